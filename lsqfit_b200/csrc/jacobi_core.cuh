// Parallel (round-robin) two-sided cyclic Jacobi on a small symmetric matrix held by one CTA.
// Shared by the single-block whitening kernel (csrc/whiten.cu) and by the 2b x 2b sub-problems of
// the block-Jacobi solver for large blocks (csrc/whiten_large.cu).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace b200lm {

__device__ __forceinline__ void rr_pair(int n2, int r, int k, int& p, int& q) {
    // round-robin tournament on n2 (even) players, round r in [0, n2-1), pair k in [0, n2/2)
    const int m = n2 - 1;
    int a, b;
    if (k == 0) { a = m; b = r; }
    else { a = (r + k) % m; b = (r - k + m) % m; }
    p = min(a, b); q = max(a, b);
}

// A (n x n, leading dimension ld) is diagonalised in place, V (initialised by the caller, usually
// to the identity) accumulates the rotations: A_in = V A_out V^T.  A rotation is applied when
// |a_pq| > eps/2 sqrt(|a_pp a_qq|) (relative criterion: small eigenvalues keep high relative
// accuracy).  s_c, s_s: n/2+1 doubles of shared scratch; s_pq: 2*(n/2+1) ints; s_flag: 2 ints.
// Returns the number of sweeps that applied at least one rotation.
template <int THREADS>
__device__ int jacobi_diagonalize(double* A, double* V, int n, int ld, int max_sweeps,
                                  double* s_c, double* s_s, int* s_pq, int* s_flag) {
    const int tid = threadIdx.x;
    const int n2 = (n + 1) & ~1;            // pad to even with a phantom index n
    const int npair = n2 / 2;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (tid == 0) s_flag[1] = 0;
        __syncthreads();
        for (int r = 0; r < n2 - 1; ++r) {
            if (tid == 0) s_flag[0] = 0;
            __syncthreads();
            for (int k = tid; k < npair; k += THREADS) {
                int p, q;
                rr_pair(n2, r, k, p, q);
                double c = 1.0, s = 0.0;
                if (q < n) {
                    const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
                    if (fabs(apq) > 1.1102230246251565e-16 * sqrt(fabs(app * aqq)) && apq != 0.0) {
                        const double tau = (aqq - app) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = rsqrt(1.0 + t * t);
                        s = t * c;
                        s_flag[0] = 1;
                    }
                }
                s_c[k] = c; s_s[k] = s;
                s_pq[2 * k] = p; s_pq[2 * k + 1] = q;
            }
            __syncthreads();
            if (s_flag[0]) {
                // columns: A <- A J, V <- V J
                for (int e = tid; e < npair * n; e += THREADS) {
                    const int k = e / n, i = e % n;
                    const double s = s_s[k];
                    if (s == 0.0) continue;
                    const int p = s_pq[2 * k], q = s_pq[2 * k + 1];
                    const double c = s_c[k];
                    const double aip = A[i * ld + p], aiq = A[i * ld + q];
                    A[i * ld + p] = c * aip - s * aiq;
                    A[i * ld + q] = s * aip + c * aiq;
                    const double vip = V[i * ld + p], viq = V[i * ld + q];
                    V[i * ld + p] = c * vip - s * viq;
                    V[i * ld + q] = s * vip + c * viq;
                }
                __syncthreads();
                // rows: A <- J^T A
                for (int e = tid; e < npair * n; e += THREADS) {
                    const int k = e / n, j = e % n;
                    const double s = s_s[k];
                    if (s == 0.0) continue;
                    const int p = s_pq[2 * k], q = s_pq[2 * k + 1];
                    const double c = s_c[k];
                    const double apj = A[p * ld + j], aqj = A[q * ld + j];
                    A[p * ld + j] = c * apj - s * aqj;
                    A[q * ld + j] = s * apj + c * aqj;
                }
                if (tid == 0) s_flag[1] += 1;
            }
            __syncthreads();
        }
        if (s_flag[1] == 0) break;
        __syncthreads();
    }
    __syncthreads();
    return sweep;
}

}  // namespace b200lm
