// Blocked FP64 Cholesky and triangular solves for ONE large normal matrix (config 5: np = 2000),
// built on the DMMA GEMM (csrc/dgemm.cuh).  The batched per-warp path (lm_kernel.cuh) keeps the
// whole factorisation in registers; here the matrix is 32 MB, so the factorisation is the classic
// right-looking blocked algorithm with 64-wide panels:
//     diag block  : unblocked Cholesky + explicit inverse of the 64 x 64 triangle, one CTA, in smem
//     panel       : L21 = A21 . inv(L11)^T                      (DMMA GEMM, in place)
//     trailing    : A22 -= L21 . L21^T, lower tiles only        (DMMA GEMM)
// Triangular solves L X = B / L^T X = B walk the same panels, multiplying by the stored inverse
// diagonal blocks (no divisions, no serial substitution on the critical path).
// Replaces the per-iteration LAPACK dgesdd of the reference's fitter (src/lsqfit/_scipy.py:156-161
// -> scipy trf) and the SVD pseudo-inverse covariance (src/lsqfit/_scipy.py:171-175) on this path.
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/b200lm.h"
#include "handle.h"
#include "dgemm.cuh"

namespace b200lm {

constexpr int PB = 64;          // panel width
constexpr int PBP = PB + 1;     // smem pitch

// One CTA: S = A[k0:k0+bs, k0:k0+bs] + shift I = L L^T; writes L into A's lower triangle and
// inv(L) (zero padded to 64 x 64, row-major) to Linv.  info: first failing pivot (1-based) or 0.
// Right-looking in panels of 8 columns: warp 0 factors a panel in REGISTERS (lane = row, two rows per lane; pivots and
// the pivot row travel by shuffles -- no barrier inside a panel), then all 256 threads apply the rank-8 update to the
// trailing triangle.  16 barriers per block instead of 64, and the serial chain per column is
// shuffle -> rsqrt -> multiply -> fma (the first version -- one barrier and two divisions per column -- took 95 us per
// block, 3 ms of a 4.2 ms factorisation of n = 2000).
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, int lda, int k0, int bs, double shift,
                                                         double* __restrict__ Linv, int* __restrict__ info) {
    extern __shared__ double sm[];
    double* S = sm;                 // working copy, becomes L
    double* X = sm + PB * PBP;      // the inverse
    __shared__ double rdiag[PB];    // 1 / L_ii
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < PB * PB; idx += 256) {
        const int i = idx / PB, j = idx % PB;
        double v = 0.0;
        if (i < bs && j <= i) v = A[(size_t)(k0 + i) * lda + k0 + j] + (i == j ? shift : 0.0);
        else if (i >= bs && i == j) v = 1.0;                 // padding: identity (keeps the arithmetic finite)
        S[i * PBP + j] = v;
        X[i * PBP + j] = 0.0;
    }
    __syncthreads();
    constexpr int PW = 8;
    for (int jb = 0; jb < PB; jb += PW) {
        if (warp == 0) {
            // rows r0 = jb + lane and r1 = jb + 32 + lane of the panel's 8 columns
            const int r0 = jb + lane, r1 = jb + 32 + lane;
            double v0[PW], v1[PW];
#pragma unroll
            for (int c = 0; c < PW; ++c) {
                v0[c] = r0 < PB ? S[r0 * PBP + jb + c] : 0.0;
                v1[c] = r1 < PB ? S[r1 * PBP + jb + c] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < PW; ++c) {
                double d = __shfl_sync(0xffffffffu, v0[c], c);          // pivot: row jb + c lives in lane c, first row set
                if (!(d > 0.0) || !isfinite(d)) {
                    if (lane == 0 && jb + c < bs) atomicCAS(info, 0, k0 + jb + c + 1);
                    d = 1.0;
                }
                const double rs = rsqrt(d);
                const double l0 = lane == c ? d * rs : (lane > c ? v0[c] * rs : 0.0);
                const double l1 = v1[c] * rs;
                v0[c] = l0; v1[c] = l1;
#pragma unroll
                for (int k = c + 1; k < PW; ++k) {
                    const double lk = __shfl_sync(0xffffffffu, l0, k);    // L[jb + k][jb + c]
                    v0[k] = fma(-l0, lk, v0[k]);
                    v1[k] = fma(-l1, lk, v1[k]);
                }
            }
#pragma unroll
            for (int c = 0; c < PW; ++c) {
                if (r0 < PB) S[r0 * PBP + jb + c] = v0[c];
                if (r1 < PB) S[r1 * PBP + jb + c] = v1[c];
            }
            if (lane < PW) rdiag[jb + lane] = 1.0 / v0[lane];             // (lane c holds L[jb+c][jb+c] in v0[c])
        }
        __syncthreads();
        // trailing triangle: S[i][k] -= sum_c L[i][jb+c] L[k][jb+c],  jb + 8 <= k <= i < 64
        const int m = PB - jb - PW;
        for (int idx = tid; idx < m * m; idx += 256) {
            const int i = jb + PW + idx / m, k = jb + PW + idx % m;
            if (k <= i) {
                double acc = S[i * PBP + k];
#pragma unroll
                for (int c = 0; c < PW; ++c) acc = fma(-S[i * PBP + jb + c], S[k * PBP + jb + c], acc);
                S[i * PBP + k] = acc;
            }
        }
        __syncthreads();
    }
    // fix the diagonal reciprocals (lane c of warp 0 wrote 1 / v0[c] only for its own column: recompute for all)
    if (tid < PB) rdiag[tid] = 1.0 / S[tid * PBP + tid];
    // write L back
    for (int idx = tid; idx < bs * bs; idx += 256) {
        const int i = idx / bs, j = idx % bs;
        if (j <= i) A[(size_t)(k0 + i) * lda + k0 + j] = S[i * PBP + j];
    }
    __syncthreads();
    // X = inv(L): column c by 4 lanes (c = tid / 4), forward substitution down the rows
    const int c = tid >> 2, part = tid & 3;
    const int c0 = (tid >> 5) * 8;          // first column of this warp
    for (int i = c0; i < bs; ++i) {
        double s = 0.0;
        if (c < bs && i > c)
            for (int k = c + part; k < i; k += 4) s = fma(S[i * PBP + k], X[k * PBP + c], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0 && c < bs && i >= c) X[i * PBP + c] = ((i == c ? 1.0 : 0.0) - s) * rdiag[i];
        __syncwarp();
    }
    __syncthreads();
    for (int idx = tid; idx < PB * PB; idx += 256) {
        const int i = idx / PB, j = idx % PB;
        Linv[idx] = (i < bs && j < bs) ? X[i * PBP + j] : 0.0;
    }
}

static const size_t POTRF_SMEM = 2 * PB * PBP * sizeof(double);

// L (n x n, ld ldl) <- chol(A + shift I), lower; d_linv: ceil(n/64) blocks of 64 x 64.
cudaError_t potrf(int n, const double* A, int lda, double shift, double* L, int ldl, double* d_linv, int* d_info,
                  cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
    if (e != cudaSuccess) return e;
    if (A != L) {
        e = cudaMemcpy2DAsync(L, (size_t)ldl * 8, A, (size_t)lda * 8, (size_t)n * 8, n, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
    }
    e = cudaMemsetAsync(d_info, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    for (int k0 = 0, kb = 0; k0 < n; k0 += PB, ++kb) {
        const int bs = n - k0 < PB ? n - k0 : PB, k1 = k0 + bs;
        double* li = d_linv + (size_t)kb * PB * PB;
        potrf_diag_kernel<<<1, 256, POTRF_SMEM, s>>>(L, ldl, k0, bs, shift, li, d_info);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (k1 < n) {
            double* A21 = L + (size_t)k1 * ldl + k0;
            // L21 = A21 . inv(L11)^T   (in place: every CTA owns its 128 rows, reads finish before the epilogue)
            e = dgemm(false, true, 1, n - k1, bs, bs, 1.0, A21, 0, ldl, li, 0, PB, 0.0, A21, 0, ldl, s);
            if (e != cudaSuccess) return e;
            double* A22 = L + (size_t)k1 * ldl + k1;
            e = dgemm(false, true, 1, n - k1, n - k1, bs, -1.0, A21, 0, ldl, A21, 0, ldl, 1.0, A22, 0, ldl, s, true);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

// trans = 0: L X = B;  trans = 1: L^T X = B.   B (n x nrhs) is used as workspace and destroyed.
cudaError_t trsm(int n, int nrhs, const double* L, int ldl, const double* d_linv, int trans, double* B, int ldb,
                 double* X, int ldx, cudaStream_t s) {
    const int nblk = (n + PB - 1) / PB;
    cudaError_t e;
    if (!trans) {
        for (int kb = 0; kb < nblk; ++kb) {
            const int k0 = kb * PB, bs = n - k0 < PB ? n - k0 : PB, k1 = k0 + bs;
            const double* li = d_linv + (size_t)kb * PB * PB;
            e = dgemm(false, false, 1, bs, nrhs, bs, 1.0, li, 0, PB, B + (size_t)k0 * ldb, 0, ldb, 0.0,
                      X + (size_t)k0 * ldx, 0, ldx, s);
            if (e != cudaSuccess) return e;
            if (k1 < n) {
                e = dgemm(false, false, 1, n - k1, nrhs, bs, -1.0, L + (size_t)k1 * ldl + k0, 0, ldl,
                          X + (size_t)k0 * ldx, 0, ldx, 1.0, B + (size_t)k1 * ldb, 0, ldb, s);
                if (e != cudaSuccess) return e;
            }
        }
    } else {
        for (int kb = nblk - 1; kb >= 0; --kb) {
            const int k0 = kb * PB, bs = n - k0 < PB ? n - k0 : PB;
            const double* li = d_linv + (size_t)kb * PB * PB;
            e = dgemm(true, false, 1, bs, nrhs, bs, 1.0, li, 0, PB, B + (size_t)k0 * ldb, 0, ldb, 0.0,
                      X + (size_t)k0 * ldx, 0, ldx, s);
            if (e != cudaSuccess) return e;
            if (k0 > 0) {
                e = dgemm(true, false, 1, k0, nrhs, bs, -1.0, L + (size_t)k0 * ldl, 0, ldl, X + (size_t)k0 * ldx, 0, ldx,
                          1.0, B, 0, ldb, s);
                if (e != cudaSuccess) return e;
            }
        }
    }
    return cudaSuccess;
}

// Single right-hand side: the whole blocked substitution in ONE CTA (the GEMM-based path above costs
// 64 launches of tiny products per solve; the trust-region iteration does three solves per
// factorisation).  The vector lives in shared memory; per 64-wide panel the diagonal block is applied
// through its stored inverse, then the remaining entries are updated with the panel of L:
//   forward  (L y = b):   warp per row, lanes across the 64 panel columns  (coalesced 512-byte rows)
//   backward (L^T x = b): thread per column, walking down the 64 panel rows (coalesced across threads)
template <int TRANS>
__global__ void __launch_bounds__(1024) trsv_kernel(int n, const double* __restrict__ L, int ldl,
                                                    const double* __restrict__ linv, const double* __restrict__ b,
                                                    double* __restrict__ x) {
    extern __shared__ double tv_sm[];
    double* v = tv_sm;            // [n] running right-hand side
    double* xk = tv_sm + n;       // [64] solution of the current panel
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < n; i += 1024) v[i] = b[i];
    __syncthreads();
    const int nblk = (n + PB - 1) / PB;
    for (int q = 0; q < nblk; ++q) {
        const int kb = TRANS ? nblk - 1 - q : q;
        const int k0 = kb * PB, bs = min(PB, n - k0), k1 = k0 + bs;
        const double* li = linv + (size_t)kb * PB * PB;
        // x_k = inv(L_kk) b_k   or   inv(L_kk)^T b_k : 16 lanes per entry
        {
            const int i = tid >> 4, sub = tid & 15;
            double acc = 0.0;
            if (i < bs)
                for (int j = sub; j < bs; j += 16) acc = fma(TRANS ? li[j * PB + i] : li[i * PB + j], v[k0 + j], acc);
            for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, 16);
            if (i < bs && sub == 0) xk[i] = acc;
        }
        __syncthreads();
        if (tid < bs) x[k0 + tid] = xk[tid];
        if (!TRANS) {
            for (int i = k1 + w; i < n; i += 32) {
                const double* row = L + (size_t)i * ldl + k0;
                double acc = 0.0;
                for (int j = lane; j < bs; j += 32) acc = fma(row[j], xk[j], acc);
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) v[i] -= acc;
            }
        } else {
            for (int c = tid; c < k0; c += 1024) {
                double acc = 0.0;
                for (int j = 0; j < bs; ++j) acc = fma(L[(size_t)(k0 + j) * ldl + c], xk[j], acc);
                v[c] -= acc;
            }
        }
        __syncthreads();
    }
}

cudaError_t trsv(int n, const double* L, int ldl, const double* d_linv, int trans, const double* b, double* x,
                 cudaStream_t s) {
    const size_t smem = ((size_t)n + PB) * sizeof(double);
    cudaError_t e;
    if (trans) {
        e = cudaFuncSetAttribute(trsv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        trsv_kernel<1><<<1, 1024, smem, s>>>(n, L, ldl, d_linv, b, x);
    } else {
        e = cudaFuncSetAttribute(trsv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        trsv_kernel<0><<<1, 1024, smem, s>>>(n, L, ldl, d_linv, b, x);
    }
    return cudaGetLastError();
}

}  // namespace b200lm

using namespace b200lm;

extern "C" int b200lm_potrf(int device, int n, const double* d_A, int lda, double shift, double* d_L, int ldl,
                            double* d_linv, int* d_info, void* stream) {
    if (n <= 0 || !d_A || !d_L || !d_linv || !d_info || lda < n || ldl < n)
        return set_error(nullptr, B200LM_EINVAL, "bad potrf argument");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    e = potrf(n, d_A, lda, shift, d_L, ldl, d_linv, d_info, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "potrf");
    return B200LM_OK;
}

extern "C" int b200lm_trsm(int device, int n, int nrhs, const double* d_L, int ldl, const double* d_linv, int trans,
                           double* d_B, int ldb, double* d_X, int ldx, void* stream) {
    if (n <= 0 || nrhs <= 0 || !d_L || !d_linv || !d_B || !d_X || ldl < n || ldb < nrhs || ldx < nrhs || d_B == d_X)
        return set_error(nullptr, B200LM_EINVAL, "bad trsm argument");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    // one right-hand side that fits in shared memory: the single-CTA substitution kernel
    if (nrhs == 1 && ldb == 1 && ldx == 1 && (size_t)(n + 64) * sizeof(double) <= 200 * 1024)
        e = trsv(n, d_L, ldl, d_linv, trans, d_B, d_X, (cudaStream_t)stream);
    else
        e = trsm(n, nrhs, d_L, ldl, d_linv, trans, d_B, ldb, d_X, ldx, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "trsm");
    return B200LM_OK;
}
