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
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, int lda, int k0, int bs, double shift,
                                                         double* __restrict__ Linv, int* __restrict__ info) {
    extern __shared__ double sm[];
    double* S = sm;                 // working copy, later the inverse
    double* L = sm + PB * PBP;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < PB * PB; idx += 256) {
        const int i = idx / PB, j = idx % PB;
        double v = 0.0;
        if (i < bs && j <= i) v = A[(size_t)(k0 + i) * lda + k0 + j] + (i == j ? shift : 0.0);
        S[i * PBP + j] = v;
        L[i * PBP + j] = 0.0;
    }
    __syncthreads();
    const int ti = tid >> 4, tk = tid & 15;
    for (int j = 0; j < bs; ++j) {
        double d = S[j * PBP + j];
        if (!(d > 0.0) || !isfinite(d)) {
            if (tid == 0) atomicCAS(info, 0, k0 + j + 1);
            d = 1.0;
        }
        const double inv = 1.0 / d, rs = 1.0 / sqrt(d);
        // column j of L (from the not-yet-overwritten S), by the first bs threads
        if (tid < bs && tid >= j) L[tid * PBP + j] = tid == j ? sqrt(d) : S[tid * PBP + j] * rs;
        // trailing update of the lower triangle
        for (int i = j + 1 + ti; i < bs; i += 16) {
            const double lij = S[i * PBP + j] * inv;
            for (int k = j + 1 + tk; k <= i; k += 16) S[i * PBP + k] = fma(-lij, S[k * PBP + j], S[i * PBP + k]);
        }
        __syncthreads();
    }
    // write L back
    for (int idx = tid; idx < bs * bs; idx += 256) {
        const int i = idx / bs, j = idx % bs;
        if (j <= i) A[(size_t)(k0 + i) * lda + k0 + j] = L[i * PBP + j];
    }
    // X = inv(L): column c by 4 lanes (c = tid / 4), forward substitution down the rows
    const int c = tid >> 2, part = tid & 3;
    const int c0 = (tid >> 5) * 8;          // first column of this warp
    for (int i = 0; i < PB; ++i) S[i * PBP + c] = 0.0;   // each column zeroed by its own 4 lanes (same value)
    __syncwarp();
    for (int i = c0; i < bs; ++i) {
        double s = 0.0;
        if (c < bs && i > c)
            for (int k = c + part; k < i; k += 4) s = fma(L[i * PBP + k], S[k * PBP + c], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (part == 0 && c < bs && i >= c) S[i * PBP + c] = ((i == c ? 1.0 : 0.0) - s) / L[i * PBP + i];
        __syncwarp();
    }
    __syncthreads();
    for (int idx = tid; idx < PB * PB; idx += 256) Linv[idx] = S[(idx / PB) * PBP + (idx % PB)];
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
