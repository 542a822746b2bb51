// Post-fit covariance propagation for fit.p as dense fp64 GEMMs.
//
// Replaces nonlinear_fit._getp (reference src/lsqfit/__init__.py:897-922) and the chivw
// call inside it (src/lsqfit/_utilities.pyx:110-139):
//     D      = cov . G^T . C^-1 = (cov . J^T) . W        (np x N)   W = whitening operator
//     cov(p) = D . C . D^T                                (np x np)
// where J = W [G; I] is the whitened Jacobian at the solution (recomputed by the
// residual/Jacobian kernel), W is the dense [nchiv][N] form of yp_pdf.i_invwgts and C is
// the (svd-corrected) covariance of y(+)prior.  The reference builds D column by column in
// a Python loop over N GVars (:907-911) and each p[a] with wsum_der (:914-918).
#include <cuda_runtime.h>
#include <vector>
#include "../../include/b200lm.h"
#include "handle.h"

namespace b200lm {

constexpr int GT = 32;      // tile edge
// C[b] (M x N) = A[b] (M x K) . B[b] (K x N)        (TRANSB = false)
// C[b] (M x N) = A[b] (M x K) . B[b]^T, B (N x K)   (TRANSB = true)
// batch strides in elements; a stride of 0 shares the operand across the batch.
template <bool TRANSB>
__global__ void __launch_bounds__(256) bgemm_kernel(int M, int N, int K,
                                                    const double* __restrict__ A, long long sA, int lda,
                                                    const double* __restrict__ B, long long sB, int ldb,
                                                    double* __restrict__ C, long long sC, int ldc) {
    __shared__ double As[GT][GT + 1];
    __shared__ double Bs[GT][GT + 1];
    const int b = blockIdx.z;
    A += (size_t)b * sA; B += (size_t)b * sB; C += (size_t)b * sC;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 x 16 threads, 2 x 2 outputs each
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
    for (int k0 = 0; k0 < K; k0 += GT) {
        for (int e = threadIdx.x; e < GT * GT; e += 256) {
            const int r = e / GT, c = e % GT;
            const int m = m0 + r, k = k0 + c;
            As[r][c] = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.0;
            if (TRANSB) {
                const int n = n0 + r, kk = k0 + c;           // Bs[n][k]
                Bs[r][c] = (n < N && kk < K) ? B[(size_t)n * ldb + kk] : 0.0;
            } else {
                const int kk = k0 + r, n = n0 + c;           // Bs[k][n]
                Bs[r][c] = (kk < K && n < N) ? B[(size_t)kk * ldb + n] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < GT; ++k) {
            const double a0 = As[ty][k], a1 = As[ty + 16][k];
            const double b0 = TRANSB ? Bs[tx][k] : Bs[k][tx];
            const double b1 = TRANSB ? Bs[tx + 16][k] : Bs[k][tx + 16];
            c00 = fma(a0, b0, c00); c01 = fma(a0, b1, c01);
            c10 = fma(a1, b0, c10); c11 = fma(a1, b1, c11);
        }
        __syncthreads();
    }
    const int m = m0 + ty, n = n0 + tx;
    if (m < M && n < N) C[(size_t)m * ldc + n] = c00;
    if (m < M && n + 16 < N) C[(size_t)m * ldc + n + 16] = c01;
    if (m + 16 < M && n < N) C[(size_t)(m + 16) * ldc + n] = c10;
    if (m + 16 < M && n + 16 < N) C[(size_t)(m + 16) * ldc + n + 16] = c11;
}

template <bool TRANSB>
static cudaError_t bgemm(int batch, int M, int N, int K, const double* A, long long sA, int lda,
                         const double* B, long long sB, int ldb, double* C, long long sC, int ldc,
                         cudaStream_t s) {
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
        dim3 grid((N + GT - 1) / GT, (M + GT - 1) / GT, nb);
        bgemm_kernel<TRANSB><<<grid, 256, 0, s>>>(M, N, K, A + (size_t)b0 * sA, sA, lda,
                                                  B + (size_t)b0 * sB, sB, ldb, C + (size_t)b0 * sC, sC, ldc);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace b200lm

using namespace b200lm;

extern "C" int b200lm_propagate(b200lm_handle h, int B, const double* d_x, const double* d_cov,
                                const double* d_C, double* d_D, double* d_covp, void* stream) {
    if (!h) return set_error(h, B200LM_EINVAL, "NULL handle");
    if (!h->have_weights || !h->have_const) return set_error(h, B200LM_EINVAL, "plan is not complete");
    if (B < 0 || !d_x || !d_cov || !d_D) return set_error(h, B200LM_EINVAL, "NULL required argument");
    if (d_covp && !d_C) return set_error(h, B200LM_EINVAL, "d_C is required for cov(p)");
    if (B == 0) return B200LM_OK;
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaSetDevice");
    cudaStream_t s = (cudaStream_t)stream;
    const int np = h->np, N = h->N, nchiv = h->nchiv;
    // dense whitening operator W [nchiv][N], built once per weight set
    if (!h->d_wfull) {
        std::vector<double> W((size_t)nchiv * N, 0.0);
        const int ndiag = (int)h->h_diag_idx.size();
        for (int i = 0; i < ndiag; ++i) W[(size_t)i * N + h->h_diag_idx[i]] = h->h_diag_w[i];
        std::vector<double> wt(h->wt_total);
        if (h->wt_total) {
            e = cudaMemcpy(wt.data(), h->d_blk_wt, wt.size() * sizeof(double), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) return cuda_fail(h, e, "read weights");
        }
        for (auto& b : h->h_blk)
            for (int r = 0; r < b.n_out; ++r)
                for (int j = 0; j < b.n_in; ++j)
                    W[(size_t)(b.chiv_off + r) * N + h->h_blk_idx[b.idx_off + j]] = wt[b.wt_off + (size_t)r * b.ldw + j];
        e = cudaMalloc((void**)&h->d_wfull, W.size() * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpy(h->d_wfull, W.data(), W.size() * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return cuda_fail(h, e, "upload dense whitening operator");
    }
    // scratch: J [B][nchiv][np], M [B][np][nchiv], T [B][np][N]
    const size_t nJ = (size_t)B * nchiv * np, nM = nJ, nT = (size_t)B * np * N;
    const size_t need = (nJ + nM + nT) * sizeof(double);
    if (need > h->scratch_bytes) {
        if (h->d_scratch) cudaFree(h->d_scratch);
        h->d_scratch = nullptr; h->scratch_bytes = 0;
        e = cudaMalloc((void**)&h->d_scratch, need);
        if (e != cudaSuccess) return cuda_fail(h, e, "scratch allocation");
        h->scratch_bytes = need;
    }
    double* dJ = h->d_scratch;
    double* dM = dJ + nJ;
    double* dT = dM + nM;
    // J at the solution.  The residual/Jacobian kernel needs a mean vector only for the
    // residuals, which are not used here: the zeroed head of the D buffer serves as a dummy.
    e = cudaMemsetAsync(d_D, 0, (size_t)N * sizeof(double), s);
    if (e != cudaSuccess) return cuda_fail(h, e, "memset");
    int rc = b200lm_residual_jacobian(h, B, d_x, np, d_D, 0, nullptr, dJ, nullptr, stream);
    if (rc) return rc;
    // M = cov . J^T      (np x nchiv)
    e = bgemm<true>(B, np, nchiv, np, d_cov, (long long)np * np, np, dJ, (long long)nchiv * np, np,
                    dM, (long long)np * nchiv, nchiv, s);
    // D = M . W          (np x N)
    if (e == cudaSuccess)
        e = bgemm<false>(B, np, N, nchiv, dM, (long long)np * nchiv, nchiv, h->d_wfull, 0, N,
                         d_D, (long long)np * N, N, s);
    if (e == cudaSuccess && d_covp) {
        // T = D . C ; cov(p) = T . D^T
        e = bgemm<false>(B, np, N, N, d_D, (long long)np * N, N, d_C, 0, N, dT, (long long)np * N, N, s);
        if (e == cudaSuccess)
            e = bgemm<true>(B, np, np, N, dT, (long long)np * N, N, d_D, (long long)np * N, N,
                            d_covp, (long long)np * np, np, s);
    }
    if (e != cudaSuccess) return cuda_fail(h, e, "propagate");
    h->launches += d_covp ? 4 : 2;
    h->last_stream = s;
    return B200LM_OK;
}
