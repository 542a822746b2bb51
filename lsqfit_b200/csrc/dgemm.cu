// Launcher + C entry point of the DMMA FP64 GEMM (csrc/dgemm.cuh).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200lm.h"
#include "handle.h"
#include "dgemm.cuh"

namespace b200lm {

template <bool A_KC, bool B_KC, bool ALIGNED>
static cudaError_t launch(const GemmArgs& g, int batch, cudaStream_t s) {
    const size_t smem = dgemm_smem_bytes<A_KC, B_KC>();
    cudaError_t e = cudaFuncSetAttribute(dgemm_kernel<A_KC, B_KC, ALIGNED>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        GemmArgs h = g;
        h.A += (size_t)b0 * g.sA; h.B += (size_t)b0 * g.sB; h.C += (size_t)b0 * g.sC;
        const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
        dim3 grid((g.N + GN - 1) / GN, (g.M + GM - 1) / GM, nb);
        dgemm_kernel<A_KC, B_KC, ALIGNED><<<grid, GTHREADS, smem, s>>>(h);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <bool A_KC, bool B_KC>
static cudaError_t launch_al(const GemmArgs& g, int batch, bool aligned, cudaStream_t s) {
    return aligned ? launch<A_KC, B_KC, true>(g, batch, s) : launch<A_KC, B_KC, false>(g, batch, s);
}

cudaError_t dgemm(bool transA, bool transB, int batch, int M, int N, int K, double alpha,
                  const double* A, long long sA, int lda, const double* B, long long sB, int ldb,
                  double beta, double* C, long long sC, int ldc, cudaStream_t stream, bool lower) {
    if (batch <= 0 || M <= 0 || N <= 0) return cudaSuccess;
    GemmArgs g;
    g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
    g.A = A; g.sA = sA; g.lda = lda; g.B = B; g.sB = sB; g.ldb = ldb; g.C = C; g.sC = sC; g.ldc = ldc;
    g.lower = lower ? 1 : 0;
    if (M <= 32 && batch >= 64 && !lower) {
        // many small products: the small-tile kernel
        for (int b0 = 0; b0 < batch; b0 += 65535) {
            GemmArgs h = g;
            h.A += (size_t)b0 * g.sA; h.B += (size_t)b0 * g.sB; h.C += (size_t)b0 * g.sC;
            const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
            dim3 grid((N + SM_N - 1) / SM_N, (M + SM_M - 1) / SM_M, nb);
            dgemm_small_kernel<<<grid, SM_THREADS, 0, stream>>>(h, transA ? 1 : 0, transB ? 1 : 0);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    const bool aligned = ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && (lda % 2 == 0) && (ldb % 2 == 0) &&
                         (sA % 2 == 0) && (sB % 2 == 0);
    // A_KC: A stored [M][K]  <=> !transA ;  B_KC: B stored [N][K] <=> transB
    if (!transA && !transB) return launch_al<true, false>(g, batch, aligned, stream);
    if (!transA && transB)  return launch_al<true, true>(g, batch, aligned, stream);
    if (transA && !transB)  return launch_al<false, false>(g, batch, aligned, stream);
    return launch_al<false, true>(g, batch, aligned, stream);
}

}  // namespace b200lm

using namespace b200lm;

extern "C" int b200lm_dgemm(int device, int transA, int transB, int batch, int M, int N, int K, double alpha,
                            const double* d_A, long long sA, int lda, const double* d_B, long long sB, int ldb,
                            double beta, double* d_C, long long sC, int ldc, void* stream) {
    if (!d_A || !d_B || !d_C || M < 0 || N < 0 || K < 0 || batch < 0)
        return set_error(nullptr, B200LM_EINVAL, "bad dgemm argument");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    e = dgemm(transA != 0, transB != 0, batch, M, N, K, alpha, d_A, sA, lda, d_B, sB, ldb, beta, d_C, sC, ldc,
              (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "dgemm");
    return B200LM_OK;
}
