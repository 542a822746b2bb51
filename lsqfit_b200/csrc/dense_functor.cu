// Dense residual/Jacobian rows of the multi-exponential correlator for LARGE parameter counts
// (BASELINE config 5: K = 1000 exponentials, np = 2000, ny = 5000), where one fit spans the whole
// GPU instead of one warp.  Writes the un-whitened model rows that the DMMA GEMM then multiplies by
// the whitening matrix:
//     G[i][k] = exp(-E_k t_i),  G[i][K+k] = -a_k t_i exp(-E_k t_i),  delta_i = sum_k a_k exp(-E_k t_i) - y_i
// (the model of reference examples/y-vs-x.py:58-61; chiv's delta of src/lsqfit/_utilities.pyx:76-77).
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/b200lm.h"
#include "handle.h"

namespace b200lm {

template <bool WITH_G>
__global__ void __launch_bounds__(256) multiexp_dense_kernel(int ny, int K, const double* __restrict__ t,
                                                             const double* __restrict__ p, const double* __restrict__ y,
                                                             double* __restrict__ G, int ld, double* __restrict__ delta) {
    __shared__ double red[256];
    const int i = blockIdx.x;
    if (i >= ny) return;
    const double ti = t[i];
    double* row = WITH_G ? G + (size_t)i * ld : nullptr;
    double f = 0.0;
    for (int k = threadIdx.x; k < K; k += 256) {
        const double a = p[k], e = exp(-p[K + k] * ti);
        if (WITH_G) {
            row[k] = e;
            row[K + k] = -ti * (a * e);
        }
        f = fma(a, e, f);
    }
    red[threadIdx.x] = f;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) delta[i] = red[0] - y[i];
}

}  // namespace b200lm

using namespace b200lm;

extern "C" int b200lm_multiexp_dense(int device, int ny, int K, const double* d_t, const double* d_p,
                                     const double* d_y, double* d_G, int ld, double* d_delta, void* stream) {
    if (ny <= 0 || K <= 0 || !d_t || !d_p || !d_y || !d_delta || (d_G && ld < 2 * K))
        return set_error(nullptr, B200LM_EINVAL, "bad multiexp_dense argument");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    if (d_G)
        multiexp_dense_kernel<true><<<ny, 256, 0, (cudaStream_t)stream>>>(ny, K, d_t, d_p, d_y, d_G, ld, d_delta);
    else
        multiexp_dense_kernel<false><<<ny, 256, 0, (cudaStream_t)stream>>>(ny, K, d_t, d_p, d_y, nullptr, 0, d_delta);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "multiexp_dense");
    return B200LM_OK;
}
