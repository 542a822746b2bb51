// Data-parallel evaluation of ONE fit's model over all of its data rows -- the building blocks of the single-fit
// paths (lsqfit_b200/dense.py), where the parallelism is over rows instead of over fits:
//   model_rows_kernel<F>    un-whitened rows  G[i][:] = df/dp (i),  delta[i] = f_i(p) - y_i  for any registered functor
//                           (fcn(x, p) on GVars + the delta of reference src/lsqfit/_utilities.pyx:76-77); the dense
//                           GEMMs then apply the whitening
//   normal_diag_kernel<F>   uncorrelated data (1x1 weights, reference src/lsqfit/_utilities.pyx:85-89): J^T J, J^T r and
//                           r^T r of w_i (f_i - y_i) in ONE pass over the rows -- x, y, w are read once (24 B per
//                           row for nx = 1: HBM bound), every thread keeps the packed upper triangle in registers, and
//                           the reduction is deterministic (fixed-order sums: warp shuffle tree, per-CTA partial rows
//                           in global memory, summed in index order by sum_partials_kernel)
// examples/uncorrelated.py:30-41 of the reference (3 parameters, 5e4 ... 2e6 points) is the shape this is for.
#pragma once
#include "lm_kernel.cuh"

namespace b200lm {

template <class F>
__global__ void __launch_bounds__(256) model_rows_kernel(int ny, int nx, const double* __restrict__ x,
                                                         const double* __restrict__ p, const double* __restrict__ y,
                                                         double* __restrict__ G, int ld, double* __restrict__ delta) {
    constexpr int NP = F::NP;
    __shared__ double ps[NP];
    for (int j = threadIdx.x; j < NP; j += blockDim.x) ps[j] = p[j];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ny; i += (long long)gridDim.x * blockDim.x) {
        double f;
        if (G) f = F::value_grad(x + i * nx, (int)i, ps, 1.0, G + i * ld);
        else f = F::value(x + i * nx, (int)i, ps);
        delta[i] = f - y[i];
    }
}

template <class F>
cudaError_t launch_model_rows(int ny, int nx, const double* x, const double* p, const double* y, double* G, int ld,
                              double* delta, int sm_count, cudaStream_t stream) {
    int grid = (ny + 255) / 256;
    if (grid > 8 * sm_count) grid = 8 * sm_count;
    model_rows_kernel<F><<<grid, 256, 0, stream>>>(ny, nx, x, p, y, G, ld, delta);
    return cudaGetLastError();
}

template <int NP> struct NormalAccLayout {
    static constexpr int NTRI = NP * (NP + 1) / 2;
    static constexpr int NACC = NTRI + NP + 1;        // packed upper triangle (row-major, i <= j) | J^T r | r^T r
};

template <class F>
__global__ void __launch_bounds__(256) normal_diag_kernel(int ny, int nx, const double* __restrict__ x,
                                                          const double* __restrict__ p, const double* __restrict__ y,
                                                          const double* __restrict__ w, double* __restrict__ partial) {
    constexpr int NP = F::NP, NACC = NormalAccLayout<NP>::NACC;
    __shared__ double ps[NP];
    __shared__ double red[8][NACC];
    for (int j = threadIdx.x; j < NP; j += blockDim.x) ps[j] = p[j];
    __syncthreads();
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ny; i += (long long)gridDim.x * blockDim.x) {
        double g[NP];
        const double wi = w[i];
        const double f = F::value_grad(x + i * nx, (int)i, ps, wi, g);
        const double r = wi * (f - y[i]);
        int k = 0;
#pragma unroll
        for (int a = 0; a < NP; ++a)
#pragma unroll
            for (int b = a; b < NP; ++b) { acc[k] = fma(g[a], g[b], acc[k]); ++k; }
#pragma unroll
        for (int a = 0; a < NP; ++a) acc[k + a] = fma(g[a], r, acc[k + a]);
        acc[NACC - 1] = fma(r, r, acc[NACC - 1]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < NACC; k += blockDim.x) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += red[q][k];
        partial[(size_t)blockIdx.x * NACC + k] = v;
    }
}

__global__ void sum_partials_kernel(int nparts, int nacc, const double* __restrict__ partial, double* __restrict__ out);

template <class F>
cudaError_t launch_normal_diag(int ny, int nx, const double* x, const double* p, const double* y, const double* w,
                               double* partial, int max_parts, double* out, int sm_count, cudaStream_t stream) {
    constexpr int NACC = NormalAccLayout<F::NP>::NACC;
    int grid = (ny + 1023) / 1024;                   // >= 4 rows per thread: the accumulator reduction is amortised
    if (grid > 4 * sm_count) grid = 4 * sm_count;
    if (grid > max_parts) grid = max_parts;
    if (grid < 1) grid = 1;
    normal_diag_kernel<F><<<grid, 256, 0, stream>>>(ny, nx, x, p, y, w, partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sum_partials_kernel<<<NACC, 32, 0, stream>>>(grid, NACC, partial, out);
    return cudaGetLastError();
}

}  // namespace b200lm
