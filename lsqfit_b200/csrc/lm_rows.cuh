// Data-parallel evaluation of ONE fit's model over all of its data rows -- the building blocks of the single-fit
// paths (lsqfit_b200/dense.py), where the parallelism is over rows instead of over fits:
//   model_rows_kernel<F>    un-whitened rows  G[i][:] = df/dp (i),  delta[i] = f_i(p) - y_i  for any registered functor
//                           (fcn(x, p) on GVars + the delta of reference src/lsqfit/_utilities.pyx:76-77); the dense
//                           GEMMs then apply the whitening
//   normal_diag_kernel<F>   uncorrelated data (1x1 weights, reference src/lsqfit/_utilities.pyx:85-89): J^T J, J^T r and
//                           r^T r of w_i (f_i - y_i) in ONE pass over the rows -- x, y, w are read once (24 B per
//                           row for nx = 1: HBM bound), every thread keeps the packed upper triangle in registers, and
//                           the reduction is deterministic (fixed-order sums: warp shuffle tree, per-CTA partial rows
//                           in global memory, summed in index order by the last CTA to arrive -- one launch)
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

// rows per thread and trip of the row loop: the loads of ROWS_UNROLL independent rows are issued before their
// arithmetic (the loop body has no stores, so the compiler hoists them), which keeps ~4x the bytes in flight per thread
// (the three knobs are compile-time so that tools/micro/nd_bench.cu can sweep them; defaults = the measured best.
// 2e6 rows x 3 parameters from HBM, us per launch: unroll 1 / 256 threads / 4 CTAs per SM 25.1; 4/256/4 26.8; 8/256/4 24.8;
// 4/256/8 32.9; 4/128/8 29.0; 4/512/2 21.8 -- fewer, larger CTAs win: the per-CTA reduction of the accumulators is a
// third of the launch.  Those figures are for the dual-number form of the model (~50 FP64 instructions per 24-byte
// row, AT the machine's ridge of 37 TFLOP/s : 6.5 TB/s = 5.7 flop/B; 18.7 us with the rows in L2); with the
// hand-written gradient (OffsetExpModel, ~28 instructions) 4/512/2 takes 20.6 us from HBM and 16.7 us from L2 -- what
// is left is the latency of a launch this short (13 rows per thread), not either roofline.)
#ifndef B200LM_ND_UNROLL
#define B200LM_ND_UNROLL 4
#endif
#ifndef B200LM_ND_THREADS
#define B200LM_ND_THREADS 512
#endif
#ifndef B200LM_ND_GRIDMUL
#define B200LM_ND_GRIDMUL 2
#endif
constexpr int ROWS_UNROLL = B200LM_ND_UNROLL;
constexpr int ND_THREADS = B200LM_ND_THREADS;

template <class F>
__global__ void __launch_bounds__(ND_THREADS) normal_diag_kernel(int ny, int nx, const double* __restrict__ x,
                                                          const double* __restrict__ p, const double* __restrict__ y,
                                                          const double* __restrict__ w, double* partial,
                                                          unsigned int* done, double* __restrict__ out) {
    constexpr int NP = F::NP, NACC = NormalAccLayout<NP>::NACC;
    __shared__ double ps[NP];
    __shared__ double red[ND_THREADS / 32][NACC];
    __shared__ int s_last;
    for (int j = threadIdx.x; j < NP; j += blockDim.x) ps[j] = p[j];
    __syncthreads();
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // (a thread visits its rows in the same order as the plain grid-stride loop: the sums do not depend on the unrolling)
    auto row = [&](long long r_, double wi, double yi) {
        double g[NP];
        const double f = F::value_grad(x + r_ * nx, (int)r_, ps, wi, g);
        const double r = wi * (f - yi);
        int k = 0;
#pragma unroll
        for (int a = 0; a < NP; ++a)
#pragma unroll
            for (int b = a; b < NP; ++b) { acc[k] = fma(g[a], g[b], acc[k]); ++k; }
#pragma unroll
        for (int a = 0; a < NP; ++a) acc[k + a] = fma(g[a], r, acc[k + a]);
        acc[NACC - 1] = fma(r, r, acc[NACC - 1]);
    };
    for (; i + (ROWS_UNROLL - 1) * stride < ny; i += ROWS_UNROLL * stride) {
        double wi[ROWS_UNROLL], yi[ROWS_UNROLL];
#pragma unroll
        for (int u = 0; u < ROWS_UNROLL; ++u) { wi[u] = __ldg(w + i + u * stride); yi[u] = __ldg(y + i + u * stride); }
#pragma unroll
        for (int u = 0; u < ROWS_UNROLL; ++u) row(i + u * stride, wi[u], yi[u]);
    }
    for (; i < ny; i += stride) row(i, w[i], y[i]);
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
        for (int q = 0; q < ND_THREADS / 32; ++q) v += red[q][k];
        partial[(size_t)blockIdx.x * NACC + k] = v;
    }
    // ---- the last CTA to arrive adds the per-CTA rows in a FIXED order (deterministic results): one warp per accumulator,
    // lane l adds rows l, l + 32, ... in order, then a shuffle tree.  No second launch.
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(done, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int nparts = gridDim.x;
    for (int k = warp; k < NACC; k += (int)(blockDim.x >> 5)) {
        double v = 0.0;
        for (int q = lane; q < nparts; q += 32) v += __ldcg(partial + (size_t)q * NACC + k);
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) out[k] = v;
    }
    if (threadIdx.x == 0) *done = 0u;             // ready for the next launch on this handle
}

// partial: max_parts rows of NACC doubles followed by the arrival counter (an unsigned int, zero between launches)
template <class F>
cudaError_t launch_normal_diag(int ny, int nx, const double* x, const double* p, const double* y, const double* w,
                               double* partial, int max_parts, double* out, int sm_count, cudaStream_t stream) {
    constexpr int NACC = NormalAccLayout<F::NP>::NACC;
    int grid = (ny + ND_THREADS * ROWS_UNROLL - 1) / (ND_THREADS * ROWS_UNROLL);   // >= one full unrolled trip per thread
    if (grid > B200LM_ND_GRIDMUL * sm_count) grid = B200LM_ND_GRIDMUL * sm_count;
    if (grid > max_parts) grid = max_parts;
    if (grid < 1) grid = 1;
    unsigned int* done = reinterpret_cast<unsigned int*>(partial + (size_t)max_parts * NACC);
    normal_diag_kernel<F><<<grid, ND_THREADS, 0, stream>>>(ny, nx, x, p, y, w, partial, done, out);
    return cudaGetLastError();
}

}  // namespace b200lm
