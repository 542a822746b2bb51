// FP64 peak micro-benchmark for the roofline denominator (MEASURED_PEAKS.json has no FP64
// figure).  Measures, with CUDA events after warm-up:
//   dfma      : independent DFMA chains (vector FP64 pipe)
//   dmma884   : mma.sync.aligned.m8n8k4.f64   (FP64 tensor path)
//   dmma1684 / dmma1688 / dmma16816 : the sm_90+ shapes m16n8k4 / k8 / k16
//   mix       : half the warps DFMA, half DMMA (do the two paths share a pipe?)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak profiles/fp64_peak.cu
// Prints one JSON object.
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double* out, double a, double b) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double* c, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void k_dmma884(double* out, double a, double b) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mma884(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma1684(double* out, double a, double b) {
    double c[16], av[2] = {a, a + 1};
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mma1684(c + 4 * i, av, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma1688(double* out, double a, double b) {
    double c[16], av[4] = {a, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mma1688(c + 4 * i, av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma16816(double* out, double a, double b) {
    double c[16], av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = a + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) bv[i] = b + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mma16816(c + 4 * i, av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// even warps DFMA, odd warps DMMA m8n8k4
__global__ void k_mix(double* out, double a, double b) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    if ((threadIdx.x >> 5) & 1) {
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) mma884(c[2 * i], c[2 * i + 1], a, b);
        }
    } else {
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static double run(K kern, int grid, int block, double* out, double flops_per_thread_iter, double* ms_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<<<grid, block>>>(out, 0.999999, 1e-6);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        kern<<<grid, block>>>(out, 0.999999, 1e-6);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    *ms_out = best;
    return flops_per_thread_iter * (double)ITERS * grid * block / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int block = 256, grid = sms * 8;
    double* out;
    CHECK(cudaMalloc(&out, (size_t)grid * block * sizeof(double)));
    double ms;
    // per thread per iteration: dfma 16 FMA = 32 flop; dmma884: 8 mma x 512 flop / 32 lanes = 128;
    // m16n8k4: 4 x 1024/32 = 128; m16n8k8: 4 x 2048/32 = 256; m16n8k16: 4 x 4096/32 = 512
    const double t_dfma = run(k_dfma, grid, block, out, 32.0, &ms);
    const double t_884 = run(k_dmma884, grid, block, out, 128.0, &ms);
    const double t_1684 = run(k_dmma1684, grid, block, out, 128.0, &ms);
    const double t_1688 = run(k_dmma1688, grid, block, out, 256.0, &ms);
    const double t_16816 = run(k_dmma16816, grid, block, out, 512.0, &ms);
    const double t_mix = run(k_mix, grid, block, out, 0.5 * 32.0 + 0.5 * 128.0, &ms);
    CHECK(cudaGetLastError());
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double best = t_dfma;
    if (t_884 > best) best = t_884;
    if (t_1684 > best) best = t_1684;
    if (t_1688 > best) best = t_1688;
    if (t_16816 > best) best = t_16816;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"dfma_tflops\": %.3f, \"dmma_m8n8k4_tflops\": %.3f, "
           "\"dmma_m16n8k4_tflops\": %.3f, \"dmma_m16n8k8_tflops\": %.3f, \"dmma_m16n8k16_tflops\": %.3f, "
           "\"mix_dfma_dmma_tflops\": %.3f, \"fp64_tflops\": %.3f, "
           "\"how\": \"profiles/fp64_peak.cu: register-resident FMA / mma.sync f64 chains, %d CTAs x %d threads, best of 5, CUDA events\"}\n",
           prop.name, sms, clk, t_dfma, t_884, t_1684, t_1688, t_16816, t_mix, best, grid, block);
    return 0;
}
