// Dependent-issue latencies of the instructions on the LM kernel's critical path (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu ; ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void probe(double* out, long long* cyc, double seed) {
    __shared__ double sm[1024];
    double x = seed + threadIdx.x, y = 1.0000001, c1 = 0.5;
    float f = (float)seed;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    int idx = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (MODE == 0) x = fma(x, y, 1e-9);                                 // DFMA chain
        if (MODE == 1) x = x * y;                                           // DMUL chain
        if (MODE == 2) x = x + y;                                           // DADD chain
        if (MODE == 3) dmma(x, c1, y, y);                                   // DMMA chain
        if (MODE == 4) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);       // 64-bit shuffle chain
        if (MODE == 5) { f = (float)x; x = (double)f + 1e-9; }              // F2F both ways + DADD
        if (MODE == 6) { f = rsqrtf(f) + 1.0f; }                            // MUFU.RSQ + FADD
        if (MODE == 7) { idx = (int)sm[idx] + ((idx + 1) & 31) - 1; }       // LDS + F2I dependent
        if (MODE == 8) { x = sm[((int)(__double_as_longlong(x) & 31))]; }   // LDS.64 dependent via bits
        if (MODE == 9) { x = 1.0 / x + 1.5; }                               // fp64 division
        if (MODE == 10) { x = sqrt(x) + 1.5; }                              // fp64 sqrt
        if (MODE == 11) { x = exp(-x) + 1.0; }                              // fp64 exp
        if (MODE == 12) { asm volatile("bar.sync 1, 128;"); }                // named barrier, 4 warps
        if (MODE == 13) { __syncwarp(); x = fma(x, y, 1e-9); }
        if (MODE == 14) { sm[threadIdx.x] = x; __syncwarp(); x = sm[threadIdx.x ^ 1] + 1e-9; }   // STS -> LDS round trip
        if (MODE == 15) { x = rsqrt(x) + 1.5; }                             // fp64 rsqrt
        if (MODE == 16) { x += __shfl_xor_sync(0xffffffffu, x, 16); x += __shfl_xor_sync(0xffffffffu, x, 8);
                          x += __shfl_xor_sync(0xffffffffu, x, 4); x += __shfl_xor_sync(0xffffffffu, x, 2);
                          x += __shfl_xor_sync(0xffffffffu, x, 1); x *= 1e-3; }                  // warp_sum
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + f + idx + c1;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char* name, int threads, double* out, long long* cyc) {
    probe<MODE><<<1, threads>>>(out, cyc, 1.25);
    cudaDeviceSynchronize();
    probe<MODE><<<1, threads>>>(out, cyc, 1.25);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("{\"op\": \"%s\", \"threads\": %d, \"cycles_per_op\": %.1f}\n", name, threads, (double)c / N);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    run<0>("DFMA dependent", 32, out, cyc);
    run<1>("DMUL dependent", 32, out, cyc);
    run<2>("DADD dependent", 32, out, cyc);
    run<3>("DMMA m8n8k4 dependent", 32, out, cyc);
    run<4>("SHFL.64 dependent", 32, out, cyc);
    run<5>("F2F d->f->d + DADD", 32, out, cyc);
    run<6>("MUFU.RSQ + FADD (fp32)", 32, out, cyc);
    run<7>("LDS + F2I dependent", 32, out, cyc);
    run<8>("LDS.64 dependent", 32, out, cyc);
    run<9>("fp64 1/x + DADD", 32, out, cyc);
    run<10>("fp64 sqrt + DADD", 32, out, cyc);
    run<11>("fp64 exp + DADD", 32, out, cyc);
    run<12>("bar.sync 128 (4 warps)", 128, out, cyc);
    run<13>("syncwarp + DFMA", 32, out, cyc);
    run<14>("STS+syncwarp+LDS+DADD", 32, out, cyc);
    run<15>("fp64 rsqrt + DADD", 32, out, cyc);
    run<16>("warp_sum (5 shfl.64 + DADD) + DMUL", 32, out, cyc);
    // throughput: 4 warps on one SM each running an independent chain of the same op
    run<0>("DFMA dependent, 4 warps", 128, out, cyc);
    run<3>("DMMA dependent, 4 warps", 128, out, cyc);
    run<0>("DFMA dependent, 16 warps", 512, out, cyc);
    run<3>("DMMA dependent, 16 warps", 512, out, cyc);
    run<3>("DMMA dependent, 32 warps", 1024, out, cyc);
    run<0>("DFMA dependent, 32 warps", 1024, out, cyc);
    return 0;
}
