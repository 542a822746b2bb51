// Cycle counts of the warp-level factor/solve routines of lm_kernel.cuh, one warp on one SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lsqfit_b200/csrc -o tools/micro/chol_probe tools/micro/chol_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "lm_kernel.cuh"
using namespace b200lm;

template <int NP>
__global__ void probe(long long* cyc, double* out, int reps) {
    constexpr int LDA = NP | 1;
    __shared__ double A[NP * LDA], dsc[NP], LT0[NP * LDA], LT1[NP * LDA], idg[NP], colb[64];
    const int lane = threadIdx.x & 31;
    // SPD test matrix: diagonally dominant
    for (int e = lane; e < NP * LDA; e += 32) {
        const int i = e / LDA, j = e % LDA;
        A[e] = (i == j) ? 2.0 + 0.1 * i : 0.5 / (1.0 + abs(i - j));
    }
    if (lane < NP) dsc[lane] = 1.0 / sqrt(2.0 + 0.1 * lane);
    __syncwarp();
    double p = 0, res[3], pn, w2; bool ok;
    double gh = 0.3 + 0.01 * lane;
    // warm-up
    factor_solve<NP, LDA>(A, dsc, LT0, idg, colb, lane, 0.1, gh, false, &p, res);
    if (NP <= 16) factor_solve2<(NP <= 16 ? NP : 16), LDA>(A, dsc, LT0, LT1, colb, lane, 0.1, 0.2, gh, &p, &pn, &w2, &ok);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        factor_solve<NP, LDA>(A, dsc, LT0, idg, colb, lane, 0.1 + 1e-3 * r, gh + p * 1e-9, false, &p, res);
    }
    long long t1 = clock64();
    if (NP <= 16) {
        for (int r = 0; r < reps; ++r) {
            factor_solve2<(NP <= 16 ? NP : 16), LDA>(A, dsc, LT0, LT1, colb, lane, 0.1 + 1e-3 * r, 0.2, gh + p * 1e-9, &p, &pn, &w2, &ok);
        }
    }
    long long t2 = clock64();
    if (lane == 0) { cyc[0] = (t1 - t0) / reps; cyc[1] = (t2 - t1) / reps; }
    out[lane] = p + res[0] + res[1];
}

template <int NP> void run(long long* cyc, double* out) {
    probe<NP><<<1, 32>>>(cyc, out, 200);
    cudaDeviceSynchronize();
    long long c[2]; cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
    double o[32]; cudaMemcpy(o, out, 256, cudaMemcpyDeviceToHost);
    printf("{\"np\": %d, \"factor_solve_cycles\": %lld, \"factor_solve2_cycles\": %lld, \"check\": %.12g}\n", NP, c[0], c[1], o[0]);
}
int main() {
    long long* cyc; double* out;
    cudaMalloc(&cyc, 16); cudaMalloc(&out, 256);
    run<4>(cyc, out); run<6>(cyc, out); run<8>(cyc, out); run<12>(cyc, out); run<16>(cyc, out); run<19>(cyc, out);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
