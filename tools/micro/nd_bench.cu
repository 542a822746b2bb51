// Micro-benchmark of normal_diag_kernel (lsqfit_b200/csrc/lm_rows.cuh) on the examples/uncorrelated.py shape:
// 2e6 rows, 3 parameters, 24 B per row.  SIX data sets (288 MB > the 126 MB L2) are visited round-robin, so every
// launch reads its rows from HBM; the time per launch is the total over 30 back-to-back launches / 30 (CUDA events
// on the launching stream), plus one L2-warm figure for comparison.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DB200LM_ND_UNROLL=4 -DB200LM_ND_THREADS=256
//        -DB200LM_ND_GRIDMUL=4 -I include -o tools/micro/nd_bench tools/micro/nd_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../lsqfit_b200/csrc/lm_rows.cuh"
#include "../../lsqfit_b200/csrc/functors.cuh"

#ifdef ND_BENCH_DUAL
typedef b200lm::ADFunctor<b200lm::OffsetExpBody, 3> OffsetExp;      // the dual-number form (before)
#else
typedef b200lm::OffsetExpModel OffsetExp;
#endif
using b200lm::launch_normal_diag;
using b200lm::ND_MAX_PARTS_PER_SM;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main() {
    const int ny = 2000000, NSET = 6, NACC = 10;
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sm = pr.multiProcessorCount;
    std::vector<double> hx(ny), hy(ny), hw(ny);
    for (int i = 0; i < ny; ++i) {
        hx[i] = 0.2 + 1.8 * i / (ny - 1.0);
        hy[i] = 0.5 + 0.4 * ::exp(-0.7 * hx[i]) + 1e-3 * ::sin(12.9898 * i);
        hw[i] = 1e3;
    }
    double *dx[NSET], *dy[NSET], *dw[NSET], *dp, *dpart, *dout;
    for (int s = 0; s < NSET; ++s) {
        CK(cudaMalloc(&dx[s], ny * sizeof(double))); CK(cudaMalloc(&dy[s], ny * sizeof(double))); CK(cudaMalloc(&dw[s], ny * sizeof(double)));
        CK(cudaMemcpy(dx[s], hx.data(), ny * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dy[s], hy.data(), ny * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dw[s], hw.data(), ny * sizeof(double), cudaMemcpyHostToDevice));
    }
    const int max_parts = ND_MAX_PARTS_PER_SM * sm;
    const double hp[3] = {0.45, 0.35, 0.8};
    CK(cudaMalloc(&dp, 3 * sizeof(double))); CK(cudaMemcpy(dp, hp, sizeof hp, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dpart, (size_t)max_parts * NACC * sizeof(double) + 8)); CK(cudaMemset(dpart, 0, (size_t)max_parts * NACC * sizeof(double) + 8));
    CK(cudaMalloc(&dout, NACC * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 12; ++i) CK(launch_normal_diag<OffsetExp>(ny, 1, dx[i % NSET], dp, dy[i % NSET], dw[i % NSET], dpart, max_parts, dout, sm, 0));
    CK(cudaDeviceSynchronize());
    double out0[NACC], out1[NACC];
    CK(cudaMemcpy(out0, dout, sizeof out0, cudaMemcpyDeviceToHost));
    float best = 1e9f, worst = 0.f;
    for (int rep = 0; rep < 5; ++rep) {
        const int NL = 30;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < NL; ++i) CK(launch_normal_diag<OffsetExp>(ny, 1, dx[i % NSET], dp, dy[i % NSET], dw[i % NSET], dpart, max_parts, dout, sm, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= NL;
        if (ms < best) best = ms;
        if (ms > worst) worst = ms;
    }
    CK(cudaMemcpy(out1, dout, sizeof out1, cudaMemcpyDeviceToHost));
    bool same = true;
    for (int k = 0; k < NACC; ++k) same = same && out0[k] == out1[k];
    // L2-warm: the same data set every launch
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 30; ++i) CK(launch_normal_diag<OffsetExp>(ny, 1, dx[0], dp, dy[0], dw[0], dpart, max_parts, dout, sm, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float warm; CK(cudaEventElapsedTime(&warm, e0, e1));
    warm /= 30;
    printf("{\"unroll\": %d, \"threads\": %d, \"gridmul\": %d, \"sms\": %d, \"us_per_launch_hbm\": %.2f, \"us_worst\": %.2f, \"gbs\": %.1f, "
           "\"us_l2_warm\": %.2f, \"deterministic\": %s, \"chi2\": %.10g}\n",
           B200LM_ND_UNROLL, B200LM_ND_THREADS, B200LM_ND_GRIDMUL, sm, 1e3 * best, 1e3 * worst, 24.0 * ny / (best * 1e-3) / 1e9,
           1e3 * warm, same ? "true" : "false", out0[NACC - 1]);
    return 0;
}
