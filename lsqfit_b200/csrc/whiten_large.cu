// svdcut whitening of ONE large correlated block (n > 512; BASELINE config 5: n = 5000) by a
// two-sided BLOCK Jacobi eigen-solver that spans the whole GPU.
//
// Replaces, at this size, the eigen-decomposition inside gvar.PDF / gvar.svd that the reference
// calls at src/lsqfit/__init__.py:1895,1898 (LAPACK on the host there); semantics as in
// csrc/whiten.cu and oracle/whiten.py.
//
// The matrix is cut into nb = ceil(n/32) block columns.  One outer sweep visits all block pairs
// (I, J) in nb-1 rounds of nb/2 disjoint pairs (round-robin tournament); per round
//   bj_diag : one CTA per pair diagonalises the (<=64)^2 sub-matrix [A_II A_IJ; A_JI A_JJ] exactly
//             with the shared-memory Jacobi of jacobi_core.cuh and leaves the orthogonal Q_p
//   bj_cols : A[:, IuJ] <- A[:, IuJ] Q_p  and  V[:, IuJ] <- V[:, IuJ] Q_p   (64-row slabs staged in smem)
//   bj_rows : A[IuJ, :] <- Q_p^T A[IuJ, :]
// until the off-diagonal mass is at rounding level.  Eigenvalues keep Jacobi's high relative
// accuracy, which the svdcut count (nmod) depends on.  Post-processing (sorting, clamp/drop, W,
// corrected covariance via the DMMA GEMM of dgemm.cuh, logdet) follows the small-block kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <numeric>
#include <vector>
#include "../../include/b200lm.h"
#include "handle.h"
#include "jacobi_core.cuh"
#include "dgemm.cuh"

namespace b200lm {

constexpr int BJ_B = 32;            // block width
constexpr int BJ_M = 2 * BJ_B;      // sub-problem size
constexpr int BJ_LD = BJ_M + 1;
constexpr int BJ_THREADS = 256;

struct BJArgs {
    double* A; double* V; int n; int ld; int nb; int nbe; int round;
    double* Q;              // [nbe/2][BJ_M][BJ_M]
    double* offsq;          // sum of squares of the off-diagonal blocks met in this sweep
    int* rotated;           // number of pairs whose sub-problem was not yet diagonal in this sweep
};

__device__ __forceinline__ bool bj_pair(const BJArgs& a, int pair, int& I, int& J, int& nI, int& nJ) {
    rr_pair(a.nbe, a.round, pair, I, J);
    if (J >= a.nb) return false;                      // phantom block: this pair idles
    nI = min(BJ_B, a.n - I * BJ_B);
    nJ = min(BJ_B, a.n - J * BJ_B);
    return true;
}
// global column/row index of local index l in the (I, J) pair
__device__ __forceinline__ int bj_gidx(int l, int I, int J, int nI) { return l < nI ? I * BJ_B + l : J * BJ_B + (l - nI); }

__global__ void __launch_bounds__(BJ_THREADS) bj_diag_kernel(const __grid_constant__ BJArgs a) {
    extern __shared__ double bj_sm[];
    double* S = bj_sm;
    double* Qs = bj_sm + BJ_M * BJ_LD;
    __shared__ double s_c[BJ_M / 2 + 1], s_s[BJ_M / 2 + 1];
    __shared__ int s_flag[2];
    __shared__ int s_pq[BJ_M + 2];
    __shared__ double s_red[BJ_THREADS];
    const int tid = threadIdx.x;
    int I, J, nI, nJ;
    double* Qg = a.Q + (size_t)blockIdx.x * BJ_M * BJ_M;
    if (!bj_pair(a, blockIdx.x, I, J, nI, nJ)) return;
    const int m = nI + nJ;
    double off = 0.0;
    for (int e = tid; e < m * m; e += BJ_THREADS) {
        const int r = e / m, c = e % m;
        const double v = a.A[(size_t)bj_gidx(r, I, J, nI) * a.ld + bj_gidx(c, I, J, nI)];
        S[r * BJ_LD + c] = v;
        Qs[r * BJ_LD + c] = (r == c) ? 1.0 : 0.0;
        if (r < nI && c >= nI) off = fma(v, v, off);
    }
    s_red[tid] = off;
    __syncthreads();
    for (int o = BJ_THREADS / 2; o > 0; o >>= 1) {
        if (tid < o) s_red[tid] += s_red[tid + o];
        __syncthreads();
    }
    if (tid == 0) atomicAdd(a.offsq, 2.0 * s_red[0]);
    // symmetrise the copy (the global matrix is symmetric up to rounding)
    for (int e = tid; e < m * m; e += BJ_THREADS) {
        const int r = e / m, c = e % m;
        if (r < c) { const double v = 0.5 * (S[r * BJ_LD + c] + S[c * BJ_LD + r]); S[r * BJ_LD + c] = v; }
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += BJ_THREADS) {
        const int r = e / m, c = e % m;
        if (r > c) S[r * BJ_LD + c] = S[c * BJ_LD + r];
    }
    __syncthreads();
    // a few inner sweeps per visit are enough: the outer iteration finishes the job, and a full
    // diagonalisation of every sub-problem is what dominated the first version (sync-latency bound)
    const int nsw = jacobi_diagonalize<BJ_THREADS>(S, Qs, m, BJ_LD, 3, s_c, s_s, s_pq, s_flag);
    if (tid == 0 && nsw > 0) atomicAdd(a.rotated, 1);      // this pair still needed rotations
    for (int e = tid; e < m * m; e += BJ_THREADS) {
        const int r = e / m, c = e % m;
        Qg[r * BJ_M + c] = Qs[r * BJ_LD + c];
    }
}

// Slab products on the FP64 tensor path.  Shared-memory pitches: 68 == 4 (mod 16) for operands read
// as A fragments (row = lane/4, k = lane%4), 72 == 8 (mod 16) for operands read as B fragments
// (k = lane%4, col = lane/4): every fragment load touches each bank pair exactly twice.
constexpr int BJ_PA = BJ_M + 4;     // 68
constexpr int BJ_PB = BJ_M + 8;     // 72

__device__ __forceinline__ void bj_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// X[rows, IuJ] <- X[rows, IuJ] . Q   (grid: x = row slab of 64, y = pair).  8 warps: warp w owns
// output rows 8w..8w+7 of the slab and all 8 column tiles (16 k-steps x 8 DMMA).
__global__ void __launch_bounds__(BJ_THREADS) bj_cols_kernel(const __grid_constant__ BJArgs a, double* X, int nrows) {
    extern __shared__ double bj_sm[];
    double* S = bj_sm;                       // [64][BJ_PA]   A operand: S[r][k]
    double* Qs = bj_sm + 64 * BJ_PA;         // [BJ_M][BJ_PB] B operand: Q[k][c]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int I, J, nI, nJ;
    if (!bj_pair(a, blockIdx.y, I, J, nI, nJ)) return;
    const int m = nI + nJ;
    const int r0 = blockIdx.x * 64;
    const int nr = min(64, nrows - r0);
    const double* Qg = a.Q + (size_t)blockIdx.y * BJ_M * BJ_M;
    for (int e = tid; e < BJ_M * BJ_M; e += BJ_THREADS) {
        const int r = e / BJ_M, c = e % BJ_M;
        Qs[r * BJ_PB + c] = (r < m && c < m) ? Qg[r * BJ_M + c] : 0.0;
    }
    for (int e = tid; e < 64 * BJ_M; e += BJ_THREADS) {
        const int r = e / BJ_M, c = e % BJ_M;
        S[r * BJ_PA + c] = (r < nr && c < m) ? X[(size_t)(r0 + r) * a.ld + bj_gidx(c, I, J, nI)] : 0.0;
    }
    __syncthreads();
    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    const double* sa = S + (8 * w + (lane >> 2)) * BJ_PA + (lane & 3);
    const double* qb = Qs + (lane & 3) * BJ_PB + (lane >> 2);
#pragma unroll 4
    for (int k4 = 0; k4 < BJ_M; k4 += 4) {
        const double af = sa[k4];
#pragma unroll
        for (int t = 0; t < 8; ++t) bj_dmma(acc[t][0], acc[t][1], af, qb[k4 * BJ_PB + 8 * t]);
    }
    __syncthreads();
    // C fragment -> slab (row 8w + lane/4, cols 8t + 2(lane%4) + {0,1})
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        double* d = S + (8 * w + (lane >> 2)) * BJ_PA + 8 * t + 2 * (lane & 3);
        d[0] = acc[t][0]; d[1] = acc[t][1];
    }
    __syncthreads();
    for (int e = tid; e < nr * m; e += BJ_THREADS) {
        const int rr = e / m, c = e % m;
        X[(size_t)(r0 + rr) * a.ld + bj_gidx(c, I, J, nI)] = S[rr * BJ_PA + c];
    }
}

// A[IuJ, cols] <- Q^T . A[IuJ, cols]   (grid: x = column slab of 64, y = pair).  Warp w owns output
// rows 8w..8w+7 (rows of Q^T = columns of Q) and all 8 column tiles.
__global__ void __launch_bounds__(BJ_THREADS) bj_rows_kernel(const __grid_constant__ BJArgs a) {
    extern __shared__ double bj_sm[];
    double* Qt = bj_sm;                      // [BJ_M][BJ_PA]  A operand: Qt[r][k] = Q[k][r]
    double* S = bj_sm + BJ_M * BJ_PA;        // [BJ_M][BJ_PB]  B operand: S[k][c]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int I, J, nI, nJ;
    if (!bj_pair(a, blockIdx.y, I, J, nI, nJ)) return;
    const int m = nI + nJ;
    const int c0 = blockIdx.x * 64;
    const int nc = min(64, a.n - c0);
    const double* Qg = a.Q + (size_t)blockIdx.y * BJ_M * BJ_M;
    for (int e = tid; e < BJ_M * BJ_M; e += BJ_THREADS) {
        const int k = e / BJ_M, r = e % BJ_M;          // coalesced read of Q[k][r]
        Qt[r * BJ_PA + k] = (r < m && k < m) ? Qg[k * BJ_M + r] : 0.0;
    }
    for (int e = tid; e < BJ_M * 64; e += BJ_THREADS) {
        const int r = e / 64, c = e % 64;
        S[r * BJ_PB + c] = (r < m && c < nc) ? a.A[(size_t)bj_gidx(r, I, J, nI) * a.ld + c0 + c] : 0.0;
    }
    __syncthreads();
    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    const double* qa = Qt + (8 * w + (lane >> 2)) * BJ_PA + (lane & 3);
    const double* sb = S + (lane & 3) * BJ_PB + (lane >> 2);
#pragma unroll 4
    for (int k4 = 0; k4 < BJ_M; k4 += 4) {
        const double af = qa[k4];
#pragma unroll
        for (int t = 0; t < 8; ++t) bj_dmma(acc[t][0], acc[t][1], af, sb[k4 * BJ_PB + 8 * t]);
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        double* d = S + (8 * w + (lane >> 2)) * BJ_PB + 8 * t + 2 * (lane & 3);
        d[0] = acc[t][0]; d[1] = acc[t][1];
    }
    __syncthreads();
    for (int e = tid; e < m * nc; e += BJ_THREADS) {
        const int rr = e / nc, cc = e % nc;
        a.A[(size_t)bj_gidx(rr, I, J, nI) * a.ld + c0 + cc] = S[rr * BJ_PB + cc];
    }
}

// corr = D cov D, V = I, D = |diag|^-1/2
__global__ void wl_init_kernel(const double* cov, int n, double* A, double* V, int ld, double* Dv) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    const int i = (int)(e / n), j = (int)(e % n);
    double di = fabs(cov[(size_t)i * n + i]), dj = fabs(cov[(size_t)j * n + j]);
    di = di == 0.0 ? 1.0 : rsqrt(di);
    dj = dj == 0.0 ? 1.0 : rsqrt(dj);
    A[(size_t)i * ld + j] = cov[e] * di * dj;
    V[(size_t)i * ld + j] = (i == j) ? 1.0 : 0.0;
    if (i == j) Dv[i] = di;
}
__global__ void wl_diag_kernel(const double* A, int n, int ld, double* val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) val[i] = A[(size_t)i * ld + i];
}
// W[r][j] = V[j][order[r]] D[j] / sqrt(used[r])  (r < nkeep), else 0
__global__ void wl_w_kernel(const double* V, int n, int ld, const int* order, const double* used,
                            const double* Dv, int nkeep, double* W) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    const int r = (int)(e / n), j = (int)(e % n);      // j fastest: coalesced writes of W's row r
    W[(size_t)r * n + j] = r < nkeep ? V[(size_t)j * ld + order[r]] * Dv[j] * rsqrt(used[r]) : 0.0;
}
// T[i][k] = V[i][sel[k]] * wgt[k]   and   U[i][k] = V[i][sel[k]]
__global__ void wl_gather_kernel(const double* V, int n, int ld, const int* sel, const double* wgt, int ns,
                                 double* T, double* U, int ldt) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * ns) return;
    const int i = (int)(e / ns), k = (int)(e % ns);
    const double v = V[(size_t)i * ld + sel[k]];
    T[(size_t)i * ldt + k] = v * wgt[k];
    U[(size_t)i * ldt + k] = v;
}
// out = base (or 0) + G / (D_i D_j)
__global__ void wl_combine_kernel(const double* base, const double* G, int n, int ldg, const double* Dv, double* out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    const int i = (int)(e / n), j = (int)(e % n);
    out[e] = (base ? base[e] : 0.0) + G[(size_t)i * ldg + j] / (Dv[i] * Dv[j]);
}

#define WL_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return cuda_fail(nullptr, e_, "whiten_large"); } } while (0)

int whiten_large(int device, int n, const double* d_cov, double svdcut, double* d_w, double* d_cov_out,
                 int* d_nout, int* d_nmod, double* d_logdet, cudaStream_t s) {
    (void)device;
    const int ld = (n + 1) & ~1;                     // even leading dimension (aligned GEMM path)
    const int nb = (n + BJ_B - 1) / BJ_B, nbe = (nb + 1) & ~1, npairs = nbe / 2;
    double *A = nullptr, *V = nullptr, *Q = nullptr, *Dv = nullptr, *val = nullptr, *offsq = nullptr;
    int* d_rot = nullptr;
    double *T = nullptr, *U = nullptr, *G = nullptr, *d_used = nullptr, *d_wgt = nullptr;
    int *d_order = nullptr, *d_sel = nullptr;
    auto cleanup = [&]() {
        cudaFree(A); cudaFree(V); cudaFree(Q); cudaFree(Dv); cudaFree(val); cudaFree(offsq); cudaFree(d_rot);
        cudaFree(T); cudaFree(U); cudaFree(G); cudaFree(d_used); cudaFree(d_wgt); cudaFree(d_order); cudaFree(d_sel);
    };
    WL_TRY(cudaMalloc((void**)&A, (size_t)n * ld * sizeof(double)));
    WL_TRY(cudaMalloc((void**)&V, (size_t)n * ld * sizeof(double)));
    WL_TRY(cudaMalloc((void**)&Q, (size_t)npairs * BJ_M * BJ_M * sizeof(double)));
    WL_TRY(cudaMalloc((void**)&Dv, n * sizeof(double)));
    WL_TRY(cudaMalloc((void**)&val, n * sizeof(double)));
    WL_TRY(cudaMalloc((void**)&offsq, sizeof(double)));
    WL_TRY(cudaMalloc((void**)&d_rot, sizeof(int)));
    const size_t sm_diag = 2 * (size_t)BJ_M * BJ_LD * sizeof(double);
    const size_t sm_slab = ((size_t)BJ_M * BJ_PA + (size_t)BJ_M * BJ_PB) * sizeof(double);
    WL_TRY(cudaFuncSetAttribute(bj_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_diag));
    WL_TRY(cudaFuncSetAttribute(bj_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_slab));
    WL_TRY(cudaFuncSetAttribute(bj_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_slab));
    const int tpb = 256;
    const unsigned gnn = (unsigned)(((size_t)n * n + tpb - 1) / tpb);
    wl_init_kernel<<<gnn, tpb, 0, s>>>(d_cov, n, A, V, ld, Dv);
    WL_TRY(cudaGetLastError());

    BJArgs a;
    a.A = A; a.V = V; a.n = n; a.ld = ld; a.nb = nb; a.nbe = nbe; a.Q = Q; a.offsq = offsq; a.rotated = d_rot; a.round = 0;
    const dim3 gslab((n + 63) / 64, npairs);
    // ||corr||_F^2 <= n^2 (unit diagonal, |corr_ij| <= 1): convergence relative to n (trace)
    const double tol = (double)n * 1e-30 * n;         // off^2 <= (1e-15)^2 * n * trace-ish
    int sweeps = 0;
    for (; sweeps < 30; ++sweeps) {
        WL_TRY(cudaMemsetAsync(offsq, 0, sizeof(double), s));
        WL_TRY(cudaMemsetAsync(d_rot, 0, sizeof(int), s));
        for (int r = 0; r < nbe - 1; ++r) {
            a.round = r;
            bj_diag_kernel<<<npairs, BJ_THREADS, sm_diag, s>>>(a);
            bj_cols_kernel<<<gslab, BJ_THREADS, sm_slab, s>>>(a, A, n);
            bj_cols_kernel<<<gslab, BJ_THREADS, sm_slab, s>>>(a, V, n);
            bj_rows_kernel<<<gslab, BJ_THREADS, sm_slab, s>>>(a);
        }
        WL_TRY(cudaGetLastError());
        double h_off = 0.0;
        int h_rot = 0;
        WL_TRY(cudaMemcpyAsync(&h_off, offsq, sizeof(double), cudaMemcpyDeviceToHost, s));
        WL_TRY(cudaMemcpyAsync(&h_rot, d_rot, sizeof(int), cudaMemcpyDeviceToHost, s));
        WL_TRY(cudaStreamSynchronize(s));
        if (getenv("B200LM_VERBOSE")) fprintf(stderr, "whiten_large: sweep %d off^2 %.3e rotated pairs %d\n", sweeps, h_off, h_rot);
        if (h_rot == 0 || !(h_off > tol)) { ++sweeps; break; }
    }
    // ---- spectrum on the host (n doubles), svdcut bookkeeping --------------------------------
    wl_diag_kernel<<<(n + tpb - 1) / tpb, tpb, 0, s>>>(A, n, ld, val);
    std::vector<double> h_val(n);
    WL_TRY(cudaMemcpyAsync(h_val.data(), val, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    WL_TRY(cudaStreamSynchronize(s));
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return h_val[x] > h_val[y]; });
    const double vmax = h_val[order[0]];
    const double valmin = fabs(svdcut) * vmax;
    const bool cut = svdcut != 0.0, drop = cut && svdcut < 0.0;
    int nmod = 0;
    if (cut) for (int i = 0; i < n; ++i) nmod += h_val[i] < valmin ? 1 : 0;
    const int nkeep = drop ? n - nmod : n;
    std::vector<double> used(n, 1.0);
    double logdet = 0.0;
    for (int r = 0; r < nkeep; ++r) {
        double v = h_val[order[r]];
        if (cut && !drop && v < valmin) v = valmin;
        v = fabs(v);
        used[r] = v;
        logdet += log(v);
    }
    std::vector<double> h_D(n);
    WL_TRY(cudaMemcpy(h_D.data(), Dv, n * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) logdet -= 2.0 * log(h_D[i]);
    WL_TRY(cudaMalloc((void**)&d_order, n * sizeof(int)));
    WL_TRY(cudaMalloc((void**)&d_used, n * sizeof(double)));
    WL_TRY(cudaMemcpyAsync(d_order, order.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
    WL_TRY(cudaMemcpyAsync(d_used, used.data(), n * sizeof(double), cudaMemcpyHostToDevice, s));
    wl_w_kernel<<<gnn, tpb, 0, s>>>(V, n, ld, d_order, d_used, Dv, nkeep, d_w);
    // ---- corrected covariance ------------------------------------------------------------------
    std::vector<int> sel; std::vector<double> wgt;
    if (cut && nmod > 0) {
        if (!drop) { for (int r = n - nmod; r < n; ++r) { sel.push_back(order[r]); wgt.push_back(valmin - h_val[order[r]]); } }
        else       { for (int r = 0; r < nkeep; ++r)    { sel.push_back(order[r]); wgt.push_back(h_val[order[r]]); } }
    }
    if (!sel.empty()) {
        const int ns = (int)sel.size(), ldt = (ns + 1) & ~1;
        WL_TRY(cudaMalloc((void**)&d_sel, ns * sizeof(int)));
        WL_TRY(cudaMalloc((void**)&d_wgt, ns * sizeof(double)));
        WL_TRY(cudaMalloc((void**)&T, (size_t)n * ldt * sizeof(double)));
        WL_TRY(cudaMalloc((void**)&U, (size_t)n * ldt * sizeof(double)));
        WL_TRY(cudaMalloc((void**)&G, (size_t)n * ld * sizeof(double)));
        WL_TRY(cudaMemcpyAsync(d_sel, sel.data(), ns * sizeof(int), cudaMemcpyHostToDevice, s));
        WL_TRY(cudaMemcpyAsync(d_wgt, wgt.data(), ns * sizeof(double), cudaMemcpyHostToDevice, s));
        const unsigned gns = (unsigned)(((size_t)n * ns + tpb - 1) / tpb);
        wl_gather_kernel<<<gns, tpb, 0, s>>>(V, n, ld, d_sel, d_wgt, ns, T, U, ldt);
        // G = T . U^T  (n x n, K = ns) on the DMMA GEMM
        WL_TRY(dgemm(false, true, 1, n, n, ns, 1.0, T, 0, ldt, U, 0, ldt, 0.0, G, 0, ld, s));
        wl_combine_kernel<<<gnn, tpb, 0, s>>>(drop ? nullptr : d_cov, G, n, ld, Dv, d_cov_out);
    } else {
        WL_TRY(cudaMemcpyAsync(d_cov_out, d_cov, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
    WL_TRY(cudaMemcpyAsync(d_nout, &nkeep, sizeof(int), cudaMemcpyHostToDevice, s));
    WL_TRY(cudaMemcpyAsync(d_nmod, &nmod, sizeof(int), cudaMemcpyHostToDevice, s));
    WL_TRY(cudaMemcpyAsync(d_logdet, &logdet, sizeof(double), cudaMemcpyHostToDevice, s));
    WL_TRY(cudaStreamSynchronize(s));
    cleanup();
    return B200LM_OK;
}

}  // namespace b200lm
