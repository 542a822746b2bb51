// Wave variant of the batched LM kernel: a CTA keeps S = 32 fits in flight and takes all of them through
// one trial point per pass, in CTA-wide phases -- built for THROUGHPUT on large batches, where the
// warp-per-fit kernels (lm_kernel.cuh, lm_team.cuh) are bound by the dependency latency of one warp's
// 16 x 16 factorisations and by their code size (instruction fetch).
//
//   solve phase   EIGHT lanes per fit (warp w owns slots 4w .. 4w+3; shift index q = 0/1 x four sub-lanes).  The scalar
//                 trust-region bookkeeping of the reference's solver (scipy trf behind src/lsqfit/_scipy.py:156-161:
//                 radius update, acceptance, ftol/xtol/gtol tests) is replicated in the eight lanes; the vector work of a
//                 fit (np <= 16 entries: two per lane -- scale update, step, trial point, A.step) is shared, with
//                 shuffles inside the group.  The two-shift Newton/secant iteration on the secular equation is that of
//                 solve_tr_dual (lm_kernel.cuh), ONE round per pass: a fit whose shift has not converged sits the
//                 evaluation out and continues in the next pass, so nobody waits for the slowest fit.
//   factorisation the 2 S factorisations of a round run on TWO FULL WARPS, one thread per (fit, shift): a plain scalar
//                 L D L^T of d (J^T J) d + alpha I on a private column of shared memory (M[e][thread]: conflict free,
//                 compile-time offsets, no shuffles; the unrolled code keeps the matrix in registers between column
//                 steps), so a warp does 32 factorisations in the instructions the lane-per-row scheme needs for two.
//   eval phase    all 8 warps, chunks of 8 fits (warp w <-> fit w of the chunk):
//       E1  rows of [G | delta] of the correlated block, 17 columns per fit side by side in Z (64 x 136)
//       E2  Y = W . Z on the FP64 tensor path: ONE GEMM for the chunk -- the whitening matrix is shared by
//           the batch, so its fragments are loaded once per 8 fits and no tile is padded per fit
//           (reference: chiv, src/lsqfit/_utilities.pyx:65-94, here for 8 fits at once)
//       E3  [J | r]^T [J | r] of each fit from its 17 columns of Y (DMMA), 1x1 prior rows added analytically
//   Fits leave and enter slot by slot (device work queue), so the 32 slots stay full until the queue is empty.
// The kernel ends with x, chi2, nit, status; covariance, log det, optional f / J (and polish) are produced by
// the one-warp kernel compiled without a trust-region loop (fit_kernel<F, 2>: one evaluation + one factorisation per fit).
#pragma once
#include "lm_kernel.cuh"

namespace b200lm {

template <class F, int S_ = 8, int CF_ = 4>
struct WaveLayout {
    static constexpr int NP = F::NP;
    static constexpr int NC = NP + 1;                 // columns of one fit in a chunk: [G | delta]
    static constexpr int NPP = NP * (NP + 1) / 2;     // packed lower triangle (row-major: e(i,j) = i(i+1)/2 + j)
    static constexpr int S = S_;                      // fits in flight per CTA (<= 32)
    static constexpr int S1 = S + 1;                  // slot stride of the per-fit arrays (odd: scatter stores conflict free)
    static constexpr int ST = 2 * S;                  // solver threads (two per fit)
    static constexpr int CF = CF_;                    // fits per chunk = warps per CTA (divides 8)
    static constexpr int MTW = 8 / CF;                // 8-row output tiles per warp in the chunk GEMM
    static constexpr int THREADS = 32 * CF;
    static constexpr int NTF = (NC + 7) / 8;          // 8-column tiles covering one fit's columns
    static constexpr int NTRI = NTF * (NTF + 1) / 2;
    static constexpr int ZC = CF * NC;
    static constexpr int NTZ = (ZC + 7) / 8;          // n-tiles of the chunk GEMM
    static constexpr int ZW = (8 * NTZ > (CF - 1) * NC + 8 * NTF) ? 8 * NTZ : (CF - 1) * NC + 8 * NTF;   // widest column touched
    static constexpr int LDZ = ((ZW + 3) / 8) * 8 + 4;      // == 4 (mod 8): DMMA fragment loads hit 16 distinct banks per half warp
    static constexpr int KMAX = 64;                   // block size limit (inputs and outputs)
    static constexpr int ZROWS = KMAX + 1;
    static constexpr int PWLD = (NP + 1) & ~1;
    // chunk buffer / factorisation columns (+ the exchange area between the solver lanes and the factorisation warps)
    static constexpr int ZM = (ZROWS * LDZ > (NPP + 6) * ST) ? ZROWS * LDZ : (NPP + 6) * ST;
    // doubles of shared memory after the staged weights
    static constexpr int O_A = 0;                                 // [2][NPP][S1]
    static constexpr int O_G = O_A + 2 * NPP * S1;                // [2][NP][S1]
    static constexpr int O_PV = O_G + 2 * NP * S1;                // current point
    static constexpr int O_PN = O_PV + NP * S1;                   // trial point
    static constexpr int O_SI = O_PN + NP * S1;                   // column scale  (|J_j|, running max)
    static constexpr int O_DS = O_SI + NP * S1;                   // 1 / scale
    static constexpr int O_GN = O_DS + NP * S1;                   // cached Gauss-Newton step
    static constexpr int O_BP = O_GN + NP * S1;                   // best step of the secular iteration
    static constexpr int O_PT = O_BP + NP * S1;                   // [NP][ST] step of the thread's last factorisation
    static constexpr int O_PW = O_PT + NP * ST;                   // [CF][PWLD] contiguous parameter vector per warp
    static constexpr int O_COST = O_PW + CF * PWLD;               // [S]
    static constexpr int O_ZM = (O_COST + S + 1) & ~1;
    static constexpr int O_INT = O_ZM + ZM;                       // ints: req[S], tgt[S], fit[S], blk_idx[KMAX]
    static constexpr int TOTAL = O_INT + (3 * S + KMAX + 1) / 2 + 1;
    // wt_total > 0: the block weights are staged in shared memory; 0: read through L1 (several CTAs per SM share them)
    __host__ __device__ static size_t bytes(int wt_total) { return ((size_t)((wt_total + 1) & ~1) + TOTAL) * sizeof(double); }
};

__device__ __forceinline__ constexpr int wtri(int i, int j) { return i * (i + 1) / 2 + j; }      // i >= j

// One factorisation + solves by ONE thread:  M = d_i A_ij d_j + alpha delta_ij = L D L^T (in the thread's column of
// shared memory),  p = -M^-1 (d g)  -> PT,  |p|,  p^T M^-1 p.  ok = numerically positive definite (same pivot test as
// ldl_solve in lm_kernel.cuh).  All offsets are compile-time constants: no address arithmetic, no shuffles.
template <int NP, int S1, int ST>
__device__ __noinline__ void wave_factor_solve(const double* A_s, const double* G_s, const double* DS_s,
                                               double* M_t, double* PT_t, double alpha,
                                               double* pn_out, double* w2_out, int* ok_out) {
    __builtin_assume(__isShared(A_s));
    __builtin_assume(__isShared(G_s));
    __builtin_assume(__isShared(DS_s));
    __builtin_assume(__isShared(M_t));
    __builtin_assume(__isShared(PT_t));
    // register diet: only ONE NP-vector is live in any phase (this function's registers come on top of the
    // caller's per-fit state; several CTAs per SM need the kernel under ~128 registers)
    {
        double d[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) d[j] = DS_s[j * S1];
#pragma unroll
        for (int i = 0; i < NP; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const double a = A_s[wtri(i, j) * S1] * d[i] * d[j];
                M_t[wtri(i, j) * ST] = (i == j) ? a + alpha : a;
            }
#pragma unroll
        for (int i = 0; i < NP; ++i) PT_t[i * ST] = -(d[i] * G_s[i * S1]);       // right-hand side -d g
    }
    bool ok = true;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const double piv = M_t[wtri(j, j) * ST];
        // the diagonal of the unfactored matrix is >= alpha and >= the pivot's own scale; test as ldl_solve does
        const double dj = DS_s[j * S1];
        const double mdiag = fma(A_s[wtri(j, j) * S1] * dj, dj, alpha);
        ok = ok && (piv > 8.0 * NP * 2.220446049250313e-16 * mdiag) && (piv < 1e300);
        const double r = fast_rcp(piv);
        M_t[wtri(j, j) * ST] = r;                           // the diagonal keeps 1 / D_j
        double c[NP];
#pragma unroll
        for (int i = j + 1; i < NP; ++i) c[i] = M_t[wtri(i, j) * ST];
#pragma unroll
        for (int i = j + 1; i < NP; ++i) M_t[wtri(i, j) * ST] = c[i] * r;
#pragma unroll
        for (int k = j + 1; k < NP; ++k) {
            const double lk = c[k] * r;
#pragma unroll
            for (int i = k; i < NP; ++i) M_t[wtri(i, k) * ST] = fma(-lk, c[i], M_t[wtri(i, k) * ST]);
        }
    }
    double y[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) y[i] = PT_t[i * ST];
#pragma unroll
    for (int i = 1; i < NP; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) y[i] = fma(-M_t[wtri(i, j) * ST], y[j], y[i]);
#pragma unroll
    for (int i = 0; i < NP; ++i) y[i] *= M_t[wtri(i, i) * ST];
#pragma unroll
    for (int i = NP - 2; i >= 0; --i)
#pragma unroll
        for (int j = i + 1; j < NP; ++j) y[i] = fma(-M_t[wtri(j, i) * ST], y[j], y[i]);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NP; ++i) { s = fma(y[i], y[i], s); PT_t[i * ST] = y[i]; }
#pragma unroll
    for (int i = 1; i < NP; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) y[i] = fma(-M_t[wtri(i, j) * ST], y[j], y[i]);
    double w2 = 0.0;
#pragma unroll
    for (int i = 0; i < NP; ++i) w2 = fma(y[i] * y[i], M_t[wtri(i, i) * ST], w2);
    *pn_out = ok ? sqrt(s) : 0.0;
    *w2_out = ok ? w2 : 0.0;
    *ok_out = ok ? 1 : 0;
}

struct WCand {
    bool ok;
    double a, pn, w2;
};

// shared-memory map of a wave CTA (offsets: WaveLayout)
struct WaveSmem {
    double *Wsm, *Abuf, *Gbuf, *PV, *PN, *SI, *DS, *GN, *BP, *PT, *PW, *COST, *ZM;
    int *s_req, *s_tgt, *s_fit, *s_bidx;
};
struct WaveTicks {
    long long e1, e2, e3;
};

template <class WL>
__device__ __forceinline__ WaveSmem wave_smem(double* smem, int wtot) {
    WaveSmem m;
    double* base = smem + wtot;
    m.Wsm = smem;
    m.Abuf = base + WL::O_A; m.Gbuf = base + WL::O_G; m.PV = base + WL::O_PV; m.PN = base + WL::O_PN;
    m.SI = base + WL::O_SI; m.DS = base + WL::O_DS; m.GN = base + WL::O_GN; m.BP = base + WL::O_BP;
    m.PT = base + WL::O_PT; m.PW = base + WL::O_PW; m.COST = base + WL::O_COST; m.ZM = base + WL::O_ZM;
    m.s_req = reinterpret_cast<int*>(base + WL::O_INT);
    m.s_tgt = m.s_req + WL::S; m.s_fit = m.s_tgt + WL::S; m.s_bidx = m.s_fit + WL::S;
    return m;
}

// Evaluation phase of a pass (all warps of the CTA): every slot with a request (s_req: 1 = at PV, 2 = at PN) gets
// J^T J (packed lower), J^T r and the cost of its point into buffer s_tgt; chunks of CF fits, warp w <-> fit w of the chunk.
template <class F, class WL, bool WSMEM>
__device__ __forceinline__ void wave_eval_phase(const FitParams& P, const WaveSmem& m, const BlockDesc& bd, int nk4, int mtiles,
                                                WaveTicks& tk) {
    constexpr int NP = WL::NP, NC = WL::NC, NPP = WL::NPP, S = WL::S, S1 = WL::S1, CF = WL::CF;
    constexpr int NTF = WL::NTF, NTRI = WL::NTRI, NTZ = WL::NTZ, LDZ = WL::LDZ, MTW = WL::MTW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* const Wsm = m.Wsm; double* const Abuf = m.Abuf; double* const Gbuf = m.Gbuf;
    double* const PV = m.PV; double* const PN = m.PN; double* const PW = m.PW; double* const COST = m.COST; double* const ZM = m.ZM;
    int* const s_req = m.s_req; int* const s_tgt = m.s_tgt; int* const s_fit = m.s_fit; int* const s_bidx = m.s_bidx;
    long long& tk_e1 = tk.e1; long long& tk_e2 = tk.e2; long long& tk_e3 = tk.e3;
    (void)tid; (void)NPP; (void)S;
    // ================= evaluation of all requested points, 8 fits per chunk =================
    const unsigned reqmask = __ballot_sync(B200LM_FULL, lane < S && s_req[lane] != 0);
    const int nact = __popc(reqmask);
    double* const Z = ZM;
    for (int c0 = 0; c0 < nact; c0 += CF) {
        const int nfc = min(CF, nact - c0);               // fits in this chunk
        const bool valid = warp < nfc;
        const int slot = valid ? __fns(reqmask, 0, c0 + warp + 1) : 0;
        const int b = valid ? s_fit[slot] : 0;
        const int req = valid ? s_req[slot] : 0;
        const int tgt = valid ? s_tgt[slot] : 0;
        const double* mean = P.mean + (size_t)b * P.mean_stride;
        double* pw = PW + warp * WL::PWLD;
        // ---- E1: rows of [G | delta] ----
        const long long t_e1 = B200LM_CLOCK();
        if (valid) {
            if (lane < NP) pw[lane] = (req == 1 ? PV : PN)[lane * S1 + slot];
            __syncwarp();
            // the means of this fit's rows come from global memory: issue the loads before the model arithmetic
            double mpre[(WL::KMAX + 31) / 32];
#pragma unroll
            for (int u = 0; u < (WL::KMAX + 31) / 32; ++u) {
                const int r = lane + 32 * u;
                mpre[u] = r < bd.n_in ? mean[s_bidx[r]] : 0.0;
            }
            int u_row = 0;
            for (int r = lane; r < nk4; r += 32, ++u_row) {
                double* zr = Z + r * LDZ + NC * warp;
                if (r < bd.n_in) {
                    const int idx = s_bidx[r];
                    const double mean_r = u_row == 0 ? mpre[0] : mpre[(WL::KMAX + 31) / 32 - 1];
                    double dlt;
                    if (idx < P.ny) {
                        // functors that can split their terms evaluate both halves in one lane: the two
                        // unrolled halves are independent, so the latencies of their exponentials overlap
                        double f;
                        if constexpr (SplitOf<F>::value >= 2) {
                            const double f0 = F::value_grad_part(P.x + (size_t)idx * P.nx, idx, pw, 1.0, zr, 0);
                            const double f1 = F::value_grad_part(P.x + (size_t)idx * P.nx, idx, pw, 1.0, zr, 1);
                            f = f0 + f1;
                        } else {
                            f = F::value_grad(P.x + (size_t)idx * P.nx, idx, pw, 1.0, zr);
                        }
                        dlt = f - mean_r;
                    } else {
                        const int j0 = idx - P.ny;
#pragma unroll
                        for (int j = 0; j < NP; ++j) zr[j] = (j == j0) ? 1.0 : 0.0;
                        dlt = pw[j0] - mean_r;
                    }
                    zr[NP] = dlt;
                } else {
#pragma unroll
                    for (int j = 0; j < NC; ++j) zr[j] = 0.0;
                }
            }
        }
        __syncthreads();
        const long long t_e2 = B200LM_CLOCK();
        tk_e1 += t_e2 - t_e1;
        // ---- E2: Y = W . Z; warp w owns the row tiles w, w + CF, ... of the block's residuals ----
        double acc[MTW][NTZ][2];
        const int ntz = (nfc * NC + 7) >> 3;
        {
#pragma unroll
            for (int m = 0; m < MTW; ++m)
#pragma unroll
                for (int t = 0; t < NTZ; ++t) { acc[m][t][0] = 0.0; acc[m][t][1] = 0.0; }
            const double* wsrc = WSMEM ? Wsm : P.blk_wt;
            const double* wa = wsrc + bd.wt_off + (8 * warp + (lane >> 2)) * bd.ldw + (lane & 3);
            const double* zb = Z + (lane & 3) * LDZ + (lane >> 2);
            const int mstride = 8 * CF * bd.ldw;
#pragma unroll 2
            for (int s4 = 0; s4 < nk4; s4 += 4) {
                double af[MTW];
#pragma unroll
                for (int m = 0; m < MTW; ++m)
                    af[m] = (warp + CF * m < mtiles) ? (WSMEM ? wa[m * mstride + s4] : __ldg(wa + m * mstride + s4)) : 0.0;
#pragma unroll
                for (int t = 0; t < NTZ; ++t) {
                    if (t < ntz) {
                        const double bf = zb[s4 * LDZ + 8 * t];
#pragma unroll
                        for (int m = 0; m < MTW; ++m) dmma(acc[m][t][0], acc[m][t][1], af[m], bf);
                    }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < MTW; ++m) {
            if (warp + CF * m < mtiles) {
                double* yo = Z + (8 * (warp + CF * m) + (lane >> 2)) * LDZ + 2 * (lane & 3);
#pragma unroll
                for (int t = 0; t < NTZ; ++t)
                    if (t < ntz) *reinterpret_cast<double2*>(yo + 8 * t) = make_double2(acc[m][t][0], acc[m][t][1]);
            }
        }
        __syncthreads();
        const long long t_e3 = B200LM_CLOCK();
        tk_e2 += t_e3 - t_e2;
        // ---- E3: normal equations of fit `warp` of the chunk from its columns of Y ----
        if (valid) {
            double na[NTRI][2];
#pragma unroll
            for (int u = 0; u < NTRI; ++u) { na[u][0] = 0.0; na[u][1] = 0.0; }
            const double* yb = Z + (lane & 3) * LDZ + NC * warp + (lane >> 2);
            for (int s4 = 0; s4 < 8 * mtiles; s4 += 4) {
                double f[NTF];
#pragma unroll
                for (int t = 0; t < NTF; ++t) f[t] = yb[s4 * LDZ + 8 * t];
                int u = 0;
#pragma unroll
                for (int ta = 0; ta < NTF; ++ta)
#pragma unroll
                    for (int tb = ta; tb < NTF; ++tb) { dmma(na[u][0], na[u][1], f[ta], f[tb]); ++u; }
            }
            double* At = Abuf + tgt * NPP * S1 + slot;
            double* Gt = Gbuf + tgt * NP * S1 + slot;
            double csum = 0.0;
            int u = 0;
#pragma unroll
            for (int ta = 0; ta < NTF; ++ta)
#pragma unroll
                for (int tb = ta; tb < NTF; ++tb) {
                    const int i = 8 * ta + (lane >> 2);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = 8 * tb + 2 * (lane & 3) + e;
                        const double v = na[u][e];
                        if (i < NP && j < NP) {
                            if (j <= i) At[wtri(i, j) * S1] = v;
                            else if (ta != tb) At[wtri(j, i) * S1] = v;
                        } else if (j == NP && i < NP) {
                            Gt[i * S1] = v;
                        } else if (j == NP && i == NP) {
                            csum += v;
                        }
                    }
                    ++u;
                }
            __syncwarp();
            // 1x1 prior rows: J row = w e_j
            for (int i = lane; i < P.nd_pr; i += 32) {
                const int idx = P.dpr_idx[i];
                const int j = idx - P.ny;
                const double w = P.dpr_w[i];
                const double r = w * (pw[j] - mean[idx]);
                At[wtri(j, j) * S1] += w * w;
                Gt[j * S1] += w * r;
                csum = fma(r, r, csum);
            }
            csum = warp_sum(csum);
            if (lane == 0) COST[slot] = 0.5 * csum;
        }
        __syncthreads();
        tk_e3 += B200LM_CLOCK() - t_e3;
    }
}

// ---- reductions inside the eight lanes of a fit ----
__device__ __forceinline__ double sum8(unsigned gm, double v) {
    v += __shfl_xor_sync(gm, v, 1);
    v += __shfl_xor_sync(gm, v, 2);
    v += __shfl_xor_sync(gm, v, 4);
    return v;
}
__device__ __forceinline__ double max8(unsigned gm, double v) {
    v = fmax(v, __shfl_xor_sync(gm, v, 1));
    v = fmax(v, __shfl_xor_sync(gm, v, 2));
    v = fmax(v, __shfl_xor_sync(gm, v, 4));
    return v;
}

template <class F, int S_, int CF_, int MINB, bool WSMEM>
__global__ void __launch_bounds__(WaveLayout<F, S_, CF_>::THREADS, MINB) fit_wave_kernel(const __grid_constant__ FitParams P) {
    typedef WaveLayout<F, S_, CF_> WL;
    constexpr int NP = WL::NP, NPP = WL::NPP, S = WL::S, S1 = WL::S1, ST = WL::ST, CF = WL::CF;
    static_assert(NP <= 16 && S == 4 * CF, "two vector entries per lane; four slots per warp");
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wtot = WSMEM ? (P.wt_total + 1) & ~1 : 0;
    const WaveSmem wm = wave_smem<WL>(smem, wtot);
    double* const Abuf = wm.Abuf; double* const Gbuf = wm.Gbuf; double* const PV = wm.PV; double* const PN = wm.PN;
    double* const SI = wm.SI; double* const DS = wm.DS; double* const GN = wm.GN; double* const BP = wm.BP;
    double* const PT = wm.PT; double* const COST = wm.COST; double* const ZM = wm.ZM;
    static_assert(WL::ZM >= NPP * WL::ST + 5 * WL::ST + WL::S, "room for the exchange area behind the factorisation columns");
    WaveTicks wtk;
    wtk.e1 = wtk.e2 = wtk.e3 = 0;
    const BlockDesc bd = P.blk[0];
    if (WSMEM) for (int i = tid; i < P.wt_total; i += blockDim.x) wm.Wsm[i] = P.blk_wt[i];
    for (int i = tid; i < bd.n_in; i += blockDim.x) wm.s_bidx[i] = P.blk_idx[bd.idx_off + i];
    if (tid < S) { wm.s_req[tid] = 0; wm.s_tgt[tid] = 0; wm.s_fit[tid] = -1; }
    __syncthreads();
    const int nk4 = (bd.n_in + 3) & ~3;
    const int mtiles = (bd.n_out + 7) >> 3;

    // ---- lane roles ----
    const int s = 4 * warp + (lane >> 3);          // slot
    const int q = (lane >> 2) & 1, t = lane & 3;   // shift index, sub-lane
    const int t8 = lane & 7;                       // lane within the fit's group: vector entries t8 and t8 + 8
    const unsigned gm = 0xffu << (lane & ~7);
    const int mq = 2 * s + q;                      // column of the factorisation work space / of PT
    const bool lead = t8 == 0;
    const int j0 = t8, j1 = t8 + 8;
    const bool h0 = j0 < NP, h1 = j1 < NP;

    int fit = -1, st = 0, cur = 0, nfev = 0, nfac = 0;
    double cost = 0.0, Delta = 1.0, alpha = 0.0, predicted = 0.0, step2 = 0.0, sh2 = 0.0, xthr = 0.0;
    bool fresh = true, queue_empty = false;
    bool gn_valid = false, gn_full = false, lm_valid = false;
    double gn_pn = 0.0, gn_w2 = 0.0, lm_a = 0.0, lm_b = 0.0;
    int it = 0;
    double au = 0.0, al = 0.0;
    WCand best, second;
    best.ok = false; best.a = best.pn = best.w2 = 0.0; second = best;
    bool use_gn = false;
    unsigned long long tot_nfev = 0, tot_nfac = 0;
    long long tk_solve = 0, tk_pass = 0, tk_fact = 0, n_pass = 0, tk_d = 0, tk_sec = 0, tk_step = 0, tk_bar = 0;

    for (;;) {
        const long long t_p0 = B200LM_CLOCK();
        ++n_pass;
        int status = -2;
        // ---------------- decide ----------------
        if (st == 1) {
            const double* Ac = Abuf + cur * NPP * S1 + s;
            cost = COST[s];
            nfev = 1; nfac = 0;
            if (!isfinite(cost)) status = -1;
            else {
                double d2 = 0.0;
                if (h0) {
                    double si = 1.0;
                    if (P.scaler == 1) { si = sqrt(Ac[wtri(j0, j0) * S1]); if (si == 0.0) si = 1.0; }
                    SI[j0 * S1 + s] = si;
                    const double tt = PV[j0 * S1 + s] * si;
                    d2 = tt * tt;
                }
                if (h1) {
                    double si = 1.0;
                    if (P.scaler == 1) { si = sqrt(Ac[wtri(j1, j1) * S1]); if (si == 0.0) si = 1.0; }
                    SI[j1 * S1 + s] = si;
                    const double tt = PV[j1 * S1 + s] * si;
                    d2 = fma(tt, tt, d2);
                }
                Delta = sqrt(sum8(gm, d2));
                if (Delta == 0.0) Delta = 1.0;
                alpha = 0.0; lm_valid = false; fresh = true; st = 3; it = 0;
            }
        } else if (st == 2) {
            const double cost_new = COST[s];
            ++nfev;
            if (!isfinite(cost_new)) {
                Delta = 0.25 * sqrt(sh2);
                st = 3;
            } else {
                const double actual = cost - cost_new;
                bool lt25, gt25, gt75;
                if (predicted > 0.0) {
                    lt25 = actual < 0.25 * predicted; gt25 = actual > 0.25 * predicted; gt75 = actual > 0.75 * predicted;
                } else if (predicted == 0.0 && actual == 0.0) {
                    lt25 = false; gt25 = true; gt75 = true;
                } else {
                    lt25 = true; gt25 = false; gt75 = false;
                }
                const bool ft = actual < P.ftol * cost && gt25;
                const bool xt = step2 < xthr * xthr;
                int term = -2;
                if (ft && xt) term = 4; else if (ft) term = 2; else if (xt) term = 3;
                if (term == -2) {
                    if (lt25) {
                        const double Dn = 0.25 * sqrt(sh2);
                        alpha *= Delta / Dn;
                        Delta = Dn;
                    } else if (gt75 && sh2 > (0.95 * 0.95) * Delta * Delta) {
                        alpha *= 0.5;
                        Delta = 2.0 * Delta;
                    }
                }
                if (actual > 0.0) {
                    cur ^= 1;
                    cost = cost_new;
                    const double* An = Abuf + cur * NPP * S1 + s;
                    if (h0) {
                        PV[j0 * S1 + s] = PN[j0 * S1 + s];
                        if (P.scaler == 1) SI[j0 * S1 + s] = fmax(SI[j0 * S1 + s], sqrt(An[wtri(j0, j0) * S1]));
                    }
                    if (h1) {
                        PV[j1 * S1 + s] = PN[j1 * S1 + s];
                        if (P.scaler == 1) SI[j1 * S1 + s] = fmax(SI[j1 * S1 + s], sqrt(An[wtri(j1, j1) * S1]));
                    }
                    fresh = true;
                }
                if (term != -2) {
                    const double* Gn = Gbuf + cur * NP * S1 + s;
                    double gf = 0.0;
                    if (h0) gf = fabs(Gn[j0 * S1]);
                    if (h1) gf = fmax(gf, fabs(Gn[j1 * S1]));
                    gf = max8(gm, gf);
                    status = gf < P.gtol ? 1 : term;
                } else {
                    st = 3;
                }
            }
        }
        __syncwarp();
        const long long t_d1 = B200LM_CLOCK();
        tk_d += t_d1 - t_p0;
        // ---------------- head of an outer iteration ----------------
        if (st == 3 && status == -2 && it == 0) {
            const double* G = Gbuf + cur * NP * S1 + s;
            if (!fresh && nfev >= P.maxit) fresh = true;
            if (fresh) {
                double xx = 0.0, gmax = 0.0;
                if (h0) {
                    const double pj = PV[j0 * S1 + s];
                    xx = pj * pj; gmax = fabs(G[j0 * S1]);
                    DS[j0 * S1 + s] = 1.0 / SI[j0 * S1 + s];
                }
                if (h1) {
                    const double pj = PV[j1 * S1 + s];
                    xx = fma(pj, pj, xx); gmax = fmax(gmax, fabs(G[j1 * S1]));
                    DS[j1 * S1 + s] = 1.0 / SI[j1 * S1 + s];
                }
                xx = sum8(gm, xx); gmax = max8(gm, gmax);
                if (gmax < P.gtol) status = 1;
                else if (nfev >= P.maxit) status = 0;
                xthr = P.xtol * (P.xtol + sqrt(xx));
                gn_valid = false;
                fresh = false;
            }
        }
        __syncwarp();
        // ---------------- one round of the secular iteration ----------------
        const bool solving = (st == 3 && status == -2);
        const double* A = Abuf + cur * NPP * S1 + s;
        const double* G = Gbuf + cur * NP * S1 + s;
        const double* D = DS + s;
        if (solving && it == 0) {
            double gg = 0.0;
            if (h0) { const double x = D[j0 * S1] * G[j0 * S1]; gg = x * x; }
            if (h1) { const double x = D[j1 * S1] * G[j1 * S1]; gg = fma(x, x, gg); }
            gg = sum8(gm, gg);
            au = sqrt(gg) / Delta; al = 0.0;
            best.ok = false; second.ok = false; use_gn = false;
        }
        bool done = !solving;
        double a_lo = 0.0, a_hi = 0.0;
        int lo_kind = 0, hi_kind = 0;             // 0 candidate, 1 Gauss-Newton (shift 0), 2 ignore
        if (!done) {
            const double invD = fast_rcp(Delta);
            const double pr = (invD - lm_a) * fast_rcp(lm_b);
            const double x2p = (lm_valid && pr > 0.0 && pr < 1e300) ? pr : -1.0;
            const bool gnok = gn_valid && gn_full;
            const double gnl = gnok ? (gn_pn - Delta) * gn_pn * fast_rcp(gn_w2) : 0.0;
            if (it == 0) {
                const bool caseA = !gn_valid && !(alpha > 0.0);
                const bool warm = alpha > 0.0;
                const double x1 = warm ? alpha : (x2p > 0.0 ? x2p : fmax(gnl, 0.001 * au));
                double x2 = warm ? x2p : -1.0;
                x2 = (x2 < 0.0 || fabs(x2 - x1) < 1e-3 * x1) ? (gn_valid ? 1.5 * x1 : 0.0) : x2;
                a_lo = caseA ? 0.0 : x1;
                lo_kind = caseA ? 1 : 0;
                a_hi = caseA ? (x2p > 0.0 ? x2p : 0.0) : x2;
                hi_kind = caseA ? (x2p > 0.0 ? 0 : 2) : (x2 == 0.0 ? 1 : 0);
            } else {
                al = gnok ? fmax(al, gnl) : al;
                au = (best.ok && best.pn < Delta) ? fmin(au, best.a) : au;
                al = (best.ok && !(best.pn < Delta)) ? fmax(al, best.a) : al;
                au = (second.ok && second.pn < Delta) ? fmin(au, second.a) : au;
                al = (second.ok && !(second.pn < Delta)) ? fmax(al, second.a) : al;
                done = best.ok && fabs(best.pn - Delta) < 0.1 * Delta && (gn_valid || best.pn >= Delta);
                const double phi = best.pn - Delta;
                const double ratio = -phi * best.pn * fast_rcp(best.w2);
                double an = best.ok ? best.a - (phi + Delta) * ratio * invD : (alpha > 0.0 ? 2.0 * alpha : 0.0);
                an = (an > al && an < au) ? an : fmax(0.001 * au, sqrt(al * au));
                const double y0 = fast_rcp(best.pn), y1 = fast_rcp(second.pn);
                const bool sec = best.ok && second.ok && second.a != best.a && y1 != y0;
                double as = sec ? best.a + (invD - y0) * (second.a - best.a) * fast_rcp(y1 - y0) : -1.0;
                const double alt = (best.ok && best.pn < Delta) ? 0.5 * (an + al) : fmin(1.5 * an, 0.5 * (an + au));
                as = (!(as > al && as < au) || fabs(as - an) < 1e-3 * an) ? alt : as;
                as = as > 0.0 ? as : 1.5 * an;
                as = (!gn_valid && best.ok && best.pn < Delta) ? 0.0 : as;
                a_lo = an; a_hi = as; hi_kind = (as == 0.0) ? 1 : 0;
            }
        }
        // The factorisations themselves run on TWO FULL WARPS (one thread per (fit, shift) column, 32 active lanes per
        // instruction): the fully unrolled scalar code keeps the matrix in registers between column steps, and packed
        // lanes halve its shared-memory wavefronts.  Measured alternatives: four sub-lanes sharing a column through shared
        // memory 43 k cycles per round, one lane per column on all eight warps (8 active lanes) 31 k, packed 21 k.
        double* const XCH = ZM + NPP * ST;                  // free tail of the work space: alpha[ST] | res[ST][3] | flags
        int* const XI = reinterpret_cast<int*>(XCH + 4 * ST);
        if (t == 0) {
            XCH[mq] = q ? a_hi : a_lo;
            XI[mq] = done ? 0 : 1;
            if (q == 0) XI[ST + s] = cur;
        }
        __syncthreads();
        const long long t_f0 = B200LM_CLOCK();
        if (tid < ST) {
            double pn_f = 0.0, w2_f = 0.0;
            int ok_f = 0;
            if (XI[tid]) {
                const int sf = tid >> 1, cf = XI[ST + sf];
                wave_factor_solve<NP, S1, ST>(Abuf + cf * NPP * S1 + sf, Gbuf + cf * NP * S1 + sf, DS + sf, ZM + tid, PT + tid,
                                              XCH[tid], &pn_f, &w2_f, &ok_f);
            }
            XCH[ST + 3 * tid] = pn_f; XCH[ST + 3 * tid + 1] = w2_f; XCH[ST + 3 * tid + 2] = (double)ok_f;
        }
        __syncthreads();
        tk_fact += B200LM_CLOCK() - t_f0;
        {
            double pn_m = XCH[ST + 3 * mq], w2_m = XCH[ST + 3 * mq + 1];
            int ok_m = (int)XCH[ST + 3 * mq + 2];
            const double pn_o = __shfl_xor_sync(B200LM_FULL, pn_m, 4);
            const double w2_o = __shfl_xor_sync(B200LM_FULL, w2_m, 4);
            const int ok_o = __shfl_xor_sync(B200LM_FULL, ok_m, 4);
            if (!done) {
                nfac += (hi_kind == 2) ? 1 : 2;
                WCand cd[2];
                cd[0].a = a_lo; cd[1].a = a_hi;
                cd[0].ok = (q ? ok_o : ok_m) != 0; cd[0].pn = q ? pn_o : pn_m; cd[0].w2 = q ? w2_o : w2_m;
                cd[1].ok = (q ? ok_m : ok_o) != 0; cd[1].pn = q ? pn_m : pn_o; cd[1].w2 = q ? w2_m : w2_o;
                const int kind[2] = {lo_kind, hi_kind};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const WCand c2 = cd[h];
                    const double* src = PT + 2 * s + h;                      // the step vector of shift h
                    if (kind[h] == 1) {
                        gn_valid = true; gn_full = c2.ok; gn_pn = c2.pn; gn_w2 = c2.w2;
                        if (h0) GN[j0 * S1 + s] = src[j0 * ST];
                        if (h1) GN[j1 * S1 + s] = src[j1 * ST];
                    } else if (kind[h] == 0 && c2.ok && c2.a > 0.0) {
                        const double dd = fabs(c2.pn - Delta);
                        if (!best.ok || dd < fabs(best.pn - Delta)) {
                            second = best; best = c2;
                            if (h0) BP[j0 * S1 + s] = src[j0 * ST];
                            if (h1) BP[j1 * S1 + s] = src[j1 * ST];
                        } else if (!second.ok || dd < fabs(second.pn - Delta)) {
                            second = c2;
                        }
                    }
                }
                if (gn_valid && gn_full && gn_pn <= Delta) { alpha = 0.0; use_gn = true; done = true; }
                else if (it > 0 && !cd[0].ok && (hi_kind == 1 || !cd[1].ok)) {
                    al = fmax(al, fmax(a_lo, a_hi));
                    alpha = fmax(2.0 * fmax(a_lo, a_hi), 0.001 * au);
                    if (alpha > au) au = 2.0 * alpha;
                }
            }
        }
        __syncwarp();
        const long long t_s1 = B200LM_CLOCK();
        tk_sec += t_s1 - t_d1;
        const bool fin = solving && (use_gn || it >= 5 ||
                                     (best.ok && fabs(best.pn - Delta) < 0.1 * Delta && (gn_valid || best.pn >= Delta)));
        if (solving && !fin) ++it;
        if (fin) {
            it = 0;
            double scale = 1.0;
            const double* SH = BP + s;
            if (use_gn) {
                SH = GN + s;
            } else if (best.ok) {
                double bb;
                const double ib = fast_rcp(best.pn);
                const bool two = second.ok && second.a != best.a;
                bb = two ? (fast_rcp(second.pn) - ib) * fast_rcp(second.a - best.a) : best.w2 * ib * ib * ib;
                lm_valid = bb > 0.0; lm_b = bb; lm_a = ib - bb * best.a;
                const double phi = best.pn - Delta;
                const double ratio = -phi * best.pn * fast_rcp(best.w2);
                const double an = best.a - (phi + Delta) * ratio * fast_rcp(Delta);
                alpha = an > 0.0 ? an : best.a;
                if (best.pn > 0.0) scale = Delta * ib;
            } else {
                double gg = 0.0;
                if (h0) { const double x = D[j0 * S1] * G[j0 * S1]; BP[j0 * S1 + s] = -x; gg = x * x; }
                if (h1) { const double x = D[j1 * S1] * G[j1 * S1]; BP[j1 * S1 + s] = -x; gg = fma(x, x, gg); }
                gg = sum8(gm, gg);
                scale = gg > 0.0 ? Delta / sqrt(gg) : 0.0;
            }
            // step (two entries per lane), trial point; the step vector goes to the fit's first PT column
            double* STP = PT + 2 * s;
            double s2 = 0.0, h2 = 0.0;
            if (h0) {
                const double sh = SH[j0 * S1] * scale, sj = D[j0 * S1] * sh;
                h2 = sh * sh; s2 = sj * sj;
                STP[j0 * ST] = sj;
                PN[j0 * S1 + s] = PV[j0 * S1 + s] + sj;
            }
            if (h1) {
                const double sh = SH[j1 * S1] * scale, sj = D[j1 * S1] * sh;
                h2 = fma(sh, sh, h2); s2 = fma(sj, sj, s2);
                STP[j1 * ST] = sj;
                PN[j1 * S1 + s] = PV[j1 * S1 + s] + sj;
            }
            __syncwarp(gm);
            // predicted reduction = -(1/2 step^T A step + g^T step): rows j0, j1 of A.step per lane
            double pred = 0.0;
            if (h0) {
                double As0 = 0.0, As1 = 0.0;
#pragma unroll
                for (int j = 0; j < NP; j += 2) {
                    As0 = fma(A[(j0 >= j ? wtri(j0, j) : wtri(j, j0)) * S1], STP[j * ST], As0);
                    if (j + 1 < NP) As1 = fma(A[(j0 >= j + 1 ? wtri(j0, j + 1) : wtri(j + 1, j0)) * S1], STP[(j + 1) * ST], As1);
                }
                pred = -STP[j0 * ST] * (0.5 * (As0 + As1) + G[j0 * S1]);
            }
            if (h1) {
                double As0 = 0.0, As1 = 0.0;
#pragma unroll
                for (int j = 0; j < NP; j += 2) {
                    As0 = fma(A[(j1 >= j ? wtri(j1, j) : wtri(j, j1)) * S1], STP[j * ST], As0);
                    if (j + 1 < NP) As1 = fma(A[(j1 >= j + 1 ? wtri(j1, j + 1) : wtri(j + 1, j1)) * S1], STP[(j + 1) * ST], As1);
                }
                pred = fma(-STP[j1 * ST], 0.5 * (As0 + As1) + G[j1 * S1], pred);
            }
            predicted = sum8(gm, pred); step2 = sum8(gm, s2); sh2 = sum8(gm, h2);
            st = 2;
        }
        tk_step += B200LM_CLOCK() - t_s1;
        // ---------------- a finished fit leaves, the next one enters ----------------
        if (status != -2) {
            if (h0) P.x_out[(size_t)fit * NP + j0] = PV[j0 * S1 + s];      // (the eight lanes hold t8 = 0..7: every entry once)
            if (h1) P.x_out[(size_t)fit * NP + j1] = PV[j1 * S1 + s];
            if (P.wave_A && status != -1) {
                // J^T J at the solution (the buffer of the current point) for the finalisation pass
                const double* Ac = Abuf + cur * NPP * S1 + s;
                double* Ao = P.wave_A + (size_t)fit * NPP;
                for (int e = t8; e < NPP; e += 8) Ao[e] = Ac[e * S1];
            }
            if (lead) {
                P.chi2[fit] = 2.0 * cost;
                P.nit[fit] = nfev;
                P.status[fit] = status;
                tot_nfev += nfev; tot_nfac += nfac;
            }
            st = 0; fit = -1;
        }
        if (st == 0 && !queue_empty) {
            int b = 0;
            if (lead) {
                b = atomicAdd(P.counter, 1);
                if (P.order && b < P.B) b = P.order[b];
            }
            b = __shfl_sync(gm, b, lane & ~7);
            if (b < P.B) {
                fit = b; st = 1; cur = 0;
                const double* p0 = P.p0 + (size_t)b * P.p0_stride;
                if (h0) PV[j0 * S1 + s] = p0[j0];
                if (h1) PV[j1 * S1 + s] = p0[j1];
            } else {
                queue_empty = true;
            }
        }
        if (lead) {
            wm.s_req[s] = st == 1 ? 1 : (st == 2 ? 2 : 0);
            wm.s_tgt[s] = st == 1 ? cur : cur ^ 1;
            wm.s_fit[s] = fit;
        }
        const long long t_b0 = B200LM_CLOCK();
        tk_solve += t_b0 - t_p0;
        const int busy = __syncthreads_count(lead && st != 0);
        tk_bar += B200LM_CLOCK() - t_b0;
        if (busy == 0) break;
        wave_eval_phase<F, WL, WSMEM>(P, wm, bd, nk4, mtiles, wtk);
        tk_pass += B200LM_CLOCK() - t_p0;
    }
#ifdef B200LM_PHASE_TICKS
    if (tid == 0 && P.stats) {
        atomicAdd(&P.stats[3], (unsigned long long)(wtk.e1 + wtk.e2 + wtk.e3));
        atomicAdd(&P.stats[4], (unsigned long long)tk_solve);
        atomicAdd(&P.stats[5], (unsigned long long)tk_pass);
        atomicAdd(&P.stats[6], (unsigned long long)wtk.e1);
        atomicAdd(&P.stats[7], (unsigned long long)wtk.e2);
        atomicAdd(&P.stats[8], (unsigned long long)wtk.e3);
        atomicAdd(&P.stats[9], (unsigned long long)n_pass);
        atomicAdd(&P.stats[11], (unsigned long long)tk_fact);
        atomicAdd(&P.stats[12], (unsigned long long)tk_d);
        atomicAdd(&P.stats[13], (unsigned long long)tk_sec);
        atomicAdd(&P.stats[14], (unsigned long long)tk_step);
        atomicAdd(&P.stats[15], (unsigned long long)tk_bar);
    }
#endif
    if (lead && P.stats) {
        atomicAdd(&P.stats[0], tot_nfev);
        atomicAdd(&P.stats[1], tot_nfev);
        atomicAdd(&P.stats[2], tot_nfac);
    }
}

// shapes the wave kernel takes: exactly one correlated block of at most 64 points (any mix of data and prior
// entries), every other entry a 1x1 PRIOR row, scipy policy.  Everything else stays with the warp / team kernels.
// Two configurations: small CTAs (8 fits, 4 warps, weights through L1) of which several share an SM -- their
// solve and evaluation phases interleave --, or one large CTA per SM (32 fits, 8 warps, weights staged).
template <class F>
cudaError_t launch_fit_wave(FitParams P, int sm_count, size_t smem_budget, cudaStream_t stream) {
    // one CTA of 32 fits / 8 warps per SM, block weights staged.  Measured on C3 at B = 160 k (tools/wave_cfg.py):
    // smaller CTAs sharing an SM (16 fits / 4 warps x 2: 91 ms, 8 fits / 4 warps x 3: 140 ms against 85 ms) lose more in the
    // latency-bound evaluation phase, whose time doubles with half the warps, than their interleaving wins
    typedef WaveLayout<F, 32, 8> WL;
    const size_t smem = WL::bytes(P.wt_total);
    if (smem > smem_budget) return cudaErrorInvalidConfiguration;
    int grid = sm_count;
    const int want = (P.B + WL::S / 2 - 1) / (WL::S / 2);      // small batches: spread over the SMs, slots at least half full
    // Small models (np <= 8) leave room for TWO such CTAs per SM (shared memory: ~95 KB each at np = 6; registers capped at
    // 128 by the second instantiation): while one CTA is in its latency-bound solve phase the other evaluates.
    if constexpr (F::NP <= 8) {
        const char* env = getenv("B200LM_WAVE_CTAS");
        const int ctas = env ? atoi(env) : 2;
        if (ctas >= 2 && 2 * (smem + 1024) <= (size_t)227 * 1024) {      // (by shape only: never by batch size)
            auto kern2 = fit_wave_kernel<F, 32, 8, 2, true>;
            cudaError_t e2 = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e2 != cudaSuccess) return e2;
            grid = 2 * sm_count;
            if (grid > want) grid = want;
            if (grid < 1) grid = 1;
            kern2<<<grid, WL::THREADS, smem, stream>>>(P);
            return cudaGetLastError();
        }
    }
    auto kern = fit_wave_kernel<F, 32, 8, 1, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    kern<<<grid, WL::THREADS, smem, stream>>>(P);
    return cudaGetLastError();
}

template <class F>
size_t wave_bytes_of(int wt_total) { return WaveLayout<F, 32, 8>::bytes(wt_total); }

}  // namespace b200lm
