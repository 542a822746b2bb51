// Wave variant of the batched LM kernel: a CTA keeps S = 32 fits in flight and takes all of them through
// one trial point per pass, in CTA-wide phases -- built for THROUGHPUT on large batches, where the
// warp-per-fit kernels (lm_kernel.cuh, lm_team.cuh) are bound by the dependency latency of one warp's
// 16 x 16 factorisations and by their code size (instruction fetch).
//
//   solve phase   64 threads: a PAIR of threads per fit.  Each thread factors d (J^T J) d + alpha I for its own
//                 shift alpha with a plain scalar L D L^T on a private column of shared memory (M[e][thread]:
//                 conflict free, immediate offsets, no shuffles), so a warp does 32 factorisations in the
//                 instructions the lane-per-row scheme needs for two; the pair runs the same safeguarded
//                 two-shift Newton/secant iteration on the secular equation as solve_tr_dual (lm_kernel.cuh),
//                 and the scalar trust-region bookkeeping of the reference's solver (scipy trf behind
//                 src/lsqfit/_scipy.py:156-161: radius update, acceptance, ftol/xtol/gtol tests).
//   eval phase    all 8 warps, chunks of 8 fits (warp w <-> fit w of the chunk):
//       E1  rows of [G | delta] of the correlated block, 17 columns per fit side by side in Z (64 x 136)
//       E2  Y = W . Z on the FP64 tensor path: ONE GEMM for the chunk -- the whitening matrix is shared by
//           the batch, so its fragments are loaded once per 8 fits and no tile is padded per fit
//           (reference: chiv, src/lsqfit/_utilities.pyx:65-94, here for 8 fits at once)
//       E3  [J | r]^T [J | r] of each fit from its 17 columns of Y (DMMA), 1x1 prior rows added analytically
//   Fits leave and enter slot by slot (device work queue), so the 32 slots stay full until the queue is empty.
// The kernel ends with x, chi2, nit, status; covariance, log det, optional f / J (and polish) are produced by
// the one-warp kernel in `finalize_only` mode (one evaluation + one factorisation per fit).
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
    static constexpr int SWT = (2 * S + 31) & ~31;    // threads of the solver warps
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
    static constexpr int ZM = (ZROWS * LDZ > NPP * ST) ? ZROWS * LDZ : NPP * ST;      // chunk buffer / factorisation columns
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

template <class F, int S_, int CF_, int MINB, bool WSMEM>
__global__ void __launch_bounds__(WaveLayout<F, S_, CF_>::THREADS, MINB) fit_wave_kernel(const __grid_constant__ FitParams P) {
    typedef WaveLayout<F, S_, CF_> WL;
    constexpr int NP = WL::NP, NC = WL::NC, NPP = WL::NPP, S = WL::S, S1 = WL::S1, ST = WL::ST, CF = WL::CF;
    constexpr int NTF = WL::NTF, NTRI = WL::NTRI, NTZ = WL::NTZ, LDZ = WL::LDZ, MTW = WL::MTW, SWT = WL::SWT;
    static_assert(NP <= 16, "unrolled thread-level factorisation");
    static_assert(S <= 32 && (8 % CF) == 0 && SWT <= 32 * CF, "slots fit one ballot; warps divide the 8 row tiles");
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wtot = WSMEM ? (P.wt_total + 1) & ~1 : 0;
    double* const Wsm = smem;
    double* const base = smem + wtot;
    double* const Abuf = base + WL::O_A;
    double* const Gbuf = base + WL::O_G;
    double* const PV = base + WL::O_PV;
    double* const PN = base + WL::O_PN;
    double* const SI = base + WL::O_SI;
    double* const DS = base + WL::O_DS;
    double* const GN = base + WL::O_GN;
    double* const BP = base + WL::O_BP;
    double* const PT = base + WL::O_PT;
    double* const PW = base + WL::O_PW;
    double* const COST = base + WL::O_COST;
    double* const ZM = base + WL::O_ZM;
    int* const s_req = reinterpret_cast<int*>(base + WL::O_INT);
    int* const s_tgt = s_req + S;
    int* const s_fit = s_tgt + S;
    int* const s_bidx = s_fit + S;

    const BlockDesc bd = P.blk[0];
    if (WSMEM) for (int i = tid; i < P.wt_total; i += blockDim.x) Wsm[i] = P.blk_wt[i];
    for (int i = tid; i < bd.n_in; i += blockDim.x) s_bidx[i] = P.blk_idx[bd.idx_off + i];
    if (tid < S) { s_req[tid] = 0; s_tgt[tid] = 0; s_fit[tid] = -1; }
    __syncthreads();
    const int nk4 = (bd.n_in + 3) & ~3;
    const int mtiles = (bd.n_out + 7) >> 3;

    // ---- per-slot state of the solver threads (identical in both threads of a pair) ----
    const int s = tid >> 1, q = tid & 1;              // (threads beyond ST own no slot: their state stays empty)
    const unsigned pmask = 3u << (lane & ~1);
    int fit = -1, st = 0, cur = 0, nfev = 0, nfac = 0;       // st: 0 empty, 1 first evaluation pending, 2 trial pending, 3 ready to solve
    double cost = 0.0, Delta = 1.0, alpha = 0.0, predicted = 0.0, step2 = 0.0, sh2 = 0.0, xthr = 0.0;
    bool fresh = true, queue_empty = tid >= ST;
    bool gn_valid = false, gn_full = false, lm_valid = false;
    double gn_pn = 0.0, gn_w2 = 0.0, lm_a = 0.0, lm_b = 0.0;
    // state of the secular iteration of the current trial (ONE factorisation round per pass: a fit whose shift has
    // not converged yet sits the evaluation out and continues in the next pass -- nobody waits for the slowest pair)
    int it = 0;
    double au = 0.0, al = 0.0;
    WCand best, second;
    best.ok = false; best.a = best.pn = best.w2 = 0.0; second = best;
    bool use_gn = false;
    unsigned long long tot_nfev = 0, tot_nfac = 0;
    long long tk_solve = 0, tk_e1 = 0, tk_e2 = 0, tk_e3 = 0, tk_pass = 0, tk_fact = 0;      // thread 0 (B200LM_PHASE_TICKS)
    long long n_pass = 0, tk_d = 0, tk_sec = 0, tk_step = 0, tk_fin = 0;

    for (;;) {
        const long long t_p0 = B200LM_CLOCK();
        ++n_pass;
        if (tid < SWT) {
            double* const Ac = Abuf + cur * NPP * S1 + s;
            double* const Gc = Gbuf + cur * NP * S1 + s;
            int status = -2;
            // ---------------- decide: result of the evaluation requested in the previous pass ----------------
            if (st == 1) {
                cost = COST[s];
                nfev = 1; nfac = 0;
                if (!isfinite(cost)) status = -1;
                else {
                    double d2 = 0.0;
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        double si = 1.0;
                        if (P.scaler == 1) { si = sqrt(Ac[wtri(j, j) * S1]); if (si == 0.0) si = 1.0; }
                        SI[j * S1 + s] = si;
                        const double t = PV[j * S1 + s] * si;
                        d2 = fma(t, t, d2);
                    }
                    Delta = sqrt(d2);
                    if (Delta == 0.0) Delta = 1.0;
                    alpha = 0.0; lm_valid = false; fresh = true; st = 3;
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
#pragma unroll
                        for (int j = 0; j < NP; ++j) {
                            PV[j * S1 + s] = PN[j * S1 + s];
                            if (P.scaler == 1) SI[j * S1 + s] = fmax(SI[j * S1 + s], sqrt(An[wtri(j, j) * S1]));
                        }
                        fresh = true;
                    }
                    if (term != -2) {
                        const double* Gn = Gbuf + cur * NP * S1 + s;
                        double gf = 0.0;
#pragma unroll
                        for (int j = 0; j < NP; ++j) gf = fmax(gf, fabs(Gn[j * S1]));
                        status = gf < P.gtol ? 1 : term;
                    } else {
                        st = 3;
                    }
                }
            }
            __syncwarp(pmask);
            const long long t_d1 = B200LM_CLOCK();
            tk_d += t_d1 - t_p0;
            // ---------------- solve: next trial point of every fit that is ready ----------------
            if (st == 3 && status == -2 && it == 0) {
                const double* G = Gbuf + cur * NP * S1 + s;
                if (!fresh && nfev >= P.maxit) fresh = true;
                if (fresh) {
                    double xx = 0.0, gmax = 0.0;
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        const double pj = PV[j * S1 + s];
                        xx = fma(pj, pj, xx);
                        gmax = fmax(gmax, fabs(G[j * S1]));
                        DS[j * S1 + s] = 1.0 / SI[j * S1 + s];
                    }
                    if (gmax < P.gtol) status = 1;
                    else if (nfev >= P.maxit) status = 0;
                    xthr = P.xtol * (P.xtol + sqrt(xx));
                    gn_valid = false;
                    fresh = false;
                    __syncwarp(pmask);
                }
            }
            // The secular iteration runs in WARP-UNIFORM control flow: every lane executes the same loop, lanes
            // without work are predicated off, and the warp converges (__syncwarp) right before the
            // factorisation -- so the 32 factorisations of a round run as ONE pass through wave_factor_solve
            // instead of one pass per diverged pair.
            {
                const bool solving = (st == 3 && status == -2);
                const double* A = Abuf + cur * NPP * S1 + s;
                const double* G = Gbuf + cur * NP * S1 + s;
                const double* D = DS + s;
                if (solving && it == 0) {
                    double gg = 0.0;
#pragma unroll
                    for (int j = 0; j < NP; ++j) { const double t = D[j * S1] * G[j * S1]; gg = fma(t, t, gg); }
                    au = sqrt(gg) / Delta; al = 0.0;
                    best.ok = false; second.ok = false; use_gn = false;
                }
                bool done = !solving;
                {
                    double a_lo = 0.0, a_hi = 0.0;
                    int lo_kind = 0, hi_kind = 0;             // 0 candidate, 1 Gauss-Newton (shift 0), 2 ignore
                    if (!done) {
                        // straight-line selects and slow-path-free reciprocals: the 16 pairs of the warp take
                        // different cases, and branches here would serialise them
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
                            as = (!gn_valid && best.ok && best.pn < Delta) ? 0.0 : as;      // the step may be interior: Gauss-Newton is needed
                            a_lo = an; a_hi = as; hi_kind = (as == 0.0) ? 1 : 0;
                        }
                    }
                    __syncwarp();                                         // the convergence point of the warp
                    // one factorisation per thread of the pair, all pairs of the warp in the same pass
                    double pn_m = 0.0, w2_m = 0.0;
                    int ok_m = 0;
                    if (!done) {
                        const long long t_f0 = B200LM_CLOCK();
                        wave_factor_solve<NP, S1, ST>(A, G, D, ZM + tid, PT + tid, q ? a_hi : a_lo, &pn_m, &w2_m, &ok_m);
                        tk_fact += B200LM_CLOCK() - t_f0;
                    }
                    __syncwarp();
                    const double pn_o = __shfl_xor_sync(B200LM_FULL, pn_m, 1);
                    const double w2_o = __shfl_xor_sync(B200LM_FULL, w2_m, 1);
                    const int ok_o = __shfl_xor_sync(B200LM_FULL, ok_m, 1);
                    if (!done) {
                        nfac += (hi_kind == 2) ? 1 : 2;
                        WCand cd[2];
                        cd[0].a = a_lo; cd[1].a = a_hi;
                        cd[0].ok = (q ? ok_o : ok_m) != 0; cd[0].pn = q ? pn_o : pn_m; cd[0].w2 = q ? w2_o : w2_m;
                        cd[1].ok = (q ? ok_m : ok_o) != 0; cd[1].pn = q ? pn_m : pn_o; cd[1].w2 = q ? w2_m : w2_o;
                        const int kind[2] = {lo_kind, hi_kind};
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const WCand t = cd[h];
                            const bool mine = (q == h);
                            if (kind[h] == 1) {                                  // the Gauss-Newton step
                                gn_valid = true; gn_full = t.ok; gn_pn = t.pn; gn_w2 = t.w2;
                                if (mine) {
#pragma unroll
                                    for (int j = 0; j < NP; ++j) GN[j * S1 + s] = PT[j * ST + tid];
                                }
                            } else if (kind[h] == 0 && t.ok && t.a > 0.0) {      // a candidate shift
                                const double dd = fabs(t.pn - Delta);
                                if (!best.ok || dd < fabs(best.pn - Delta)) {
                                    second = best; best = t;
                                    if (mine) {
#pragma unroll
                                        for (int j = 0; j < NP; ++j) BP[j * S1 + s] = PT[j * ST + tid];
                                    }
                                } else if (!second.ok || dd < fabs(second.pn - Delta)) {
                                    second = t;
                                }
                            }
                        }
                        if (gn_valid && gn_full && gn_pn <= Delta) { alpha = 0.0; use_gn = true; done = true; }
                        else if (it > 0 && !cd[0].ok && (hi_kind == 1 || !cd[1].ok)) {   // not positive definite at these shifts: push up
                            al = fmax(al, fmax(a_lo, a_hi));
                            alpha = fmax(2.0 * fmax(a_lo, a_hi), 0.001 * au);
                            if (alpha > au) au = 2.0 * alpha;
                        }
                    }
                }
                __syncwarp();
                const long long t_s1 = B200LM_CLOCK();
                tk_sec += t_s1 - t_d1;
                // converged (same test the next round would start with), or out of rounds: take the step
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
                        // model of 1/|p(alpha)| for the next trial and warm start of its shift (as solve_tr_dual)
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
                        // no shift could be factorised: steepest descent to the boundary
#pragma unroll
                        for (int j = 0; j < NP; ++j) BP[j * S1 + s] = -(D[j * S1] * G[j * S1]);
                        double gg = 0.0;
#pragma unroll
                        for (int j = 0; j < NP; ++j) { const double t = D[j * S1] * G[j * S1]; gg = fma(t, t, gg); }
                        scale = gg > 0.0 ? Delta / sqrt(gg) : 0.0;
                        __syncwarp(pmask);
                    }
                    // step, trial point, predicted reduction = -(1/2 step^T A step + g^T step): rows split over the pair
                    double s2 = 0.0, h2 = 0.0, pred = 0.0;
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        const double sh = SH[j * S1] * scale;
                        const double sj = D[j * S1] * sh;
                        h2 = fma(sh, sh, h2);
                        s2 = fma(sj, sj, s2);
                        PT[j * ST + tid] = sj;
                        if ((j & 1) == q) PN[j * S1 + s] = PV[j * S1 + s] + sj;
                    }
#pragma unroll
                    for (int i2 = 0; i2 < NP; i2 += 2) {
                        const int i = i2 + q;
                        if (i < NP) {
                            double As0 = 0.0, As1 = 0.0;
#pragma unroll
                            for (int j = 0; j < NP; j += 2) {
                                As0 = fma(A[(i >= j ? wtri(i, j) : wtri(j, i)) * S1], PT[j * ST + tid], As0);
                                if (j + 1 < NP) As1 = fma(A[(i >= j + 1 ? wtri(i, j + 1) : wtri(j + 1, i)) * S1], PT[(j + 1) * ST + tid], As1);
                            }
                            pred = fma(-PT[i * ST + tid], 0.5 * (As0 + As1) + G[i * S1], pred);
                        }
                    }
                    pred += __shfl_xor_sync(pmask, pred, 1);
                    predicted = pred; step2 = s2; sh2 = h2;
                    st = 2;
                }
                tk_step += B200LM_CLOCK() - t_s1;
            }
            const long long t_f1 = B200LM_CLOCK();
            // ---------------- a finished fit leaves, the next one enters ----------------
            if (status != -2) {
                if (q == 0) {
#pragma unroll
                    for (int j = 0; j < NP; ++j) P.x_out[(size_t)fit * NP + j] = PV[j * S1 + s];
                    P.chi2[fit] = 2.0 * cost;
                    P.nit[fit] = nfev;
                    P.status[fit] = status;
                    tot_nfev += nfev; tot_nfac += nfac;
                }
                st = 0; fit = -1;
            }
            if (st == 0 && !queue_empty) {
                int b = 0;
                if (q == 0) b = atomicAdd(P.counter, 1);
                b = __shfl_sync(pmask, b, lane & ~1);
                if (b < P.B) {
                    fit = b; st = 1; cur = 0;
                    const double* p0 = P.p0 + (size_t)b * P.p0_stride;
#pragma unroll
                    for (int j = 0; j < NP; ++j) PV[j * S1 + s] = p0[j];
                } else {
                    queue_empty = true;
                }
            }
            if (q == 0 && tid < ST) {
                s_req[s] = st == 1 ? 1 : (st == 2 ? 2 : 0);
                s_tgt[s] = st == 1 ? cur : cur ^ 1;
                s_fit[s] = fit;
            }
        }
        const long long t_b0 = B200LM_CLOCK();
        tk_solve += t_b0 - t_p0;
        const int busy = __syncthreads_count(tid < ST && q == 0 && st != 0);
        tk_fin += B200LM_CLOCK() - t_b0;                 // (diagnostics: barrier wait, reported as "finish")
        if (busy == 0) break;

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
        tk_pass += B200LM_CLOCK() - t_p0;
    }
#ifdef B200LM_PHASE_TICKS
    if (tid == 32 && P.stats) atomicAdd(&P.stats[10], (unsigned long long)tk_solve);
    if (tid == 0 && P.stats) {
        atomicAdd(&P.stats[3], (unsigned long long)(tk_e1 + tk_e2 + tk_e3));
        atomicAdd(&P.stats[4], (unsigned long long)tk_solve);
        atomicAdd(&P.stats[5], (unsigned long long)tk_pass);
        atomicAdd(&P.stats[6], (unsigned long long)tk_e1);
        atomicAdd(&P.stats[7], (unsigned long long)tk_e2);
        atomicAdd(&P.stats[8], (unsigned long long)tk_e3);
        atomicAdd(&P.stats[9], (unsigned long long)n_pass);
        atomicAdd(&P.stats[11], (unsigned long long)tk_fact);
        atomicAdd(&P.stats[12], (unsigned long long)tk_d);
        atomicAdd(&P.stats[13], (unsigned long long)tk_sec);
        atomicAdd(&P.stats[14], (unsigned long long)tk_step);
        atomicAdd(&P.stats[15], (unsigned long long)tk_fin);
    }
#endif
    if (tid < ST && q == 0 && P.stats) {
        atomicAdd(&P.stats[0], tot_nfev);
        atomicAdd(&P.stats[1], tot_nfev);
        atomicAdd(&P.stats[2], tot_nfac);
    }
}

// shapes the wave kernel takes: exactly one correlated block of at most 64 points (any mix of data and prior
// entries), every other entry a 1x1 PRIOR row, scipy policy.  Everything else stays with the warp / team kernels.
// Two configurations: small CTAs (8 fits, 4 warps, weights through L1) of which several share an SM -- their
// solve and evaluation phases interleave --, or one large CTA per SM (32 fits, 8 warps, weights staged).
template <class F, int S_, int CF_, int MINB, bool WSMEM>
cudaError_t launch_fit_wave_cfg(FitParams P, int sm_count, size_t smem_budget, cudaStream_t stream) {
    typedef WaveLayout<F, S_, CF_> WL;
    const size_t smem = WL::bytes(WSMEM ? P.wt_total : 0);
    if (smem > smem_budget) return cudaErrorInvalidConfiguration;
    auto kern = fit_wave_kernel<F, S_, CF_, MINB, WSMEM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WL::THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > MINB) per_sm = MINB;
    int grid = sm_count * per_sm;
    const int want = (P.B + (S_ / 2) - 1) / (S_ / 2);          // small batches: spread over the SMs, slots at least half full
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    kern<<<grid, WL::THREADS, smem, stream>>>(P);
    return cudaGetLastError();
}

template <class F>
cudaError_t launch_fit_wave(FitParams P, int sm_count, size_t smem_budget, cudaStream_t stream) {
    // measured on C3 (B = 160 k, tools/wave_cfg.py): one CTA of 32 fits / 8 warps per SM 84.8 ms, two CTAs of 16 fits /
    // 4 warps 91.0 ms, three CTAs of 8 fits / 4 warps 139.8 ms -- the evaluation phase is latency bound, so halving
    // the warps of a CTA doubles its time and the overlap of solve and evaluation phases between CTAs buys nothing
    if (const char* env = getenv("B200LM_WAVE_CFG"))                      // tuning knob (profiling only)
        if (atoi(env) == 2) return launch_fit_wave_cfg<F, 16, 4, 2, false>(P, sm_count, smem_budget, stream);
    return launch_fit_wave_cfg<F, 32, 8, 1, true>(P, sm_count, smem_budget, stream);
}

template <class F>
size_t wave_bytes_of(int wt_total) { return WaveLayout<F, 32, 8>::bytes(wt_total); }

}  // namespace b200lm
