// Batched trust-region Levenberg-Marquardt: one warp per fit, persistent CTAs
// pulling fits from a device work queue, every decision kept on the device.
//
// What it replaces in the reference (paths relative to the reference tree):
//   * chiv.__call__            src/lsqfit/_utilities.pyx:65-94   (whiten + residual)
//   * Dfun / valder Jacobian   src/lsqfit/_scipy.py:144-154
//   * the fitter plugin        src/lsqfit/_scipy.py:156-181 / src/lsqfit/_gsl.pyx:563-723
//   * chi2 / logdet(J^T J)     src/lsqfit/__init__.py:665-682, 706-725
//
// Algorithm (host model: tests/lm_model.py).  The trust-region *decisions* are those
// of the solver behind the reference's scipy plugin (unbounded TRF: radius update
// 0.25/0.75, Delta0 = |x0*scale|, accept iff cost decreases, ftol/xtol/gtol tests),
// with More' column scaling (GSL `scaler='more'`, scipy x_scale='jac') by default.
// The Levenberg parameter is found by Newton iteration on the secular equation,
// evaluated with a Cholesky factorisation of the scaled normal matrix
//       d (J^T J) d + alpha I
// held in shared memory -- not by an SVD/QR of J.
//
// Data layout per warp (shared memory):
//   R[rb][LDR]   row buffer: rows of [G | delta] before whitening, [J | r] after
//   A[NP][LDA]   J^T J (unscaled)        L[NP][LDA]   Cholesky factor
//   p, pn, g, sinv, dsc, idg   NP-vectors
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "lm_types.h"

namespace b200lm {

#define B200LM_FULL 0xffffffffu

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(B200LM_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(B200LM_FULL, v, o));
    return v;
}

template <class F>
struct FitLayout {
    static constexpr int NP = F::NP;
    static constexpr int NP1 = NP + 1;
    static constexpr int LDR = NP1 | 1;      // odd: lane-per-row accesses are conflict free
    static constexpr int LDA = NP | 1;
    static constexpr int NVEC = 6;
    // register budget: 16 warps/CTA leave 128 registers per thread, 12 warps leave 168
    static constexpr int MAX_WARPS = NP > 10 ? 12 : 16;
    __host__ __device__ static int per_warp_doubles(int rb) {
        int n = rb * LDR + 2 * NP * LDA + NVEC * NP;
        return (n + 1) & ~1;
    }
};

template <class F>
struct WarpCtx {
    typedef FitLayout<F> Lay;
    const FitParams& P;
    const double* wt;       // whitening matrices (shared or global)
    const double* mean;     // this fit's y(+)prior means
    double *R, *A, *L, *p, *pn, *g, *sinv, *dsc, *idg;
    int lane;
    __device__ WarpCtx(const FitParams& P_) : P(P_) {}
};

// ---------------------------------------------------------------------------
// residual only: returns cost = 1/2 sum r^2 (same value in every lane)
// ---------------------------------------------------------------------------
template <class F>
__device__ double eval_cost(WarpCtx<F>& c, const double* pv, double* fout) {
    typedef FitLayout<F> Lay;
    const FitParams& P = c.P;
    const int lane = c.lane;
    double acc = 0.0;
    for (int i = lane; i < P.nd_fn; i += 32) {
        const int row = P.dfn_idx[i];
        const double f = F::value(P.x + (size_t)row * P.nx, row, pv);
        const double r = P.dfn_w[i] * (f - c.mean[row]);
        acc = fma(r, r, acc);
        if (fout) fout[i] = r;
    }
    for (int i = lane; i < P.nd_pr; i += 32) {
        const int idx = P.dpr_idx[i];
        const double r = P.dpr_w[i] * (pv[idx - P.ny] - c.mean[idx]);
        acc = fma(r, r, acc);
        if (fout) fout[P.nd_fn + i] = r;
    }
    double* dv = c.R;
    for (int b = 0; b < P.nblk; ++b) {
        const BlockDesc bd = P.blk[b];
        for (int k = lane; k < bd.n_in; k += 32) {
            const int idx = P.blk_idx[bd.idx_off + k];
            const double v = idx < P.ny ? F::value(P.x + (size_t)idx * P.nx, idx, pv) : pv[idx - P.ny];
            dv[k] = v - c.mean[idx];
        }
        __syncwarp();
        const double* wt = c.wt + bd.wt_off;
        for (int r = lane; r < bd.n_out; r += 32) {
            double s = 0.0;
            for (int k = 0; k < bd.n_in; ++k) s = fma(wt[(size_t)k * bd.ldw + r], dv[k], s);
            acc = fma(s, s, acc);
            if (fout) fout[bd.chiv_off + r] = s;
        }
        __syncwarp();
    }
    return 0.5 * warp_sum(acc);
}

// ---------------------------------------------------------------------------
// accumulate A += S^T S, g += S^T r, cost += r^T r over `nrows` rows of S = [J | r]
// ---------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void accumulate(WarpCtx<F>& c, const double* S, int nrows, double& acc_cost) {
    typedef FitLayout<F> Lay;
    constexpr int NP = Lay::NP, NP1 = Lay::NP1, LDR = Lay::LDR, LDA = Lay::LDA;
    constexpr int T = NP1 * (NP1 + 1) / 2;
    for (int e = c.lane; e < T; e += 32) {
        int a = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
        while ((a + 1) * (a + 2) / 2 <= e) ++a;
        while (a * (a + 1) / 2 > e) --a;
        const int b = e - a * (a + 1) / 2;           // b <= a <= NP
        double s0 = 0.0, s1 = 0.0;
        int i = 0;
        for (; i + 1 < nrows; i += 2) {
            s0 = fma(S[i * LDR + a], S[i * LDR + b], s0);
            s1 = fma(S[(i + 1) * LDR + a], S[(i + 1) * LDR + b], s1);
        }
        if (i < nrows) s0 = fma(S[i * LDR + a], S[i * LDR + b], s0);
        const double s = s0 + s1;
        if (a < NP) {
            c.A[a * LDA + b] += s;
            if (a != b) c.A[b * LDA + a] += s;
        } else if (b < NP) {
            c.g[b] += s;
        } else {
            acc_cost += s;
        }
    }
}

template <class F>
__device__ __forceinline__ void emit_rows(const double* S, int nrows, int slot0, int lane,
                                          double* fout, double* Jout) {
    typedef FitLayout<F> Lay;
    constexpr int NP = Lay::NP, LDR = Lay::LDR;
    for (int i = lane; i < nrows; i += 32) {
        if (fout) fout[slot0 + i] = S[i * LDR + NP];
        if (Jout) {
#pragma unroll
            for (int j = 0; j < NP; ++j) Jout[(size_t)(slot0 + i) * NP + j] = S[i * LDR + j];
        }
    }
}

// ---------------------------------------------------------------------------
// residual + Jacobian + normal equations at pv: fills c.A (J^T J), c.g (J^T r),
// returns cost.  Optionally writes the residual vector and J to global memory.
// ---------------------------------------------------------------------------
// Householder update of the triangular factor Rt (in c.A) with `nrows` rows of S, columns
// scaled by c.dsc:  Rt <- R of qr([Rt ; S.diag(dsc)]).  Rows are owned by lanes (i mod 32);
// S is destroyed.  Used for the final covariance: (J^T J)^-1 = dsc Rt^-1 Rt^-T dsc has a
// relative error ~kappa(J)*eps instead of kappa^2*eps for the normal equations.
template <class F>
__device__ void qr_update(WarpCtx<F>& c, double* S, int nrows) {
    typedef FitLayout<F> Lay;
    constexpr int NP = Lay::NP, LDR = Lay::LDR, LDA = Lay::LDA;
    const int lane = c.lane;
    for (int i = lane; i < nrows; i += 32) {
#pragma unroll
        for (int j = 0; j < NP; ++j) S[i * LDR + j] *= c.dsc[j];
    }
    __syncwarp();
    for (int j = 0; j < NP; ++j) {
        double ss = 0.0;
        for (int i = lane; i < nrows; i += 32) { const double v = S[i * LDR + j]; ss = fma(v, v, ss); }
        ss = warp_sum(ss);
        if (ss == 0.0) continue;                             // uniform across the warp
        const double rjj = c.A[j * LDA + j];
        const double nrm = sqrt(fma(rjj, rjj, ss));
        const double alpha = rjj > 0.0 ? -nrm : nrm;
        const double v0 = rjj - alpha;
        const double beta = 1.0 / (nrm * fabs(v0));
        double t[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) t[q] = 0.0;
        for (int i = lane; i < nrows; i += 32) {
            const double sj = S[i * LDR + j];
#pragma unroll
            for (int q = 0; q < NP; ++q) if (q > j) t[q] = fma(sj, S[i * LDR + q], t[q]);
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) if (q > j) t[q] = (fma(v0, c.A[j * LDA + q], warp_sum(t[q]))) * beta;
        __syncwarp();
        for (int i = lane; i < nrows; i += 32) {
            const double sj = S[i * LDR + j];
#pragma unroll
            for (int q = 0; q < NP; ++q) if (q > j) S[i * LDR + q] = fma(-t[q], sj, S[i * LDR + q]);
        }
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < NP; ++q) if (q > j) c.A[j * LDA + q] = fma(-t[q], v0, c.A[j * LDA + q]);
            c.A[j * LDA + j] = alpha;
        }
        __syncwarp();
    }
}

// MODE 0: normal equations (A = J^T J, g = J^T r).  MODE 1: QR factor of J.diag(dsc) in c.A.
template <class F, int MODE = 0>
__device__ double eval_full(WarpCtx<F>& c, const double* pv, double* fout, double* Jout) {
    typedef FitLayout<F> Lay;
    constexpr int NP = Lay::NP, LDR = Lay::LDR, LDA = Lay::LDA;
    const FitParams& P = c.P;
    const int lane = c.lane;
    for (int e = lane; e < NP * LDA; e += 32) c.A[e] = 0.0;
    if (MODE == 0 && lane < NP) c.g[lane] = 0.0;
    __syncwarp();
    double acc = 0.0;
    // 1x1 prior rows: J row = w e_j, analytic contribution
    for (int i = lane; i < P.nd_pr; i += 32) {
        const int idx = P.dpr_idx[i];
        const int j = idx - P.ny;
        const double w = P.dpr_w[i];
        const double r = w * (pv[j] - c.mean[idx]);
        if (MODE == 0) {
            c.A[j * LDA + j] += w * w;
            c.g[j] += w * r;
        } else {
            c.A[j * LDA + j] = w * c.dsc[j];                 // a diagonal matrix is triangular
        }
        acc = fma(r, r, acc);
        if (fout) fout[P.nd_fn + i] = r;
        if (Jout) {
            for (int q = 0; q < NP; ++q) Jout[(size_t)(P.nd_fn + i) * NP + q] = (q == j) ? w : 0.0;
        }
    }
    __syncwarp();
    // 1x1 data rows, staged through the row buffer in chunks
    for (int c0 = 0; c0 < P.nd_fn; c0 += P.rb) {
        const int nrows = min(P.rb, P.nd_fn - c0);
        for (int i = lane; i < nrows; i += 32) {
            const int row = P.dfn_idx[c0 + i];
            const double w = P.dfn_w[c0 + i];
            double gr[NP];
            const double f = F::value_grad(P.x + (size_t)row * P.nx, row, pv, gr);
#pragma unroll
            for (int j = 0; j < NP; ++j) c.R[i * LDR + j] = w * gr[j];
            c.R[i * LDR + NP] = w * (f - c.mean[row]);
        }
        __syncwarp();
        if (fout || Jout) emit_rows<F>(c.R, nrows, c0, lane, fout, Jout);
        if (MODE == 0) accumulate<F>(c, c.R, nrows, acc);
        else { __syncwarp(); qr_update<F>(c, c.R, nrows); }
        __syncwarp();
    }
    // correlated blocks: J_blk = W.[G | delta].  Output rows are owned by lanes (two per lane per
    // 64-row group) and stay in registers while the input rows [G | delta] stream through the
    // row buffer in chunks of P.rb rows; the finished rows then pass through the same buffer,
    // 32 at a time, into the normal equations (or the QR update).
    for (int b = 0; b < P.nblk; ++b) {
        const BlockDesc bd = P.blk[b];
        const double* wt = c.wt + bd.wt_off;
        for (int g0 = 0; g0 < bd.n_out; g0 += 64) {
            const int r0 = g0 + lane, r1 = g0 + lane + 32;
            const bool h0 = r0 < bd.n_out, h1 = r1 < bd.n_out;
            double a0[NP + 1], a1[NP + 1];
#pragma unroll
            for (int j = 0; j <= NP; ++j) { a0[j] = 0.0; a1[j] = 0.0; }
            for (int k0 = 0; k0 < bd.n_in; k0 += P.rb) {
                const int nk = min(P.rb, bd.n_in - k0);
                for (int k = lane; k < nk; k += 32) {
                    const int idx = P.blk_idx[bd.idx_off + k0 + k];
                    if (idx < P.ny) {
                        double gr[NP];
                        const double f = F::value_grad(P.x + (size_t)idx * P.nx, idx, pv, gr);
#pragma unroll
                        for (int j = 0; j < NP; ++j) c.R[k * LDR + j] = gr[j];
                        c.R[k * LDR + NP] = f - c.mean[idx];
                    } else {
                        const int j0 = idx - P.ny;
#pragma unroll
                        for (int j = 0; j < NP; ++j) c.R[k * LDR + j] = (j == j0) ? 1.0 : 0.0;
                        c.R[k * LDR + NP] = pv[j0] - c.mean[idx];
                    }
                }
                __syncwarp();
                const double* wk = wt + (size_t)k0 * bd.ldw;
#pragma unroll 2
                for (int k = 0; k < nk; ++k) {
                    const double w0 = h0 ? wk[(size_t)k * bd.ldw + r0] : 0.0;
                    const double w1 = h1 ? wk[(size_t)k * bd.ldw + r1] : 0.0;
#pragma unroll
                    for (int j = 0; j <= NP; ++j) {
                        const double v = c.R[k * LDR + j];
                        a0[j] = fma(w0, v, a0[j]);
                        a1[j] = fma(w1, v, a1[j]);
                    }
                }
                __syncwarp();                      // everyone has finished reading this chunk
            }
            // rows g0 .. g0+31
            if (h0) {
#pragma unroll
                for (int j = 0; j <= NP; ++j) c.R[lane * LDR + j] = a0[j];
            }
            __syncwarp();
            int nrows = min(32, bd.n_out - g0);
            if (fout || Jout) emit_rows<F>(c.R, nrows, bd.chiv_off + g0, lane, fout, Jout);
            if (MODE == 0) accumulate<F>(c, c.R, nrows, acc);
            else { __syncwarp(); qr_update<F>(c, c.R, nrows); }
            __syncwarp();
            // rows g0+32 .. g0+63
            nrows = min(32, bd.n_out - g0 - 32);
            if (nrows > 0) {
                if (h1) {
#pragma unroll
                    for (int j = 0; j <= NP; ++j) c.R[lane * LDR + j] = a1[j];
                }
                __syncwarp();
                if (fout || Jout) emit_rows<F>(c.R, nrows, bd.chiv_off + g0 + 32, lane, fout, Jout);
                if (MODE == 0) accumulate<F>(c, c.R, nrows, acc);
                else { __syncwarp(); qr_update<F>(c, c.R, nrows); }
                __syncwarp();
            }
        }
    }
    return 0.5 * warp_sum(acc);
}

// ---------------------------------------------------------------------------
// small dense kernels on the warp: lane i owns row i, everything in registers
// ---------------------------------------------------------------------------
// Factor of the scaled, shifted normal matrix held in registers: lane i keeps row i of L in
// l[] (entries k <= i), column i of L in u[] (entries k >= i) and 1/L_ii in inv.
template <int NP>
struct CholReg {
    double l[NP];
    double u[NP];
    double inv;
};

// L L^T = d_i A_ij d_j + alpha delta_ij, right-looking, fully unrolled: the column of the
// current step is broadcast with warp shuffles.  false if not numerically positive definite.
template <int NP, int LDA>
__device__ __forceinline__ bool chol_factor(const double* A, const double* dsc, int lane, double alpha,
                                            CholReg<NP>& f, double& minr) {
    const int i = lane;
    const bool act = i < NP;
    const double di = act ? dsc[i] : 0.0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        double v = act ? A[i * LDA + k] * di * dsc[k] : 0.0;
        if (k == i) v = act ? v + alpha : 1.0;
        f.l[k] = v;
        f.u[k] = 0.0;
    }
    double mdiag = 1.0;                 // original diagonal entry of this lane's row
#pragma unroll
    for (int k = 0; k < NP; ++k) if (k == i) mdiag = f.l[k];
    minr = 1.0;
    bool ok = true;
    f.inv = 1.0;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const double piv = __shfl_sync(B200LM_FULL, f.l[j], j);
        const double mjj = __shfl_sync(B200LM_FULL, mdiag, j);
        // no early exit: a `break` would keep the loop from unrolling
        if (!(piv > 8.0 * NP * 2.220446049250313e-16 * mjj) || !isfinite(piv)) ok = false;
        minr = fmin(minr, piv / mjj);
        const double inv = rsqrt(piv);
        if (i == j) f.inv = inv;
        const double lij = (i == j) ? piv * inv : f.l[j] * inv;      // L[i][j]
        f.l[j] = lij;
        if (i == j) f.u[j] = lij;
#pragma unroll
        for (int k = j + 1; k < NP; ++k) {
            const double lkj = __shfl_sync(B200LM_FULL, lij, k);      // L[k][j]
            f.l[k] = fma(-lij, lkj, f.l[k]);                          // row i, column k (used for k <= i)
            if (i == j) f.u[k] = lkj;                                 // column j of L kept by lane j
        }
    }
    return ok;
}
// y = L^-1 b (lane i holds b_i, returns y_i)
template <int NP>
__device__ __forceinline__ double solve_lower(const CholReg<NP>& f, int lane, double b) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const double yj = __shfl_sync(B200LM_FULL, b * f.inv, j);
        if (lane == j) b = yj;
        else if (lane > j) b = fma(-f.l[j], yj, b);
    }
    return b;
}
// x = L^-T b
template <int NP>
__device__ __forceinline__ double solve_upper(const CholReg<NP>& f, int lane, double b) {
#pragma unroll
    for (int j = NP - 1; j >= 0; --j) {
        const double xj = __shfl_sync(B200LM_FULL, b * f.inv, j);
        if (lane == j) b = xj;
        else if (lane < j) b = fma(-f.u[j], xj, b);
    }
    return b;
}

// One factorisation + the solves every caller needs, as ONE out-of-line function (the
// unrolled factor is ~1.5k instructions; five inlined copies would thrash the i-cache and
// the register allocator).  With the lane-distributed scaled gradient gh:
//     p = -(Ah + alpha I)^-1 gh ,  res[0] = |p| ,  res[1] = |L^-1 p|^2 ,  res[2] = min pivot ratio
// Optionally leaves L (rows) and 1/L_ii in shared memory for the covariance routine.
template <int NP, int LDA>
__device__ __noinline__ bool factor_solve(const double* A, const double* dsc, double* Lsm, double* idg,
                                          int lane, double alpha, double gh, bool store,
                                          double* p_out, double* res) {
    CholReg<NP> f;
    double minr;
    const bool ok = chol_factor<NP, LDA>(A, dsc, lane, alpha, f, minr);
    const bool act = lane < NP;
    double p = 0.0, pn = 0.0, w2 = 0.0;
    if (ok) {
        p = solve_upper<NP>(f, lane, solve_lower<NP>(f, lane, act ? -gh : 0.0));
        pn = sqrt(warp_sum(act ? p * p : 0.0));
        const double w = solve_lower<NP>(f, lane, act ? p : 0.0);
        w2 = warp_sum(act ? w * w : 0.0);
        if (store) {
            if (act) {
#pragma unroll
                for (int k = 0; k < NP; ++k) Lsm[lane * LDA + k] = (k <= lane) ? f.l[k] : 0.0;
                idg[lane] = f.inv;
            }
            __syncwarp();
        }
    }
    *p_out = p;
    res[0] = pn; res[1] = w2; res[2] = minr;
    return ok;
}

// Gauss-Newton step of the current outer iteration (alpha = 0), cached across rejected trials
struct GNCache {
    bool valid, full_rank;
    double p, pn, w2;
};

// Trust-region sub-problem in scaled variables: min 1/2 s^T Ah s + gh^T s, |s| <= Delta.
// gh: lane-distributed scaled gradient.  Returns lane-distributed step; updates alpha.
// (cf. scipy/optimize/_lsq/common.py: solve_lsq_trust_region, with the secular equation
// evaluated through Cholesky factors instead of singular values.)
template <class F>
__device__ double solve_tr(WarpCtx<F>& c, double gh, double Delta, double& alpha, int& nfac, GNCache& gn) {
    typedef FitLayout<F> Lay;
    constexpr int NP = Lay::NP, LDA = Lay::LDA;
    const int lane = c.lane;
    const bool act = lane < NP;
    double res[3];
    double p = 0.0;
    if (!gn.valid) {
        ++nfac;
        gn.full_rank = factor_solve<NP, LDA>(c.A, c.dsc, c.L, c.idg, lane, 0.0, gh, false, &p, res);
        gn.p = p; gn.pn = res[0]; gn.w2 = res[1];
        gn.valid = true;
    }
    const bool full_rank = gn.full_rank;
    if (full_rank && gn.pn <= Delta) { alpha = 0.0; return gn.p; }
    p = gn.p;
    double pn = gn.pn;
    double alpha_upper = sqrt(warp_sum(act ? gh * gh : 0.0)) / Delta;
    double alpha_lower = 0.0;
    if (full_rank) {
        const double phi = pn - Delta;
        const double phi_prime = -gn.w2 / pn;
        alpha_lower = -phi / phi_prime;
    }
    if (!full_rank && alpha == 0.0)
        alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
    bool have_p = false, converged = false;
    // iterations 0..9: Newton on the secular equation; iteration 10: the step at the final alpha
    for (int it = 0; it <= 10; ++it) {
        const bool last = converged || it == 10;
        if (!last && (alpha < alpha_lower || alpha > alpha_upper))
            alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
        ++nfac;
        double pt;
        const bool ok = factor_solve<NP, LDA>(c.A, c.dsc, c.L, c.idg, lane, alpha, gh, false, &pt, res);
        if (last) {
            if (ok) { p = pt; have_p = true; }
            break;
        }
        if (!ok) {
            alpha_lower = fmax(alpha_lower, alpha);
            alpha = fmax(2.0 * alpha, 0.001 * alpha_upper);
            if (alpha > alpha_upper) alpha_upper = 2.0 * alpha;
            continue;
        }
        p = pt;
        have_p = true;
        pn = res[0];
        const double phi = pn - Delta;
        const double phi_prime = -res[1] / pn;
        if (phi < 0.0) alpha_upper = alpha;
        const double ratio = phi / phi_prime;
        alpha_lower = fmax(alpha_lower, alpha - ratio);
        alpha -= (phi + Delta) * ratio / Delta;
        if (fabs(phi) < 0.01 * Delta) converged = true;
    }
    if (!have_p) p = act ? -gh : 0.0;             // steepest descent fallback
    pn = sqrt(warp_sum(act ? p * p : 0.0));
    if (pn > 0.0) p *= Delta / pn;
    return p;
}

// covariance (J^T J)^-1 = d (L L^T)^-1 d from the factor of the scaled matrix at alpha=0.
// Uses c.A as scratch for L^-1 (column a computed by lane a).  Returns log det(J^T J).
template <class F>
__device__ double covariance_from_chol(WarpCtx<F>& c, double* cov_out) {
    typedef FitLayout<F> Lay;
    constexpr int NP = Lay::NP, LDA = Lay::LDA;
    const int a = c.lane;
    double ld = 0.0;
    if (a < NP) ld = 2.0 * (log(c.L[a * LDA + a]) - log(c.dsc[a]));
    ld = warp_sum(ld);
    __syncwarp();
    if (a < NP) {
        // column a of Linv: x_j = (delta_ja - sum_{a<=k<j} L[j][k] x_k) / L_jj , j >= a
        for (int j = 0; j < NP; ++j) {
            double s = (j == a) ? 1.0 : 0.0;
            if (j < a) { c.A[j * LDA + a] = 0.0; continue; }
            for (int k = a; k < j; ++k) s = fma(-c.L[j * LDA + k], c.A[k * LDA + a], s);
            c.A[j * LDA + a] = s * c.idg[j];
        }
    }
    __syncwarp();
    if (a < NP && cov_out) {
        for (int b = 0; b < NP; ++b) {
            double s = 0.0;
            const int k0 = a > b ? a : b;
            for (int k = k0; k < NP; ++k) s = fma(c.A[k * LDA + a], c.A[k * LDA + b], s);
            cov_out[a * NP + b] = s * c.dsc[a] * c.dsc[b];
        }
    }
    __syncwarp();
    return ld;
}

// covariance from the triangular factor Rt (c.A) of J.diag(dsc): (J^T J)^-1 =
// dsc Rt^-1 Rt^-T dsc.  Uses c.L as scratch for Rt^-1.  Returns log det(J^T J) (nan if singular).
template <class F>
__device__ double covariance_from_qr(WarpCtx<F>& c, double* cov_out) {
    typedef FitLayout<F> Lay;
    constexpr int NP = Lay::NP, LDA = Lay::LDA;
    const int a = c.lane;
    double ld = 0.0;
    if (a < NP) ld = 2.0 * (log(fabs(c.A[a * LDA + a])) - log(c.dsc[a]));
    ld = warp_sum(ld);
    if (a < NP) {
        // column a of Rt^-1 (upper triangular): back substitution from row a up to row 0
        for (int j = NP - 1; j > a; --j) c.L[j * LDA + a] = 0.0;
        for (int j = a; j >= 0; --j) {
            double s = (j == a) ? 1.0 : 0.0;
            for (int k = j + 1; k <= a; ++k) s = fma(-c.A[j * LDA + k], c.L[k * LDA + a], s);
            c.L[j * LDA + a] = s / c.A[j * LDA + j];
        }
    }
    __syncwarp();
    if (a < NP && cov_out) {
        for (int b = 0; b < NP; ++b) {
            double s = 0.0;
            const int k0 = a > b ? a : b;
            for (int k = k0; k < NP; ++k) s = fma(c.L[a * LDA + k], c.L[b * LDA + k], s);
            cov_out[a * NP + b] = s * c.dsc[a] * c.dsc[b];
        }
    }
    __syncwarp();
    return ld;
}

// ---------------------------------------------------------------------------
// the fit kernel
// ---------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void setup_ctx(WarpCtx<F>& c, double* smem, const FitParams& P) {
    typedef FitLayout<F> Lay;
    const int warp = threadIdx.x >> 5;
    c.lane = threadIdx.x & 31;
    const int wt_region = P.wt_in_smem ? ((P.wt_total + 1) & ~1) : 0;
    if (P.wt_in_smem) {
        for (int i = threadIdx.x; i < P.wt_total; i += blockDim.x) smem[i] = P.blk_wt[i];
        c.wt = smem;
    } else {
        c.wt = P.blk_wt;
    }
    double* base = smem + wt_region + (size_t)warp * Lay::per_warp_doubles(P.rb);
    c.R = base;
    c.A = c.R + (size_t)P.rb * Lay::LDR;
    c.L = c.A + Lay::NP * Lay::LDA;
    c.p = c.L + Lay::NP * Lay::LDA;
    c.pn = c.p + Lay::NP;
    c.g = c.pn + Lay::NP;
    c.sinv = c.g + Lay::NP;
    c.dsc = c.sinv + Lay::NP;
    c.idg = c.dsc + Lay::NP;
    __syncthreads();
}

template <class F>
__global__ void __launch_bounds__(FitLayout<F>::MAX_WARPS * 32, 1) fit_kernel(const __grid_constant__ FitParams P) {
    typedef FitLayout<F> Lay;
    constexpr int NP = Lay::NP, LDA = Lay::LDA;
    static_assert(NP <= 32, "one lane per parameter");
    extern __shared__ double smem[];
    WarpCtx<F> c(P);
    setup_ctx<F>(c, smem, P);
    const int lane = c.lane;
    const bool act = lane < NP;
    unsigned long long tot_nfev = 0, tot_njev = 0, tot_nfac = 0;

    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(P.counter, 1);
        b = __shfl_sync(B200LM_FULL, b, 0);
        if (b >= P.B) break;
        c.mean = P.mean + (size_t)b * P.mean_stride;
        const double* p0 = P.p0 + (size_t)b * P.p0_stride;
        if (act) c.p[lane] = p0[lane];
        __syncwarp();

        double cost = eval_full<F>(c, c.p, nullptr, nullptr);
        int nfev = 1, njev = 1, nfac = 0;
        int status = -2;                         // -2: running
        if (!isfinite(cost)) status = -1;
        // scale_inv_j = |J_j| = sqrt(A_jj)   (More': running max; scaler 0: 1)
        double sinv = 1.0;
        if (act && P.scaler == 1) {
            sinv = sqrt(c.A[lane * LDA + lane]);
            if (sinv == 0.0) sinv = 1.0;
        }
        double t = act ? c.p[lane] * sinv : 0.0;
        double Delta = sqrt(warp_sum(t * t));
        if (Delta == 0.0) Delta = 1.0;
        double alpha = 0.0;

        while (status == -2) {
            const double gi = act ? c.g[lane] : 0.0;
            const double g_norm = warp_max(fabs(gi));
            if (g_norm < P.gtol) { status = 1; break; }
            if (nfev >= P.maxit) { status = 0; break; }
            const double d = 1.0 / sinv;
            if (act) c.dsc[lane] = d;
            __syncwarp();
            const double gh = d * gi;
            double actual_reduction = -1.0, cost_new = cost;
            int term = -2;
            GNCache gn;
            gn.valid = false;
            while (actual_reduction <= 0.0 && nfev < P.maxit) {
                const double sh = solve_tr<F>(c, gh, Delta, alpha, nfac, gn);
                const double step = act ? d * sh : 0.0;
                if (act) { c.pn[lane] = c.p[lane] + step; c.idg[lane] = step; }
                __syncwarp();
                // predicted reduction = -(1/2 step^T A step + g^T step)   (unscaled == scaled)
                double As = 0.0;
                if (act) {
                    for (int j = 0; j < NP; ++j) As = fma(c.A[lane * LDA + j], c.idg[j], As);
                }
                const double predicted = -warp_sum(act ? step * (0.5 * As + gi) : 0.0);
                __syncwarp();
                cost_new = eval_cost<F>(c, c.pn, nullptr);
                ++nfev;
                const double shn = sqrt(warp_sum(act ? sh * sh : 0.0));
                if (!isfinite(cost_new)) { Delta = 0.25 * shn; continue; }
                actual_reduction = cost - cost_new;
                double ratio;
                if (predicted > 0.0) ratio = actual_reduction / predicted;
                else if (predicted == 0.0 && actual_reduction == 0.0) ratio = 1.0;
                else ratio = 0.0;
                double Delta_new = Delta;
                if (ratio < 0.25) Delta_new = 0.25 * shn;
                else if (ratio > 0.75 && shn > 0.95 * Delta) Delta_new = 2.0 * Delta;
                const double step_norm = sqrt(warp_sum(step * step));
                const double pi_ = act ? c.p[lane] : 0.0;
                const double x_norm = sqrt(warp_sum(pi_ * pi_));
                const bool ft = actual_reduction < P.ftol * cost && ratio > 0.25;
                const bool xt = step_norm < P.xtol * (P.xtol + x_norm);
                if (ft && xt) term = 4; else if (ft) term = 2; else if (xt) term = 3;
                if (term != -2) break;
                alpha *= Delta / Delta_new;
                Delta = Delta_new;
            }
            if (actual_reduction > 0.0) {
                if (act) c.p[lane] = c.pn[lane];
                __syncwarp();
                cost = eval_full<F>(c, c.p, nullptr, nullptr);   // J, J^T J, J^T r at the new point
                ++njev;
                if (act && P.scaler == 1) sinv = fmax(sinv, sqrt(c.A[lane * LDA + lane]));
            }
            if (term != -2) {
                // the solver behind the reference re-tests gtol before leaving (trf.py loop head)
                const double gf = warp_max(act ? fabs(c.g[lane]) : 0.0);
                status = gf < P.gtol ? 1 : term;
            }
        }

        // ---- optional Gauss-Newton polish (P.polish > 0; off by default) ----------------
        // The trust-region loop accepts a step only if the cost decreases, and cost differences
        // below eps*cost cannot be resolved in fp64: it stalls ~sqrt(eps*chi2) standard deviations
        // from the stationary point (so does the reference's solver).  Undamped Gauss-Newton
        // steps, accepted on the decrease of the Newton decrement instead of the cost, take the
        // solution on to the gradient's rounding level.
        if (status >= 1 && P.polish > 0) {
            // Newton decrement dec = g^T (J^T J)^-1 g = predicted decrease of chi2; sqrt(dec) is
            // the distance to the stationary point in standard deviations.
            const double d = 1.0 / sinv;
            if (act) c.dsc[lane] = d;
            __syncwarp();
            double dec_prev = 1e300;
            for (int it = 0; it <= P.polish; ++it) {
                ++nfac;
                const double gh = act ? d * c.g[lane] : 0.0;
                double sh, pres[3];
                if (!factor_solve<NP, LDA>(c.A, c.dsc, c.L, c.idg, lane, 0.0, gh, false, &sh, pres)) break;
                const double dec = -warp_sum(act ? gh * sh : 0.0);
                if (it > 0) {
                    if (!(dec < dec_prev)) {                        // the last step did not help: undo it
                        if (act) { const double t = c.p[lane]; c.p[lane] = c.pn[lane]; c.pn[lane] = t; }
                        __syncwarp();
                        cost = eval_full<F>(c, c.p, nullptr, nullptr);
                        ++njev;
                        break;
                    }
                }
                if (!(dec > 1e-22) || it == P.polish) break;        // < 1e-11 sdev from stationarity
                dec_prev = dec;
                // keep the old point in pn, step to the new one
                if (act) { const double t = c.p[lane]; c.pn[lane] = t; c.p[lane] = t + d * sh; }
                __syncwarp();
                const double cost_try = eval_full<F>(c, c.p, nullptr, nullptr);
                ++nfev; ++njev;
                if (!isfinite(cost_try)) {
                    if (act) c.p[lane] = c.pn[lane];
                    __syncwarp();
                    cost = eval_full<F>(c, c.p, nullptr, nullptr);
                    ++njev;
                    break;
                }
                cost = cost_try;
            }
        }

        // ---- results -----------------------------------------------------------
        if (P.f_out || P.J_out) {
            cost = eval_full<F>(c, c.p, P.f_out ? P.f_out + (size_t)b * P.nchiv : nullptr,
                                P.J_out ? P.J_out + (size_t)b * P.nchiv * NP : nullptr);
        }
        // covariance at the solution, scaled by the current column norms for conditioning
        double dfin = 1.0;
        if (act) {
            const double s = sqrt(c.A[lane * LDA + lane]);
            dfin = s > 0.0 ? 1.0 / s : 1.0;
            c.dsc[lane] = dfin;
        }
        __syncwarp();
        ++nfac;
        double pdummy, fres[3];
        const bool okc = factor_solve<NP, LDA>(c.A, c.dsc, c.L, c.idg, lane, 0.0, 0.0, true, &pdummy, fres);
        const double pivr = fres[2];
        double* cov_out = P.cov ? P.cov + (size_t)b * NP * NP : nullptr;
        double ld = nan("");
        if (okc && pivr > 1e-6) {
            // well conditioned: kappa^2 * eps < ~2e-10, the normal-equation factor is accurate enough
            ld = covariance_from_chol<F>(c, cov_out);
        } else if (isfinite(cost)) {
            // ill conditioned: one more pass over the rows, Householder QR of J.diag(dsc)
            eval_full<F, 1>(c, c.p, nullptr, nullptr);
            ++njev;
            ld = covariance_from_qr<F>(c, cov_out);
        } else if (cov_out && act) {
            for (int j = 0; j < NP; ++j) cov_out[lane * NP + j] = nan("");
        }
        if (act) P.x_out[(size_t)b * NP + lane] = c.p[lane];
        if (lane == 0) {
            P.chi2[b] = 2.0 * cost;
            P.nit[b] = nfev;
            P.status[b] = status;
            if (P.logdet) P.logdet[b] = ld;
        }
        tot_nfev += nfev; tot_njev += njev; tot_nfac += nfac;
        __syncwarp();
    }
    if (lane == 0 && P.stats) {
        atomicAdd(&P.stats[0], tot_nfev);
        atomicAdd(&P.stats[1], tot_njev);
        atomicAdd(&P.stats[2], tot_nfac);
    }
}

// residual + Jacobian at given parameter vectors (test hook for the chiv parity of
// reference src/lsqfit/_utilities.pyx:65-94); P.p0 holds the B parameter vectors.
template <class F>
__global__ void __launch_bounds__(FitLayout<F>::MAX_WARPS * 32, 1) resjac_kernel(const __grid_constant__ FitParams P) {
    typedef FitLayout<F> Lay;
    constexpr int NP = Lay::NP;
    extern __shared__ double smem[];
    WarpCtx<F> c(P);
    setup_ctx<F>(c, smem, P);
    const int lane = c.lane;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int b = warp_global; b < P.B; b += nwarps) {
        c.mean = P.mean + (size_t)b * P.mean_stride;
        const double* p0 = P.p0 + (size_t)b * P.p0_stride;
        if (lane < NP) c.p[lane] = p0[lane];
        __syncwarp();
        const double cost = eval_full<F>(c, c.p, P.f_out ? P.f_out + (size_t)b * P.nchiv : nullptr,
                                         P.J_out ? P.J_out + (size_t)b * P.nchiv * NP : nullptr);
        if (lane == 0 && P.chi2) P.chi2[b] = 2.0 * cost;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
struct LaunchInfo {
    int grid, block;
    size_t smem;
};

template <class F>
inline cudaError_t plan_launch(FitParams& P, int sm_count, size_t smem_budget, LaunchInfo& li) {
    typedef FitLayout<F> Lay;
    const size_t wt_bytes = ((size_t)(P.wt_total + 1) & ~(size_t)1) * sizeof(double);
    const size_t per_warp = (size_t)Lay::per_warp_doubles(P.rb) * sizeof(double);
    P.wt_in_smem = (P.wt_total > 0 && wt_bytes + 4 * per_warp <= smem_budget) ? 1 : 0;
    const size_t avail = smem_budget - (P.wt_in_smem ? wt_bytes : 0);
    int warps = (int)(avail / per_warp);
    if (warps < 1) return cudaErrorInvalidConfiguration;
    if (warps > Lay::MAX_WARPS) warps = Lay::MAX_WARPS;
    P.warps = warps;
    li.block = warps * 32;
    li.smem = (P.wt_in_smem ? wt_bytes : 0) + warps * per_warp;
    int grid = (P.B + warps - 1) / warps;
    if (grid > sm_count) grid = sm_count;
    if (grid < 1) grid = 1;
    li.grid = grid;
    return cudaSuccess;
}

template <class F>
cudaError_t launch_fit(FitParams P, int sm_count, size_t smem_budget, cudaStream_t stream) {
    LaunchInfo li;
    cudaError_t e = plan_launch<F>(P, sm_count, smem_budget, li);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fit_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)li.smem);
    if (e != cudaSuccess) return e;
    fit_kernel<F><<<li.grid, li.block, li.smem, stream>>>(P);
    return cudaGetLastError();
}

template <class F>
cudaError_t launch_resjac(FitParams P, int sm_count, size_t smem_budget, cudaStream_t stream) {
    LaunchInfo li;
    cudaError_t e = plan_launch<F>(P, sm_count, smem_budget, li);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(resjac_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)li.smem);
    if (e != cudaSuccess) return e;
    resjac_kernel<F><<<li.grid, li.block, li.smem, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace b200lm
