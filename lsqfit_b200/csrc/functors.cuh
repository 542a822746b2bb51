// Registry of CUDA device functors: the fit functions of the reference's
// examples, evaluated per data row with forward-mode duals in registers.
//
// A functor F provides
//     static constexpr int NP;   number of fit parameters (compile time)
//     static constexpr int NX;   number of x columns the row reads
//     double value     (const double* xrow, int row, const double* p)
//     double value_grad(const double* xrow, int row, const double* p, double w, G out /*[NP]*/)
//            returns f and writes out[j] = w * df/dp_j straight into the row buffer; G is double* (row-major
//            buffer of the warp kernel) or a strided accessor (transposed buffer of the team kernel)
// x is row-major [ny][NX] in global memory; p points to the warp's parameter vector in shared
// memory.  value() and value_grad() return bit-identical f (same operation order), so that
// results do not depend on which of the two evaluated the final residuals.  Loops over
// exponentials / powers are deliberately NOT unrolled: the fit kernel is instruction-fetch
// bound and a double-precision exp is ~50 instructions.
//
// Reference definitions (paths relative to the reference tree):
//   multiexp       examples/y-vs-x.py:58-61, examples/y-noerr.py:70-73
//   multiexp_de    tests/test_lsqfit.py:1643-1649          (E = cumsum(dE))
//   simple         examples/simple.py:43-48
//   offset_exp     examples/uncorrelated.py:30-31
//   poly           tests/test_lsqfit.py:878-880
//   exp_poly       examples/empbayes.py:28-29
//   xerr_logistic  examples/x-err.py:40-43
//   NIST StRD      examples/nist.py (line numbers at each body)
#pragma once
#include "dual.cuh"

namespace b200lm {

// family ids (stable ABI; mirrored in lsqfit_b200/functors.py)
enum FunctorFamily : int {
    F_MULTIEXP = 0, F_MULTIEXP_DE = 1, F_SIMPLE = 2, F_OFFSET_EXP = 3, F_POLY = 4,
    F_EXP_POLY = 5, F_XERR_LOGISTIC = 6, F_GATHER = 7, F_MULTIEXP_SHARED2 = 8, F_MULTIEXP_SHARED3 = 9,
    F_MISRA1A = 10, F_CHWIRUT = 11, F_LANCZOS = 12, F_GAUSS = 13, F_DANWOOD = 14,
    F_MISRA1B = 15, F_MISRA1C = 16, F_MISRA1D = 17, F_KIRBY2 = 18, F_HAHN1 = 19,
    F_NELSON = 20, F_MGH17 = 21, F_ROSZMAN1 = 22, F_ENSO = 23, F_MGH09 = 24,
    F_RAT42 = 25, F_MGH10 = 26, F_ECKERLE4 = 27, F_RAT43 = 28, F_BENNETT5 = 29,
    F_SPLINE_POLY = 30,
};

// ---------------------------------------------------------------------------
// Generic wrapper: Body::eval<T>(xrow, p) written once, differentiated by Dual.
// ---------------------------------------------------------------------------
template <class Body, int NP_, int NX_ = 1>
struct ADFunctor {
    static constexpr int NP = NP_;
    static constexpr int NX = NX_;
    __device__ __forceinline__ static double value(const double* __restrict__ x, int, const double* p) {
        double q[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) q[j] = p[j];
        return Body::template eval<double>(x, q);
    }
    template <class G>
    __device__ __forceinline__ static double value_grad(const double* __restrict__ x, int,
                                                        const double* p, double w, G g) {
        Dual<NP> q[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) q[j] = Dual<NP>::variable(p[j], j);
        const Dual<NP> r = Body::template eval<Dual<NP>>(x, q);
#pragma unroll
        for (int j = 0; j < NP; ++j) g[j] = w * r.d[j];
        return r.v;
    }
};

// ---------------------------------------------------------------------------
// examples/uncorrelated.py:30-31:  b0 + b1 exp(-b2 x), hand-written gradient (1, e, -b1 x e).  The one-pass
// normal-equation kernel of the single-fit path (lm_rows.cuh) runs this model over millions of rows at the machine's
// ridge point: the dual-number form costs ~50 FP64 instructions per row (products with derivative slots that are
// identically 0 or 1 cannot be folded under IEEE rules), this one ~28.
// ---------------------------------------------------------------------------
struct OffsetExpModel {
    static constexpr int NP = 3;
    static constexpr int NX = 1;
    __device__ __forceinline__ static double value(const double* __restrict__ x, int, const double* p) {
        return p[0] + p[1] * ::exp(-(p[2] * x[0]));
    }
    template <class G>
    __device__ __forceinline__ static double value_grad(const double* __restrict__ x, int,
                                                        const double* p, double w, G g) {
        const double t = x[0];
        const double e = ::exp(-(p[2] * t));
        const double we = w * e;
        g[0] = w;
        g[1] = we;
        g[2] = -(p[1] * t) * we;
        return p[0] + p[1] * e;
    }
};

// ---------------------------------------------------------------------------
// Multi-exponential correlator: sum_k a_k exp(-E_k t); params [a_0..a_K-1, E_0..E_K-1]
// Hand-written gradient: d/da_k = e_k, d/dE_k = -a_k t e_k.
// ---------------------------------------------------------------------------
template <int K>
struct MultiExp {
    static constexpr int NP = 2 * K;
    static constexpr int NX = 1;
    __device__ __forceinline__ static double value(const double* __restrict__ x, int, const double* p) {
        const double t = x[0];
        double f = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) f = fma(p[k], ::exp(-p[K + k] * t), f);
        return f;
    }
    template <class G>
    __device__ __forceinline__ static double value_grad(const double* __restrict__ x, int,
                                                        const double* p, double w, G g) {
        const double t = x[0];
        const double wt = -w * t;
        double f = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const double e = ::exp(-p[K + k] * t);
            const double a = p[k];
            g[k] = w * e;
            g[K + k] = wt * (a * e);
            f = fma(a, e, f);
        }
        return f;
    }
    // Two lanes per row (team kernel): part 0 evaluates the exponentials [0, KH), part 1 [KH, K), each
    // writes its own columns of the gradient and returns its share of f.  The loop is unrolled: the
    // exponentials of one lane are independent, so their ~150-cycle latencies overlap.
    template <class G>
    __device__ __forceinline__ static double value_grad_part(const double* __restrict__ x, int,
                                                             const double* p, double w, G g, int part) {
        constexpr int KH = (K + 1) / 2;
        const double t = x[0];
        const double wt = -w * t;
        const int k0 = part * KH;
        double f = 0.0;
#pragma unroll
        for (int u = 0; u < KH; ++u) {
            const int k = k0 + u;
            if (k < K) {
                const double e = ::exp(-p[K + k] * t);
                const double a = p[k];
                g[k] = w * e;
                g[K + k] = wt * (a * e);
                f = fma(a, e, f);
            }
        }
        return f;
    }
};

// number of lanes this functor can spread one row's value_grad over (value_grad_part); primary template: lm_kernel.cuh
template <int K> struct SplitOf<MultiExp<K>> { static constexpr int value = K >= 2 ? 2 : 1; };

// M data sets that share their energies: set m is sum_k a^(m)_k exp(-E_k t); params [a^(0)(K), ..., a^(M-1)(K), E(K)],
// x row = (t, m).  The composite model of a simultaneous (lsqfit.MultiFitter) fit of several correlators, whose
// fit function is a Python closure over per-data-set models in the reference (src/lsqfit/_extras.py:1816-1829).
template <int K, int M>
struct MultiExpShared {
    static constexpr int NP = (M + 1) * K;
    static constexpr int NX = 2;
    __device__ __forceinline__ static double value(const double* __restrict__ x, int, const double* p) {
        const double t = x[0];
        const int m = (int)x[1];
        double f = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) f = fma(p[m * K + k], ::exp(-p[M * K + k] * t), f);
        return f;
    }
    template <class G>
    __device__ __forceinline__ static double value_grad(const double* __restrict__ x, int,
                                                        const double* p, double w, G g) {
        const double t = x[0];
        const int m = (int)x[1];
        const double wt = -w * t;
#pragma unroll
        for (int j = 0; j < M * K; ++j) g[j] = 0.0;
        double f = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const double e = ::exp(-p[M * K + k] * t);
            const double a = p[m * K + k];
            g[m * K + k] = w * e;
            g[M * K + k] = wt * (a * e);
            f = fma(a, e, f);
        }
        return f;
    }
};

// E_k = dE_0 + ... + dE_k ; params [a_0..a_K-1, dE_0..dE_K-1]
template <int K>
struct MultiExpDE {
    static constexpr int NP = 2 * K;
    static constexpr int NX = 1;
    __device__ __forceinline__ static double value(const double* __restrict__ x, int, const double* p) {
        const double t = x[0];
        double f = 0.0, E = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) { E += p[K + k]; f = fma(p[k], ::exp(-E * t), f); }
        return f;
    }
    template <class G>
    __device__ __forceinline__ static double value_grad(const double* __restrict__ x, int,
                                                        const double* p, double w, G g) {
        const double t = x[0];
        const double wt = -w * t;
        double f = 0.0, E = 0.0;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            E += p[K + k];
            const double e = ::exp(-E * t);
            const double a = p[k];
            g[k] = w * e;
            g[K + k] = a * e;                 // a_k e_k for now
            f = fma(a, e, f);
        }
        // df/d(dE_j) = -t sum_{k>=j} a_k e_k
        double tail = 0.0;
#pragma unroll 1
        for (int k = K - 1; k >= 0; --k) { tail += g[K + k]; g[K + k] = wt * tail; }
        return f;
    }
};

// polynomial sum_n p_n t^n
template <int NP_>
struct Poly {
    static constexpr int NP = NP_;
    static constexpr int NX = 1;
    __device__ __forceinline__ static double value(const double* __restrict__ x, int, const double* p) {
        const double t = x[0];
        double tn = 1.0, f = 0.0;
#pragma unroll 1
        for (int n = 0; n < NP; ++n) { f = fma(p[n], tn, f); tn *= t; }
        return f;
    }
    template <class G>
    __device__ __forceinline__ static double value_grad(const double* __restrict__ x, int,
                                                        const double* p, double w, G g) {
        const double t = x[0];
        double tn = 1.0, f = 0.0;
#pragma unroll 1
        for (int n = 0; n < NP; ++n) { g[n] = w * tn; f = fma(p[n], tn, f); tn *= t; }
        return f;
    }
};

// f_i = p[idx_i], idx_i = (int) x_i: the model behind lsqfit.wavg (reference src/lsqfit/_extras.py:499-507:
// every datum is an estimate of one component of p; ragged inputs simply use fewer indices)
template <int NP_>
struct Gather {
    static constexpr int NP = NP_;
    static constexpr int NX = 1;
    __device__ __forceinline__ static double value(const double* __restrict__ x, int, const double* p) {
        const int k = (int)x[0];
        double f = 0.0;
#pragma unroll
        for (int n = 0; n < NP; ++n) f = (n == k) ? p[n] : f;
        return f;
    }
    template <class G>
    __device__ __forceinline__ static double value_grad(const double* __restrict__ x, int,
                                                        const double* p, double w, G g) {
        const int k = (int)x[0];
        double f = 0.0;
#pragma unroll
        for (int n = 0; n < NP; ++n) { g[n] = (n == k) ? w : 0.0; f = (n == k) ? p[n] : f; }
        return f;
    }
};

// exp(-sum_n p_n t^n)
template <int NP_>
struct ExpPoly {
    static constexpr int NP = NP_;
    static constexpr int NX = 1;
    __device__ __forceinline__ static double value(const double* __restrict__ x, int r, const double* p) {
        return ::exp(-Poly<NP>::value(x, r, p));
    }
    template <class G>
    __device__ __forceinline__ static double value_grad(const double* __restrict__ x, int r,
                                                        const double* p, double w, G g) {
        const double f = ::exp(-Poly<NP>::value_grad(x, r, p, 1.0, g));
        const double wf = -w * f;
#pragma unroll 1
        for (int n = 0; n < NP; ++n) g[n] *= wf;
        return f;
    }
};

// errors-in-x logistic: b0/(1+exp(b1-b2 x_i))**(1/b3), the x_i are parameters p[4+i]
template <int NY>
struct XerrLogistic {
    static constexpr int NP = 4 + NY;
    static constexpr int NX = 1;
    template <class T>
    __device__ __forceinline__ static T body(const T& b0, const T& b1, const T& b2, const T& b3, const T& xi) {
        return b0 / pow(1.0 + exp(b1 - b2 * xi), 1.0 / b3);
    }
    __device__ __forceinline__ static double value(const double* __restrict__, int row, const double* p) {
        return body<double>(p[0], p[1], p[2], p[3], p[4 + row]);
    }
    template <class G>
    __device__ __forceinline__ static double value_grad(const double* __restrict__, int row,
                                                        const double* p, double w, G g) {
        typedef Dual<5> D5;
        const D5 r = body<D5>(D5::variable(p[0], 0), D5::variable(p[1], 1), D5::variable(p[2], 2),
                              D5::variable(p[3], 3), D5::variable(p[4 + row], 4));
#pragma unroll 1
        for (int j = 4; j < NP; ++j) g[j] = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] = w * r.d[j];
        g[4 + row] = w * r.d[4];
        return r.v;
    }
};

// ---------------------------------------------------------------------------
// Bodies differentiated by Dual (T = double or Dual<NP>)
// ---------------------------------------------------------------------------
#define B200LM_BODY(name) struct name { template <class T> __device__ __forceinline__ static T eval(const double* __restrict__ x, const T* b)

// examples/simple.py:43-48 ; x = (t, kind): kind 0 -> exp(a + t b), kind 1 -> b/a
B200LM_BODY(SimpleBody) {
    if (x[1] == 0.0) return exp(b[0] + b[1] * x[0]);
    return b[1] / b[0];
} };
// examples/uncorrelated.py:30-31
B200LM_BODY(OffsetExpBody) { return b[0] + b[1] * exp(-(b[2] * x[0])); } };
// nist.py:112 misra1a, :1114 boxbod
B200LM_BODY(Misra1aBody) { return b[0] * (1.0 - exp(-(b[1] * x[0]))); } };
// nist.py:145, :225 chwirut1/2
B200LM_BODY(ChwirutBody) { return exp(-(b[0] * x[0])) / (b[1] + b[2] * x[0]); } };
// nist.py:249, :795, :822 lanczos1/2/3
B200LM_BODY(LanczosBody) {
    return b[0] * exp(-(b[1] * x[0])) + b[2] * exp(-(b[3] * x[0])) + b[4] * exp(-(b[5] * x[0]));
} };
// nist.py:344, :441, :917 gauss1/2/3
B200LM_BODY(GaussBody) {
    const double t = x[0];
    return b[0] * exp(-(b[1] * t)) + b[2] * exp(-(sqr(t - b[3]) / sqr(b[4])))
         + b[5] * exp(-(sqr(t - b[6]) / sqr(b[7])));
} };
// nist.py:462 danwood: b1 * x**b2
B200LM_BODY(DanwoodBody) { return b[0] * exp(b[1] * ::log(x[0])); } };
// nist.py:483 misra1b
B200LM_BODY(Misra1bBody) { return b[0] * (1.0 - 1.0 / sqr(1.0 + b[1] * (x[0] * 0.5))); } };
// nist.py:940 misra1c
B200LM_BODY(Misra1cBody) { return b[0] * (1.0 - 1.0 / sqrt(1.0 + b[1] * (2.0 * x[0]))); } };
// nist.py:961 misra1d
B200LM_BODY(Misra1dBody) { return b[0] * b[1] * x[0] / (1.0 + b[1] * x[0]); } };
// nist.py:573 kirby2
B200LM_BODY(Kirby2Body) {
    const double t = x[0], t2 = t * t;
    return (b[0] + b[1] * t + b[2] * t2) / (1.0 + b[3] * t + b[4] * t2);
} };
// nist.py:659 hahn1, :1093 thurber
B200LM_BODY(Hahn1Body) {
    const double t = x[0], t2 = t * t, t3 = t2 * t;
    return (b[0] + b[1] * t + b[2] * t2 + b[3] * t3) / (1.0 + b[4] * t + b[5] * t2 + b[6] * t3);
} };
// nist.py:723 nelson (fitted to log y); x = (x1, x2)
B200LM_BODY(NelsonBody) { return b[0] - b[1] * x[0] * exp(-(b[2] * x[1])); } };
// nist.py:749 mgh17
B200LM_BODY(Mgh17Body) { return b[0] + b[1] * exp(-(b[3] * x[0])) + b[2] * exp(-(b[4] * x[0])); } };
// nist.py:987 roszman1
B200LM_BODY(Roszman1Body) {
    return b[0] - b[1] * x[0] - atan(b[2] / (x[0] - b[3])) * 0.31830988618379067154;
} };
// nist.py:1043 enso
B200LM_BODY(EnsoBody) {
    const double w = 6.283185307179586476925 * x[0];
    const T a4 = w / b[3], a7 = w / b[6];
    return b[0] + b[1] * ::cos(w / 12.0) + b[2] * ::sin(w / 12.0)
         + b[4] * cos(a4) + b[5] * sin(a4) + b[7] * cos(a7) + b[8] * sin(a7);
} };
// nist.py:1064 mgh09, examples/p-corr.py:60-61
B200LM_BODY(Mgh09Body) {
    const double t = x[0], t2 = t * t;
    return b[0] * (t2 + b[1] * t) / (t2 + b[2] * t + b[3]);
} };
// nist.py:1134 rat42
B200LM_BODY(Rat42Body) { return b[0] / (1.0 + exp(b[1] - b[2] * x[0])); } };
// nist.py:1156 mgh10
B200LM_BODY(Mgh10Body) { return b[0] * exp(b[1] / (x[0] + b[2])); } };
// nist.py:1190 eckerle4
B200LM_BODY(Eckerle4Body) { return (b[0] / b[1]) * exp(-0.5 * sqr((x[0] - b[2]) / b[1])); } };
// nist.py:1212 rat43
B200LM_BODY(Rat43Body) { return b[0] / pow(1.0 + exp(b[1] - b[2] * x[0]), 1.0 / b[3]); } };
// nist.py:1291 bennett5
B200LM_BODY(Bennett5Body) { return b[0] * pow(b[1] + x[0], -1.0 / b[2]); } };

#undef B200LM_BODY

// ---------------------------------------------------------------------------
// Monotonic cubic spline (Steffen 1990) through FITTED knots + even powers: the model of examples/spline.py:50-60,
//     f(m, am) = CSpline(mknot, fknot)(m) + sum_i c_i am^(2 + 2 i),     params [mknot(NK), fknot(NK), c(NCF)],  x = (m, am).
// gvar.cspline.CSpline (third party, not vendored) defaults to Steffen's algorithm with Steffen's own end slopes and
// continues the end cubics outside the knots (extrap_order = 3); restated in oracle/models.py and PINNED there by
// examples/spline.out (logGBF = 9.2202 and every printed parameter).  Knots are parameters, so slopes, interval
// search and the min/sign selections are differentiated through by the dual numbers (selections by value).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double valof(double a) { return a; }
template <int N> __device__ __forceinline__ double valof(const Dual<N>& a) { return a.v; }
template <class T> __device__ __forceinline__ T abs_sel(const T& a) { return valof(a) < 0.0 ? -a : a; }
template <class T> __device__ __forceinline__ T min_sel(const T& a, const T& b) { return valof(a) <= valof(b) ? a : b; }
__device__ __forceinline__ double sign_of(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0); }

template <int NK, int NCF>
struct SplinePolyBody {
    static_assert(NK >= 3, "Steffen's end slopes need three knots");
    // end slope from the parabola through the three outermost knots, limited as in Steffen's paper
    template <class T>
    __device__ __forceinline__ static T end_slope(const T& s0, const T& s1, const T& h0, const T& h1) {
        const T r = h0 / (h0 + h1);
        const T pe = s0 * (1.0 + r) - s1 * r;
        if (valof(pe) * valof(s0) <= 0.0) return pe * 0.0;
        if (fabs(valof(pe)) > 2.0 * fabs(valof(s0))) return s0 * 2.0;
        return pe;
    }
    template <class T>
    __device__ __forceinline__ static T eval(const double* __restrict__ x, const T* p) {
        T h[NK - 1], sl[NK - 1], yp[NK];
#pragma unroll
        for (int j = 0; j < NK - 1; ++j) { h[j] = p[j + 1] - p[j]; sl[j] = (p[NK + j + 1] - p[NK + j]) / h[j]; }
#pragma unroll
        for (int j = 1; j < NK - 1; ++j) {
            const T pj = (sl[j - 1] * h[j] + sl[j] * h[j - 1]) / (h[j - 1] + h[j]);
            const T m = min_sel(min_sel(abs_sel(sl[j - 1]), abs_sel(sl[j])), abs_sel(pj) * 0.5);
            yp[j] = m * (sign_of(valof(sl[j - 1])) + sign_of(valof(sl[j])));
        }
        yp[0] = end_slope(sl[0], sl[1], h[0], h[1]);
        yp[NK - 1] = end_slope(sl[NK - 2], sl[NK - 3], h[NK - 2], h[NK - 3]);
        const double xv = x[0];
        int iv = 0;
#pragma unroll
        for (int k = 1; k < NK - 1; ++k) iv = (xv > valof(p[k])) ? k : iv;
        T f = p[NK] * 0.0;
#pragma unroll
        for (int j = 0; j < NK - 1; ++j) {
            if (iv == j) {
                const T t = xv - p[j];
                const T a = (yp[j] + yp[j + 1] - sl[j] * 2.0) / (h[j] * h[j]);
                const T b = (sl[j] * 3.0 - yp[j] * 2.0 - yp[j + 1]) / h[j];
                f = ((a * t + b) * t + yp[j]) * t + p[NK + j];
            }
        }
        const double am2 = x[1] * x[1];
        double pw = am2;
#pragma unroll
        for (int i = 0; i < NCF; ++i) { f = f + p[2 * NK + i] * pw; pw *= am2; }
        return f;
    }
};
template <int NK, int NCF> using SplinePoly = ADFunctor<SplinePolyBody<NK, NCF>, 2 * NK + NCF, 2>;

}  // namespace b200lm
