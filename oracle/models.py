"""Numpy fit functions of the reference's examples (oracle only).

Each function is ``f(x, p)`` with ``x`` a float array of shape ``(ny, nx)`` (the
same row-major layout the device functors read) and ``p`` a flat vector that is
either a float array or a ``dual.Dual``.  Returns the ``ny`` model values.

Reference definitions being restated (paths relative to the reference tree):
  multiexp        examples/y-vs-x.py:58-61, examples/y-noerr.py:70-73
  multiexp_de     tests/test_lsqfit.py:1643-1649 (E = cumsum(dE))
  simple          examples/simple.py:43-48
  offset_exp      examples/uncorrelated.py:30-31
  poly            tests/test_lsqfit.py:878-880
  exp_poly        examples/empbayes.py:28-29
  xerr_logistic   examples/x-err.py:40-43
  NIST forms      examples/nist.py (line of each body given below)

TEST INFRASTRUCTURE ONLY -- never imported by lsqfit_b200.
"""
import numpy as np

from . import dual as D

exp, log, cos, sin, arctan = D.exp, D.log, D.cos, D.sin, D.arctan
pi = np.pi


def _x(x, c=0):
    x = np.asarray(x, dtype=float)
    return x[:, c] if x.ndim == 2 else x


def multiexp(x, p):
    t = _x(x)
    K = len(p) // 2
    ans = 0.0
    for k in range(K):
        ans = ans + p[k] * exp(-p[K + k] * t)
    return ans


def _multiexp_shared(M):
    """M data sets with shared energies (a simultaneous MultiFitter fit; src/lsqfit/_extras.py:1816-1829 evaluates the
    per-data-set models one after the other): x rows (t, m); p = [a^(0)(K), ..., a^(M-1)(K), E(K)]"""
    def f(x, p):
        x = np.asarray(x, dtype=float)
        t, mm = x[:, 0], x[:, 1].astype(int)
        K = len(p) // (M + 1)
        parts = []
        for i in range(len(t)):
            ans = 0.0
            for k in range(K):
                ans = ans + p[mm[i] * K + k] * exp(-p[M * K + k] * t[i])
            parts.append(ans)
        return D.stack(parts) if isinstance(parts[0], D.Dual) else np.array(parts)
    return f


def multiexp_de(x, p):
    t = _x(x)
    K = len(p) // 2
    E = D.cumsum(p[K:])
    ans = 0.0
    for k in range(K):
        ans = ans + p[k] * exp(-E[k] * t)
    return ans


def simple(x, p):
    """rows with x[:,1]==0: exp(a + x b); rows with x[:,1]==1: b/a."""
    x = np.asarray(x, dtype=float)
    t, kind = x[:, 0], x[:, 1]
    a, b = p[0], p[1]
    e = exp(a + b * t)
    r = b / a
    return e * (kind == 0) + r * (kind == 1)


def offset_exp(x, p):
    return p[0] + p[1] * exp(-p[2] * _x(x))


def poly(x, p):
    t = _x(x)
    ans = 0.0
    for n in range(len(p)):
        ans = ans + p[n] * t ** n
    return ans


def exp_poly(x, p):
    t = _x(x)
    s = 0.0
    for n in range(len(p)):
        s = s + p[n] * t ** n
    return exp(-s)


def xerr_logistic(x, p):
    """b0/(1+exp(b1 - b2 x_i))**(1/b3); the x_i are parameters p[4:]."""
    b0, b1, b2, b3 = p[0], p[1], p[2], p[3]
    xi = p[4:]
    return b0 / ((1.0 + exp(b1 - b2 * xi)) ** (1.0 / b3))


def gather(x, p):
    """f_i = p[idx_i]: the fit function lsqfit.wavg builds (src/lsqfit/_extras.py:499-507)."""
    return p[np.asarray(_x(x)).astype(int)]


# ---- NIST StRD forms (examples/nist.py:<line>) ------------------------------

def misra1a(x, b):          # :112 (also boxbod :1114)
    return b[0] * (1 - exp(-b[1] * _x(x)))


def chwirut(x, b):          # :145, :225
    t = _x(x)
    return exp(-b[0] * t) / (b[1] + b[2] * t)


def lanczos(x, b):          # :249, :795, :822
    t = _x(x)
    return b[0] * exp(-b[1] * t) + b[2] * exp(-b[3] * t) + b[4] * exp(-b[5] * t)


def gauss(x, b):            # :344, :441, :917
    t = _x(x)
    return (b[0] * exp(-b[1] * t) + b[2] * exp(-(t - b[3]) ** 2 / b[4] ** 2)
            + b[5] * exp(-(t - b[6]) ** 2 / b[7] ** 2))


def danwood(x, b):          # :462
    t = _x(x)
    return b[0] * exp(b[1] * np.log(t))     # x**b2 with x > 0


def misra1b(x, b):          # :483
    return b[0] * (1 - (1 + b[1] * _x(x) / 2) ** (-2))


def misra1c(x, b):          # :940
    return b[0] * (1 - (1 + 2 * b[1] * _x(x)) ** (-.5))


def misra1d(x, b):          # :961
    t = _x(x)
    return b[0] * b[1] * t * ((1 + b[1] * t) ** (-1))


def kirby2(x, b):           # :573
    t = _x(x)
    return (b[0] + b[1] * t + b[2] * t ** 2) / (1 + b[3] * t + b[4] * t ** 2)


def hahn1(x, b):            # :659 (also thurber :1093)
    t = _x(x)
    return ((b[0] + b[1] * t + b[2] * t ** 2 + b[3] * t ** 3)
            / (1 + b[4] * t + b[5] * t ** 2 + b[6] * t ** 3))


def nelson(x, b):           # :723  (fitted to log(y), :719)
    x = np.asarray(x, dtype=float)
    x1, x2 = x[:, 0], x[:, 1]
    return b[0] - b[1] * x1 * exp(-b[2] * x2)


def mgh17(x, b):            # :749
    t = _x(x)
    return b[0] + b[1] * exp(-t * b[3]) + b[2] * exp(-t * b[4])


def roszman1(x, b):         # :987
    t = _x(x)
    return b[0] - b[1] * t - arctan(b[2] / (t - b[3])) / pi


def enso(x, b):             # :1043
    t = _x(x)
    return (b[0] + b[1] * cos(2 * pi * t / 12) + b[2] * sin(2 * pi * t / 12)
            + b[4] * cos(2 * pi * t / b[3]) + b[5] * sin(2 * pi * t / b[3])
            + b[7] * cos(2 * pi * t / b[6]) + b[8] * sin(2 * pi * t / b[6]))


def mgh09(x, b):            # :1064 (also examples/p-corr.py:60-61)
    t = _x(x)
    return b[0] * (t ** 2 + t * b[1]) / (t ** 2 + t * b[2] + b[3])


def rat42(x, b):            # :1134
    return b[0] / (1 + exp(b[1] - b[2] * _x(x)))


def mgh10(x, b):            # :1156
    return b[0] * exp(b[1] / (_x(x) + b[2]))


def eckerle4(x, b):         # :1190
    t = _x(x)
    return (b[0] / b[1]) * exp(-0.5 * ((t - b[2]) / b[1]) ** 2)


def rat43(x, b):            # :1212
    return b[0] / ((1 + exp(b[1] - b[2] * _x(x))) ** (1 / b[3]))


def bennett5(x, b):         # :1291
    return b[0] * (b[1] + _x(x)) ** (-1 / b[2])



# ---- monotonic cubic spline through fitted knots + even powers (examples/spline.py:50-60) --------------------------
SPLINE_SHAPES = {13: (4, 5), 8: (4, 0), 12: (6, 0)}          # np -> (knots, polynomial coefficients) compiled on the device


def _sel_abs(a):
    return -a if D.value(a) < 0 else a


def _sel_min(a, b):
    return a if D.value(a) <= D.value(b) else b


def _sgn(a):
    v = float(D.value(a))
    return 1.0 if v > 0 else (-1.0 if v < 0 else 0.0)


def steffen_spline(xk, yk, xs):
    """gvar.cspline.CSpline(xk, yk)(xs) with gvar's defaults (third party, not vendored; gvar >= 13.1.5 pinned by the
    reference's setup): ``alg='steffen'`` -- Steffen's monotonic cubic Hermite spline, A&A 239 (1990) 443: interior slopes
    (sign(s_i-1) + sign(s_i)) min(|s_i-1|, |s_i|, |p_i| / 2), end slopes from the parabola through the three outermost
    knots limited to [0, 2 s] -- and ``extrap_order=3`` (the end cubics continue outside the knots).  Knots may be dual
    numbers.  PINNED by examples/spline.out: tests/test_oracle_golden.py::test_spline_golden."""
    n = len(xk)
    h = [xk[i + 1] - xk[i] for i in range(n - 1)]
    s = [(yk[i + 1] - yk[i]) / h[i] for i in range(n - 1)]
    yp = [None] * n
    for i in range(1, n - 1):
        pi = (s[i - 1] * h[i] + s[i] * h[i - 1]) / (h[i - 1] + h[i])
        yp[i] = _sel_min(_sel_min(_sel_abs(s[i - 1]), _sel_abs(s[i])), 0.5 * _sel_abs(pi)) * (_sgn(s[i - 1]) + _sgn(s[i]))

    def end(s0, s1, h0, h1):
        r = h0 / (h0 + h1)
        pe = s0 * (1.0 + r) - s1 * r
        if float(D.value(pe)) * float(D.value(s0)) <= 0:
            return pe * 0.0
        if abs(float(D.value(pe))) > 2 * abs(float(D.value(s0))):
            return s0 * 2.0
        return pe
    yp[0] = end(s[0], s[1], h[0], h[1])
    yp[-1] = end(s[-1], s[-2], h[-1], h[-2])
    out = []
    for xv in xs:
        i = 0
        for k in range(1, n - 1):
            if xv > float(D.value(xk[k])):
                i = k
        t = xv - xk[i]
        a = (yp[i] + yp[i + 1] - s[i] * 2.0) / (h[i] * h[i])
        b = (s[i] * 3.0 - yp[i] * 2.0 - yp[i + 1]) / h[i]
        out.append(((a * t + b) * t + yp[i]) * t + yk[i])
    return out


def spline_poly(x, p):
    """x rows (m, am); p = [mknot(NK), fknot(NK), c(NCF)]:  CSpline(mknot, fknot)(m) + sum_i c_i am^(2 + 2 i)"""
    x = np.asarray(x, dtype=float)
    nk, ncf = SPLINE_SHAPES[len(p)]
    vals = steffen_spline([p[i] for i in range(nk)], [p[nk + i] for i in range(nk)], x[:, 0])
    res = []
    for r, v in enumerate(vals):
        for i in range(ncf):
            v = v + p[2 * nk + i] * x[r, 1] ** (2 + 2 * i)
        res.append(v)
    return D.stack(res) if any(isinstance(v, D.Dual) for v in res) else np.array([float(v) for v in res])

MODELS = dict(
    multiexp=multiexp, multiexp_de=multiexp_de, simple=simple,
    multiexp_shared2=_multiexp_shared(2), multiexp_shared3=_multiexp_shared(3),
    offset_exp=offset_exp, poly=poly, exp_poly=exp_poly,
    xerr_logistic=xerr_logistic, gather=gather, spline_poly=spline_poly,
    misra1a=misra1a, chwirut=chwirut, lanczos=lanczos, gauss=gauss,
    danwood=danwood, misra1b=misra1b, misra1c=misra1c, misra1d=misra1d,
    kirby2=kirby2, hahn1=hahn1, nelson=nelson, mgh17=mgh17,
    roszman1=roszman1, enso=enso, mgh09=mgh09, rat42=rat42, mgh10=mgh10,
    eckerle4=eckerle4, rat43=rat43, bennett5=bennett5,
)

# NIST problem name -> model form
NIST_FORM = dict(
    misra1a="misra1a", boxbod="misra1a", chwirut1="chwirut", chwirut2="chwirut",
    lanczos1="lanczos", lanczos2="lanczos", lanczos3="lanczos",
    gauss1="gauss", gauss2="gauss", gauss3="gauss", danwood="danwood",
    misra1b="misra1b", misra1c="misra1c", misra1d="misra1d", kirby2="kirby2",
    hahn1="hahn1", thurber="hahn1", nelson="nelson", mgh17="mgh17",
    roszman1="roszman1", enso="enso", mgh09="mgh09", rat42="rat42",
    mgh10="mgh10", eckerle4="eckerle4", rat43="rat43", bennett5="bennett5",
)


def value_and_jacobian(name, x, p):
    """Model values f[ny] and G[ny, np] = df/dp via forward-mode AD."""
    p = np.asarray(p, dtype=float)
    out = MODELS[name](x, D.Dual.variables(p))
    return D.value(out), D.deriv(out, p.size)
