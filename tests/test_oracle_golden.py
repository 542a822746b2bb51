"""Pin the CPU oracle against the reference's own golden outputs (CPU only).

Fixtures come from tests/golden/make_golden.py (generated from the reference
tree: examples/*.out, examples/nist.py assert strings, NIST certified values,
numeric pins in tests/test_lsqfit.py).
"""
import json
import os

import numpy as np
import pytest

from oracle import gvfmt
from oracle.fit import nonlinear_fit
from oracle.fitter import gammaQ, scipy_least_squares
from oracle.whiten import PDF, cov_blocks
from oracle import models as M


def _expected_list(s):
    return s[1:-1].replace(" +- ", "+-").split()


def test_nist_all_27(nist_problems):
    """examples/nist.py assert strings + examples/nist.out chi2/dof, Q, logGBF."""
    assert len(nist_problems) == 27
    for pr in nist_problems:
        fit = nonlinear_fit(pr["form"], np.array(pr["x"]), pr["y"], pr["ysdev"],
                            prior_mean=pr["prior_mean"], prior_cov=pr["prior_sdev"],
                            p0=pr["p0"], tol=pr["tol"])
        exp = _expected_list(pr["expected"])
        for m, s, e in zip(fit.pmean, fit.p_sdev, exp):
            assert gvfmt.agrees(m, s, e), (pr["name"], m, s, e)
        o = pr["out"]
        assert fit.dof == o["dof"]
        assert gvfmt.agrees_g(fit.chi2 / fit.dof, o["chi2_dof"], 2), pr["name"]
        assert gvfmt.agrees_g(fit.Q, o["Q"], 2), pr["name"]
        assert gvfmt.agrees_g(fit.logGBF, o["logGBF"], 5), pr["name"]
        # NIST certified values: priors are 200x wide, so p and sdev agree closely
        cert, csd = np.array(pr["certified"]), np.array(pr["certified_sdev"])
        if pr["name"] != "lanczos1":        # sigma_y ~ 1e-13: roundoff-dominated (nist.py:18-20)
            assert np.max(np.abs(fit.pmean - cert) / csd) < 1e-3, pr["name"]
        assert np.max(np.abs(fit.p_sdev / csd - 1)) < 1e-3, pr["name"]
        # fit.p (D C D^T) and fit.palt (fit.cov) must agree (check_roundoff, __init__.py:884-895)
        assert np.allclose(fit.p_sdev, fit.psdev, rtol=1e-5)


def test_nist_model_values(nist_problems):
    """oracle models == the reference's own fcn closures at start and certified points."""
    for pr in nist_problems:
        x = np.array(pr["x"])
        f = M.MODELS[pr["form"]]
        np.testing.assert_allclose(f(x, np.array(pr["p0"])), pr["f_p0"], rtol=1e-13, atol=0)
        np.testing.assert_allclose(f(x, np.array(pr["certified"])), pr["f_cert"], rtol=1e-13, atol=0)


def test_model_jacobians_fd():
    """Dual-number Jacobians vs central finite differences for every model."""
    rng = np.random.default_rng(1)
    for name in M.MODELS:
        npar = {"multiexp": 6, "multiexp_de": 6, "simple": 2, "offset_exp": 3, "poly": 5,
                "exp_poly": 4, "xerr_logistic": 4 + 7, "misra1a": 2, "chwirut": 3,
                "lanczos": 6, "gauss": 8, "danwood": 2, "misra1b": 2, "misra1c": 2,
                "misra1d": 2, "kirby2": 5, "hahn1": 7, "nelson": 3, "mgh17": 5,
                "roszman1": 4, "enso": 9, "mgh09": 4, "rat42": 3, "mgh10": 3,
                "eckerle4": 3, "rat43": 4, "bennett5": 3, "gather": 3,
                "multiexp_shared2": 6, "multiexp_shared3": 8, "spline_poly": 13}[name]
        ny = 7
        x = rng.uniform(0.5, 2.0, size=(ny, 2))
        if name == "simple":
            x[:, 1] = [0, 0, 0, 1, 0, 1, 0]
        if name == "gather":
            x[:, 0] = [0, 1, 2, 2, 0, 1, 0]
        if name.startswith("multiexp_shared"):
            x[:, 1] = [0, 1, 1, 0, 1, 0, int(name[-1]) - 1]
        p = rng.uniform(0.5, 1.5, size=npar)
        if name == "spline_poly":
            p[:4] = [0.6, 1.0, 1.4, 1.9]                  # knots in increasing order, data on both sides of them
        f, G = M.value_and_jacobian(name, x, p)
        for j in range(npar):
            h = 1e-6 * max(1.0, abs(p[j]))
            pp, pm = p.copy(), p.copy()
            pp[j] += h
            pm[j] -= h
            fd = (M.MODELS[name](x, pp) - M.MODELS[name](x, pm)) / (2 * h)
            np.testing.assert_allclose(G[:, j], fd, rtol=2e-7, atol=1e-9, err_msg=name)


def _simple_fit(ex):
    ny = len(ex["ymean"])
    ycov = np.zeros((ny, ny))
    i = 0
    for b in ex["ycov_blocks"]:
        b = np.array(b)
        ycov[i:i + len(b), i:i + len(b)] = b
        i += len(b)
    return nonlinear_fit(ex["model"], np.array(ex["x"]), ex["ymean"], ycov,
                         prior_mean=ex["prior_mean"], prior_cov=ex["prior_sdev"])


def test_simple(golden_examples):
    """examples/simple.out:2-6 and the error budget :26-31 (pins D of _getp)."""
    ex = golden_examples["simple"]
    fit = _simple_fit(ex)
    o = ex["out"]
    for m, s, e in zip(fit.pmean, fit.p_sdev, o["p"]):
        assert gvfmt.agrees(m, s, e)
    assert fit.dof == o["dof"] and fit.svdn == o["svdn"]
    assert gvfmt.agrees_g(fit.chi2 / fit.dof, o["chi2_dof"], 2)
    assert gvfmt.agrees_g(fit.Q, o["Q"], 2)
    assert gvfmt.agrees_g(fit.logGBF, o["logGBF"], 5)
    fm, fc = fit.fcn_values_cov()
    for m, s, e in zip(fm, np.sqrt(np.diag(fc)), o["fit"]):
        assert gvfmt.agrees(m, s, e)
    # error budget: partial sdev of outputs (a, b/a, b) from y and from prior
    D = fit.D                              # d p / d (y, prior)
    C = fit.yp_pdf.cov
    a, b = fit.pmean
    T = np.array([[1, 0], [-b / a ** 2, 1 / a], [0, 1]])    # outputs a, b/a, b
    vals = np.array([a, b / a, b])
    ny = fit.ny
    Dy, Dp = T @ D[:, :ny], T @ D[:, ny:]
    ey = np.sqrt(np.diag(Dy @ C[:ny, :ny] @ Dy.T)) / np.abs(vals) * 100
    ep = np.sqrt(np.diag(Dp @ C[ny:, ny:] @ Dp.T)) / np.abs(vals) * 100
    np.testing.assert_allclose(ey, o["budget"]["y"], atol=0.006)
    np.testing.assert_allclose(ep, o["budget"]["prior"], atol=0.006)
    np.testing.assert_allclose(np.sqrt(ey ** 2 + ep ** 2), o["budget"]["total"], atol=0.006)


def _yvsx(ex, nexp, p0=None):
    pm = np.concatenate([np.full(nexp, 0.5), np.arange(1, nexp + 1.)])
    return nonlinear_fit("multiexp", np.array(ex["x"])[:, None], ex["ymean"], np.array(ex["ycov"]),
                         prior_mean=pm, prior_cov=np.full(2 * nexp, 0.4), p0=p0)


def test_y_vs_x(golden_examples):
    """examples/y-vs-x.out: svdcut 1e-12 modifies exactly one mode; four fits."""
    ex = golden_examples["y-vs-x"]
    for nexp in (1, 2, 3, 4):
        fit = _yvsx(ex, nexp)
        o = ex["out"][str(nexp)]
        assert fit.svdn == o["svdn"] and fit.dof == o["dof"]
        assert gvfmt.agrees_g(fit.chi2 / fit.dof, o["chi2_dof"], 2)
        assert gvfmt.agrees_g(fit.Q, o["Q"], 2)
        assert gvfmt.agrees_g(fit.logGBF, o["logGBF"], 5)
        for m, s, e in zip(fit.pmean, fit.p_sdev, o["p"]):
            assert gvfmt.agrees(m, s, e), (nexp, m, s, e)
        if nexp == 3:
            a, E = fit.pmean[:3], fit.pmean[3:]
            C = fit.p_cov
            def ratio(i, j, off):
                g = np.zeros(6)
                g[off + i] = 1 / fit.pmean[off + j]
                g[off + j] = -fit.pmean[off + i] / fit.pmean[off + j] ** 2
                return fit.pmean[off + i] / fit.pmean[off + j], np.sqrt(g @ C @ g)
            r = ex["ratios"]
            assert gvfmt.agrees(*ratio(1, 0, 3), r["E1_E0"])
            assert gvfmt.agrees(*ratio(2, 0, 3), r["E2_E0"])
            assert gvfmt.agrees(*ratio(1, 0, 0), r["a1_a0"])
            assert gvfmt.agrees(*ratio(2, 0, 0), r["a2_a0"])


def test_p_corr(golden_examples):
    """examples/p-corr.out: correlated prior block (2x2) + diag data."""
    ex = golden_examples["p-corr"]
    y = np.array([gvfmt.parse(s)[:2] for s in ex["y"]])
    fit = nonlinear_fit(ex["model"], np.array(ex["x"])[:, None], y[:, 0], y[:, 1],
                        prior_mean=ex["prior_mean"], prior_cov=np.array(ex["prior_cov"]))
    o = ex["out"]
    assert fit.dof == o["dof"] and fit.svdn == o["svdn"]
    assert gvfmt.agrees_g(fit.chi2 / fit.dof, o["chi2_dof"], 2)
    assert gvfmt.agrees_g(fit.Q, o["Q"], 2)
    assert gvfmt.agrees_g(fit.logGBF, o["logGBF"], 5)
    for m, s, e in zip(fit.pmean, fit.p_sdev, o["p"]):
        assert gvfmt.agrees(m, s, e)
    C = fit.p_cov
    assert "%.4f" % (C[0, 1] / np.sqrt(C[0, 0] * C[1, 1])) == o["corr_p0_p1"]
    p = fit.pmean
    g = np.array([-p[1] / p[0] ** 2, 1 / p[0], 0, 0])
    assert gvfmt.agrees(p[1] / p[0], np.sqrt(g @ C @ g), o["p1_p0"])
    g = np.array([0, 0, -p[3] / p[2] ** 2, 1 / p[2]])
    assert gvfmt.agrees(p[3] / p[2], np.sqrt(g @ C @ g), o["p3_p2"])


def test_x_err(golden_examples):
    """examples/x-err.out: 19 parameters, fcn(p) without x."""
    ex = golden_examples["x-err"]
    y = np.array([gvfmt.parse(s)[:2] for s in ex["y"]])
    xp = np.array([gvfmt.parse(s)[:2] for s in ex["xprior"]])
    bp = np.array([gvfmt.parse(s)[:2] for s in ex["bprior"]])
    pm = np.concatenate([bp[:, 0], xp[:, 0]])
    ps = np.concatenate([bp[:, 1], xp[:, 1]])
    # the golden was printed by the reference's default fitter (gsl_multifit, Moré
    # scaling); scipy's unscaled trf walks to a different local minimum from the default
    # start, its scaled variant (x_scale='jac' == Moré) lands on the printed one.
    fit = nonlinear_fit(ex["model"], np.zeros((len(y), 1)), y[:, 0], y[:, 1],
                        prior_mean=pm, prior_cov=ps, x_scale="jac")
    o = ex["out"]
    assert fit.dof == o["dof"]
    assert gvfmt.agrees_g(fit.chi2 / fit.dof, o["chi2_dof"], 2)
    assert gvfmt.agrees_g(fit.Q, o["Q"], 2)
    assert gvfmt.agrees_g(fit.logGBF, o["logGBF"], 5)
    for m, s, e in zip(fit.pmean, fit.p_sdev, o["p"]):
        assert gvfmt.agrees(m, s, e), (m, s, e)


def test_gammaQ(golden_examples):
    """tests/test_lsqfit.py:1887-1901"""
    for a, x, gax, gxa in golden_examples["gammaQ"]:
        np.testing.assert_allclose(gax, gammaQ(a, x), rtol=0.01)
        np.testing.assert_allclose(gxa, gammaQ(x, a), rtol=0.01)


def test_fitter_stopping_codes():
    """tests/test_lsqfit.py:1754-1777 (scipy_least_squares called directly)."""
    nx = 3
    xans = np.arange(nx) + 1.

    def f(x):
        return (x - xans) ** 2 + (x - xans) ** 4

    ans = scipy_least_squares(x0=np.ones(nx), n=nx, f=f, tol=(1e-15, 1e-8, 1e-15), method='trf')
    np.testing.assert_allclose(ans.x, xans, rtol=1e-3)
    assert ans.stopping_criterion == 2
    ans = scipy_least_squares(x0=np.zeros(nx), n=nx, f=f, tol=(1e-8, 1e-15, 1e-15), method='lm')
    np.testing.assert_allclose(ans.x, xans, rtol=1e-3)
    assert ans.stopping_criterion == 1
    ans = scipy_least_squares(x0=np.zeros(nx), n=nx, f=f, tol=(1e-15, 1e-8, 1e-15), method='dogbox')
    assert ans.stopping_criterion == 2
    np.testing.assert_allclose(ans.x, xans, rtol=1e-3)


def test_fitters_two_point():
    """tests/test_lsqfit.py:1811-1833: str(fit.p) == '[0.904(98) 2.17(19)]' for every method."""
    for method in ("trf", "dogbox", "lm"):
        fit = nonlinear_fit(lambda x, p: p, None, [0.9, 2.2], [0.1, 0.2],
                            prior_mean=[1.0, 2.0], prior_cov=[0.5, 0.5], method=method)
        assert gvfmt.agrees(fit.pmean[0], fit.p_sdev[0], "0.904(98)")
        assert gvfmt.agrees(fit.pmean[1], fit.p_sdev[1], "2.17(19)")


def test_whiten_pins():
    """tests/test_lsqfit.py:960-962, 976-978 (1x1 weights), :1011-1017 (block inverse),
    :932-943 (logdet) and the svdcut known answers of :581-589."""
    pdf = PDF([1., 10., np.log(2.)], [2., 4., 2.], svdcut=0)
    np.testing.assert_array_equal(pdf.i_invwgts[0][0], [0, 1, 2])
    np.testing.assert_array_equal(pdf.i_invwgts[0][1], [0.5, 0.25, 0.5])
    assert pdf.nchiv == 3 and pdf.nmod == 0
    # correlated pair (0,2), index 1 alone  (test case 3, :1000-1017)
    one_var = 1e-6
    cov = np.diag([4.0 + one_var, 16.0, 16.0 + 4 * one_var])
    cov[0, 2] = cov[2, 0] = 2 * one_var
    pdf = PDF([1, 10, 2], cov, svdcut=0.0)
    np.testing.assert_array_equal(pdf.i_invwgts[0][0], [1])
    np.testing.assert_array_equal(pdf.i_invwgts[1][0], [0, 2])
    np.testing.assert_allclose(pdf.icov(), np.linalg.inv(cov), rtol=1e-10)
    np.testing.assert_allclose(pdf.logdet, np.log(np.linalg.det(cov)), rtol=1e-12)
    # wavg pins (:581-589): data [(a+b)/2, (a+c)/2, a], a,b,c = 1(1); var = 1/(1^T C^-1 1)
    cov = np.array([[0.5, 0.25, 0.5], [0.25, 0.5, 0.5], [0.5, 0.5, 1.0]])
    one = np.ones(3)
    pdf = PDF(one, cov, svdcut=1 - 1e-16)
    np.testing.assert_allclose(1 / (one @ pdf.icov() @ one), 0.4561552812808828, rtol=1e-7)
    pdf = PDF(one, cov, svdcut=1e-18)
    np.testing.assert_allclose(1 / (one @ pdf.icov() @ one), 1. / 3., rtol=1e-7)
    # doc/source/overview.rst:1569-1580: svdcut 1e-4 on [[1,1,0],[1,1,0],[0,0,1e-20]]
    pdf = PDF(np.zeros(3), np.array([[1., 1, 0], [1, 1, 0], [0, 0, 1e-20]]), svdcut=1e-4)
    np.testing.assert_allclose(pdf.cov, [[1.0001, 0.9999, 0], [0.9999, 1.0001, 0], [0, 0, 1e-20]],
                               rtol=1e-12)
    assert pdf.nmod == 1


def test_whiten_negative_svdcut_drops_modes():
    """tests/test_lsqfit.py:830-841 -- svdcut<0 removes modes, nchiv shrinks."""
    rng = np.random.default_rng(3)
    A = rng.normal(size=(6, 3))
    cov = A @ A.T + 1e-14 * np.eye(6)
    pdf = PDF(np.zeros(6), cov, svdcut=-1e-10)
    assert pdf.nchiv == 3 and pdf.nmod == 3
    blocks = cov_blocks(np.diag([1., 2, 3]))
    assert list(blocks[0]) == [0, 1, 2] and blocks[1] == []


def _wavg_cases():
    """Inputs of reference tests/test_lsqfit.py:581-596 as arrays: (a+b)/2, (a+c)/2, a with a, b, c = 1(1)."""
    C3 = np.array([[0.5, 0.25, 0.5], [0.25, 0.5, 0.5], [0.5, 0.5, 1.0]])
    return C3


def test_wavg_known_answers():
    """lsqfit.wavg known answers (reference tests/test_lsqfit.py:581-596): svd-cut variances
    0.4561552812808828 and 1/3, and the array average [2.09802, 6.09802] +- 0.995037."""
    from oracle.fit import wavg
    C3 = _wavg_cases()
    assert abs(wavg([1.0, 1.0, 1.0], C3, index=[0, 0, 0], svdcut=1 - 1e-16).cov[0, 0] - 0.4561552812808828) < 1e-12
    assert abs(wavg([1.0, 1.0, 1.0], C3, index=[0, 0, 0], svdcut=1e-18).cov[0, 0] - 1.0 / 3.0) < 1e-9
    assert abs(wavg([1.0, 1.0, 1.0], np.ones(3), index=[0, 0, 0]).cov[0, 0] - 1.0 / 3.0) < 1e-12
    # [gvar(2.1,1), 4+gvar(2.1,1)] and [gvar(1.9,10), 4+gvar(1.9,10)]: the two components of each estimate are
    # the same random variable (fully correlated within an estimate)
    cov = np.zeros((4, 4))
    cov[np.ix_([0, 1], [0, 1])] = 1.0
    cov[np.ix_([2, 3], [2, 3])] = 100.0
    f = wavg([[2.1, 6.1], [1.9, 5.9]], cov)
    np.testing.assert_allclose(f.pmean, [2.09802, 6.09802], rtol=1e-4)
    np.testing.assert_allclose(f.psdev, [0.995037, 0.995037], rtol=1e-4)


from parity_util import _y_noerr_problem, Y_NOERR_OUT      # noqa: E402  (shared with the GPU test)


def _chain_p0(pm, prev, n):
    """examples/y-noerr.py:27-45 passes ``p0 = fit.pmean`` of the previous (nexp - 1) fit; _unpack_p0
    (src/lsqfit/__init__.py:1945-1987) copies the overlapping entries of every key and takes the rest from the prior."""
    p0 = np.array(pm, dtype=float)
    if prev is not None:
        m = n - 1
        p0[:m] = prev[:m]
        p0[n:n + m] = prev[m:2 * m]
    return p0


Y_NOERR_ITNS = {1: 12, 2: 31, 3: 64, 4: 143}        # examples/y-noerr.out:23, 49, 77, 107 (gsl_multifit)


@pytest.mark.parametrize("fitter", ["scipy", "gsl"])
def test_y_noerr(fitter):
    """examples/y-noerr.out (svd cut modifies 2-3 modes of a data (+) prior covariance whose data and prior
    parts are correlated; tol = 1e-15; every fit starts from the previous fit's parameters).  The nexp = 5 fit of the
    example is not pinned: its printed logGBF (83.141) belongs to a fit that the GSL fitter left after 249
    iterations short of the minimum the first four digits of its parameters already agree with.  With the GSL
    restatement the iteration counts of the example are reproduced as well (11/29/64/141 vs 12/31/64/143)."""
    from oracle.gsl_lm import gsl_multifit
    from oracle.fitter import scipy_least_squares
    prev = None
    for n in (1, 2, 3, 4):
        x, ymod, cov, pm = _y_noerr_problem(n)
        fit = nonlinear_fit("multiexp", x[:, None], ymod, yp_cov=cov, prior_mean=pm, p0=_chain_p0(pm, prev, n),
                            svdcut=1e-12, tol=1e-15, fitter=gsl_multifit if fitter == "gsl" else scipy_least_squares)
        prev = fit.pmean
        chi2dof, dof, Q, logGBF, svdn, a_exp, E_exp = Y_NOERR_OUT[n]
        assert fit.error is None
        assert fit.dof == dof and fit.yp_pdf.nmod == svdn
        assert gvfmt.agrees_g(fit.chi2 / fit.dof, chi2dof, 2)
        assert gvfmt.agrees_g(fit.Q, Q, 2)
        assert abs(fit.logGBF - float(logGBF)) < 1.5e-3
        for m, s, e in zip(fit.pmean, fit.psdev, a_exp + E_exp):
            assert gvfmt.agrees(m, s, e, slack=1.01), (m, s, e)
        if fitter == "gsl":
            assert abs(fit.nit - Y_NOERR_ITNS[n]) <= max(1, 0.1 * Y_NOERR_ITNS[n]), (n, fit.nit)


def test_gsl_lm_iteration_counts(nist_problems):
    """PIN of oracle/gsl_lm.py (the restatement of GSL's trust/lm driver behind lsqfit.gsl_multifit): the iteration
    counts printed in examples/nist.out -- 27 fits made by the reference with this fitter at tol = 1e-10 -- are
    reproduced exactly on 19 problems (among them 303, 97, 93, 92 and 22 iterations) and to +-1 ... 3 on the rest
    (the last iterations of a fit converged to rounding level depend on the linear algebra's rounding)."""
    from oracle.gsl_lm import gsl_multifit
    exact = close = 0
    for pr in nist_problems:
        fit = nonlinear_fit(pr["form"], np.array(pr["x"]), pr["y"], pr["ysdev"], prior_mean=pr["prior_mean"],
                            prior_cov=pr["prior_sdev"], p0=pr["p0"], tol=pr["tol"], fitter=gsl_multifit)
        want = pr["out"]["nit"]
        assert fit.error is None and fit.stopping_criterion == 1, pr["name"]
        exact += int(fit.nit == want)
        close += int(abs(fit.nit - want) <= max(1, 0.1 * want))
        assert abs(fit.nit - want) <= 3, (pr["name"], fit.nit, want)
        o = pr["out"]
        assert gvfmt.agrees_g(fit.chi2 / fit.dof, o["chi2_dof"], 2), pr["name"]
        if pr["name"] != "lanczos1":
            assert gvfmt.agrees_g(fit.logGBF, o["logGBF"], 5), pr["name"]
    assert exact >= 18 and close >= 23, (exact, close)


def _spline_problem():
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spline.json")))
    m = np.concatenate([g["param"][k][0] * np.array(g["param"][k][1]) for k in "ABC"])
    am = np.concatenate([np.array(g["param"][k][1]) for k in "ABC"])
    return g, np.stack([m, am], axis=1)


@pytest.mark.parametrize("fitter", ["scipy", "gsl"])
def test_spline_golden(fitter):
    """examples/spline.out: 12 correlated points, 13 parameters of which 8 are the knots of a gvar.cspline.CSpline.
    PINS the restatement of gvar's default spline (Steffen's monotonic cubic, oracle/models.py: steffen_spline):
    chi2/dof, Q, logGBF = 9.2202 and all printed parameters; with the GSL restatement also the 9 iterations."""
    from oracle.gsl_lm import gsl_multifit
    from oracle.fitter import scipy_least_squares
    g, x = _spline_problem()
    fit = nonlinear_fit("spline_poly", x, g["ymean"], np.array(g["ycov"]), prior_mean=g["prior_mean"], prior_cov=g["prior_sdev"],
                        fitter=gsl_multifit if fitter == "gsl" else scipy_least_squares)
    o = g["out"]
    assert fit.dof == o["dof"] and fit.error is None
    assert gvfmt.agrees_g(fit.chi2 / fit.dof, o["chi2_dof"], 2)
    assert gvfmt.agrees_g(fit.Q, o["Q"], 2)
    assert gvfmt.agrees_g(fit.logGBF, o["logGBF"], 5)
    for mu, sd, e in zip(fit.pmean, fit.p_sdev, o["params"]):
        assert gvfmt.agrees(mu, sd, e, slack=1.01), (mu, sd, e)
    if fitter == "gsl":
        assert abs(fit.nit - o["nit"]) <= 1, fit.nit


def test_oracle_svd_noise_has_the_correction_covariance():
    """``noise=True`` of the whitening (gvar.PDF as called at src/lsqfit/__init__.py:1895-1898): the sample added to the
    means has the covariance the svd cut added (corrected - input), nothing else moves, and no cut means no noise."""
    from oracle.whiten import PDF
    rng = np.random.default_rng(2)
    n = 12
    sig = 0.1 * (1.0 + rng.random(n))
    idx = np.arange(n)
    cov = sig[:, None] * sig[None, :] * 0.95 ** np.abs(idx[:, None] - idx[None, :])
    full = np.zeros((n + 2, n + 2))
    full[:n, :n] = cov
    full[n:, n:] = np.diag([0.3, 0.4]) ** 2
    mean = np.linspace(0.0, 1.0, n + 2)
    pdf = PDF(mean, full, svdcut=0.1)
    assert pdf.nmod > 0 and np.all(np.linalg.eigvalsh(pdf.correction_cov) > -1e-15)
    z = pdf.noise_samples(60000, np.random.default_rng(5))
    assert np.all(np.abs(z[:, n:]) < 1e-12)
    emp = z.T @ z / z.shape[0]
    scale = np.sqrt(np.outer(np.diag(pdf.correction_cov)[:n], np.diag(pdf.correction_cov)[:n]))
    assert np.max(np.abs(emp[:n, :n] - pdf.correction_cov[:n, :n]) / scale) < 5.0 * np.sqrt(2.0 / z.shape[0])
    noisy = PDF(mean, full, svdcut=0.1, noise=True, rng=np.random.default_rng(5))
    assert np.any(noisy.mean[:n] != mean[:n]) and np.all(noisy.mean[n:] == mean[n:])
    np.testing.assert_array_equal(noisy.cov, pdf.cov)
    quiet = PDF(mean, full, svdcut=1e-15, noise=True, rng=np.random.default_rng(5))
    np.testing.assert_allclose(quiet.mean, mean, rtol=0, atol=1e-12)
