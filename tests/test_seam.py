"""The reference's fitter seam, driven end to end.

gvar cannot be installed in this image, so the real ``lsqfit`` cannot be imported.  ``tests/lsqfit_double.py``
reproduces the exact call protocol around the seam (src/lsqfit/__init__.py:562-573, 657-682, 1391-1469,
1548-1642, 1997-2042; src/lsqfit/_extras.py:1164-1212, 1540-1586, 1816-1829); ``lsqfit_b200.register`` is installed
into that module exactly as it would be into lsqfit, and ``fitter='b200_lm'`` then runs through ``nonlinear_fit``,
``bootstrapped_fit_iter``, ``simulated_fit_iter`` and ``MultiFitter`` -- none of them modified.
"""
import collections
import sys
import os

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lsqfit_double as L                                         # noqa: E402
from oracle import models as M                                    # noqa: E402
from oracle import dual as D                                      # noqa: E402


def _problem(K=2, ny=24, seed=3):
    from lsqfit_b200 import configs
    cfg = configs.correlator(K, ny=ny, dt=0.25, rel_err=1e-3)
    rng = np.random.default_rng(seed)
    val, vec = np.linalg.eigh(cfg["ycov"])
    y = cfg["f"] + (vec * np.sqrt(np.clip(val, 0, None))) @ rng.standard_normal(ny)
    return cfg, y


# ---- CPU: the double itself and the composite mapping ---------------------------------------------------------
def test_double_reproduces_oracle_fit():
    """the double's protocol + scipy plugin gives what oracle.fit.nonlinear_fit gives (same arithmetic, two call paths)"""
    from oracle.fit import nonlinear_fit as ofit
    cfg, y = _problem()
    a = L.nonlinear_fit(data=(cfg["x"], y, cfg["ycov"]), prior=(cfg["prior_mean"], cfg["prior_sdev"]),
                        fcn=M.MODELS["multiexp"], tol=1e-10)
    b = ofit("multiexp", cfg["x"], y, cfg["ycov"], prior_mean=cfg["prior_mean"], prior_cov=cfg["prior_sdev"], tol=1e-10)
    np.testing.assert_allclose(a.pmean, b.pmean, rtol=0, atol=1e-12)
    np.testing.assert_allclose(a.chi2, b.chi2, rtol=1e-12)
    assert a.nit == b.nit
    with pytest.raises(ValueError, match="unknown fitter"):
        L.nonlinear_fit(data=(cfg["x"], y, cfg["ycov"]), prior=(cfg["prior_mean"], cfg["prior_sdev"]),
                        fcn=M.MODELS["multiexp"], fitter="no_such_fitter")


def _multi_problem(K=2, seed=5):
    from lsqfit_b200.multifit import SharedExpModel
    rng = np.random.default_rng(seed)
    t1, t2 = 0.25 * np.arange(1, 17), 0.25 * np.arange(2, 14)
    E = 0.5 * np.arange(1, K + 1)
    a1, a2 = np.full(K, 0.5), np.linspace(0.3, 0.8, K)
    f1 = (a1 * np.exp(-np.outer(t1, E))).sum(axis=1)
    f2 = (a2 * np.exp(-np.outer(t2, E))).sum(axis=1)
    data = collections.OrderedDict()
    for tag, f in (("G1", f1), ("G2", f2)):
        sig = 1e-3 * f
        i = np.arange(f.size)
        cov = sig[:, None] * sig[None, :] * 0.8 ** np.abs(i[:, None] - i[None, :])
        data[tag] = (f + np.linalg.cholesky(cov) @ rng.standard_normal(f.size), cov)
    # the prior dictionary's order is NOT the device's parameter order (E in the middle)
    pm = collections.OrderedDict([("a1", a1.copy()), ("E", E.copy()), ("a2", a2.copy())])
    ps = collections.OrderedDict([("a1", np.full(K, 0.4)), ("E", np.full(K, 0.2)), ("a2", np.full(K, 0.4))])
    models = [SharedExpModel("G1", t1, "a1", "E", exp=D.exp), SharedExpModel("G2", t2, "a2", "E", exp=D.exp)]
    return models, data, (pm, ps)


def test_composite_maps_multifitter_closure():
    """_multifitfcn(flatmodels) (a closure over models, _extras.py:1816-1829) -> one device functor, x rows and the
    permutation between lsqfit's flat parameter buffer and the device order; the functor's host evaluation equals
    the closure's on random parameters."""
    from lsqfit_b200.multifit import composite
    models, data, (pm, ps) = _multi_problem()
    mf = L.MultiFitter(models)
    fcn = mf.buildfitfcn()
    po = L._Flat(pm)
    yo = L._Flat(collections.OrderedDict((m.datatag, data[m.datatag][0]) for m in models))
    functor, x, pperm = composite(fcn, po, yo)
    assert functor.name == "multiexp_shared2" and x.shape == (28, 2)
    rng = np.random.default_rng(0)
    p = rng.uniform(0.2, 1.5, po.size)
    want = L.flatfcn_dd(p, False, fcn, po, yo)
    got = functor(x, p[pperm])
    np.testing.assert_allclose(got, want, rtol=1e-14)
    # a model list the device cannot express is declined, not mis-mapped
    models[1].E = "E2"
    assert composite(fcn, po, yo) is None


# ---- GPU: fitter='b200_lm' through the unmodified protocol ------------------------------------------------------
def _install():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback)")
    import lsqfit_b200 as lb
    lb.register(L)
    assert "b200_lm" in L.nonlinear_fit.FITTERS
    return lb


@pytest.mark.gpu
def test_b200_lm_through_nonlinear_fit_and_iterators():
    lb = _install()
    from lsqfit_b200.fitter import PLAN_CACHE
    cfg, y = _problem()
    data = (cfg["x"], y, cfg["ycov"])
    prior = (cfg["prior_mean"], cfg["prior_sdev"])
    tight = (1e-15, 0.0, 0.0)
    ref = L.nonlinear_fit(data=data, prior=prior, fcn=M.MODELS["multiexp"], fitter="scipy_least_squares", tol=tight,
                          x_scale="jac")
    PLAN_CACHE.clear()
    dev = L.nonlinear_fit(data=data, prior=prior, fcn=lb.Functor("multiexp"), fitter="b200_lm", tol=tight, polish=8)
    assert dev.error is None and dev.stopping_criterion in (1, 2, 3)
    np.testing.assert_allclose(dev.pmean, ref.pmean, rtol=0, atol=2e-6 * np.max(ref.psdev))
    np.testing.assert_allclose(dev.chi2, ref.chi2, rtol=1e-9)
    s = ref.psdev
    assert np.max(np.abs(dev.cov - ref.cov) / (s[:, None] * s[None, :])) < 1e-6
    assert dev.residuals.shape == ref.residuals.shape and dev.J.shape == ref.J.shape
    # chiv stays callable for the host (fit.chi2 checks, format()) and agrees with the reference's
    np.testing.assert_allclose(dev._chiv(dev.pmean), dev.residuals, rtol=0, atol=1e-9)
    # a Python fit function cannot run on the device: loud, no CPU fallback
    with pytest.raises(ValueError, match="device functor"):
        L.nonlinear_fit(data=data, prior=prior, fcn=M.MODELS["multiexp"], fitter="b200_lm")

    # bootstrapped_fit_iter, unchanged: one nonlinear_fit per copy; copies share ONE plan (no b200lm_create per copy)
    fd = L.nonlinear_fit(data=data, prior=prior, fcn=lb.Functor("multiexp"), fitter="b200_lm")
    fr = L.nonlinear_fit(data=data, prior=prior, fcn=M.MODELS["multiexp"], fitter="scipy_least_squares", x_scale="jac")
    m0, h0 = PLAN_CACHE.misses, PLAN_CACHE.hits
    n = 25
    for bd, br in zip(fd.bootstrapped_fit_iter(n=n, seed=11), fr.bootstrapped_fit_iter(n=n, seed=11)):
        assert bd.error is None
        np.testing.assert_allclose(bd.pmean, br.pmean, rtol=0, atol=3e-4 * np.max(br.psdev))
        np.testing.assert_allclose(bd.chi2, br.chi2, rtol=1e-7)
    assert PLAN_CACHE.misses == m0 and PLAN_CACHE.hits - h0 >= n
    # simulated_fit_iter, unchanged (the _yp_pdf mean-swap path)
    for sd, sr in zip(fd.simulated_fit_iter(n=10, seed=4), fr.simulated_fit_iter(n=10, seed=4)):
        assert sd.error is None
        np.testing.assert_allclose(sd.pmean, sr.pmean, rtol=0, atol=3e-4 * np.max(sr.psdev))
        np.testing.assert_allclose(sd.chi2, sr.chi2, rtol=1e-7)
    assert PLAN_CACHE.misses == m0


@pytest.mark.gpu
def test_b200_lm_through_multifitter():
    """MultiFitter.lsqfit and its bootstrap iterator with fitter='b200_lm': the closure built by buildfitfcn is
    recognised and mapped onto the composite shared-energy functor; results return in the prior dictionary's order."""
    _install()
    models, data, prior = _multi_problem()
    tight = (1e-15, 0.0, 0.0)
    ref = L.MultiFitter(models, fitter="scipy_least_squares", x_scale="jac", tol=tight)
    rfit = ref.lsqfit(data, prior)
    dev = L.MultiFitter(models, fitter="b200_lm", polish=8, tol=tight)
    dfit = dev.lsqfit(data, prior)
    assert dfit.error is None
    np.testing.assert_allclose(dfit.pmean, rfit.pmean, rtol=0, atol=2e-6 * np.max(rfit.psdev))
    np.testing.assert_allclose(dfit.chi2, rfit.chi2, rtol=1e-9)
    s = rfit.psdev
    assert np.max(np.abs(dfit.cov - rfit.cov) / (s[:, None] * s[None, :])) < 1e-6
    np.testing.assert_allclose(dfit.J, rfit.J, rtol=0, atol=1e-7 * np.max(np.abs(rfit.J)))
    ref.fitterargs["tol"] = dev.fitterargs["tol"] = 1e-10
    dev.fitterargs.pop("polish")
    ref.lsqfit(data, prior), dev.lsqfit(data, prior)
    for bd, br in zip(dev.bootstrapped_fit_iter(8, seed=2), ref.bootstrapped_fit_iter(8, seed=2)):
        assert bd.error is None
        np.testing.assert_allclose(bd.pmean, br.pmean, rtol=0, atol=3e-4 * np.max(br.psdev))
        np.testing.assert_allclose(bd.chi2, br.chi2, rtol=1e-7)
