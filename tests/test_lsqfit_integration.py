"""Real-lsqfit integration: `fitter='b200_lm'` through the unmodified reference package.
Needs lsqfit + gvar (absent from the build image, where this test is skipped) and a GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fitter_plugin_inside_real_lsqfit():
    lsqfit = pytest.importorskip("lsqfit")
    gv = pytest.importorskip("gvar")
    import lsqfit_b200 as lb
    lb.register(lsqfit)
    assert "b200_lm" in lsqfit.nonlinear_fit.FITTERS
    x = np.array([1., 1.2, 1.4, 1.6, 1.8, 2., 2.2, 2.4, 2.6])
    ptrue = np.array([0.5, 0.4, 0.9, 1.9])
    f = lb.Functor("multiexp")
    y = gv.gvar(f(x, ptrue), 1e-3 * f(x, ptrue))
    prior = gv.gvar(["0.5(5)", "0.5(5)", "1.0(5)", "2.0(5)"])
    ref = lsqfit.nonlinear_fit(data=(x, y), prior=prior, fcn=f, fitter="scipy_least_squares", tol=1e-12)
    dev = lsqfit.nonlinear_fit(data=(x, y), prior=prior, fcn=f, fitter="b200_lm", tol=1e-12)
    assert dev.error is None
    np.testing.assert_allclose(gv.mean(dev.p), gv.mean(ref.p), atol=1e-6 * np.max(gv.sdev(ref.p)))
    np.testing.assert_allclose(dev.chi2, ref.chi2, rtol=1e-8)
    np.testing.assert_allclose(gv.evalcov(dev.p), gv.evalcov(ref.p), rtol=1e-6)
    n = 0
    for bs in dev.bootstrapped_fit_iter(n=3):          # the reference iterator, unchanged
        assert bs.error is None
        n += 1
    assert n == 3
    with pytest.raises(ValueError):                    # a plain Python fcn cannot run on the device
        lsqfit.nonlinear_fit(data=(x, y), prior=prior, fcn=lambda x, p: f(x, p), fitter="b200_lm")
