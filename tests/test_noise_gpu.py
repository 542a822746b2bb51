"""``noise=`` of nonlinear_fit / PDF (reference src/lsqfit/__init__.py:493-499, 535-536, 1895-1898: svd noise through
``gvar.PDF(..., noise=)``, prior noise as one sample of the prior).  The noise itself is random, so the checks are
statistical: the covariance of many device samples against the oracle's correction covariance (corrected - input
covariance, oracle/whiten.py), reproducibility per seed, and which means move for which flag.  Needs a GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback)")


def _problem(n=24, rho=0.97, seed=3):
    rng = np.random.default_rng(seed)
    sig = 0.05 * (1.0 + rng.random(n))
    idx = np.arange(n)
    cov = sig[:, None] * sig[None, :] * rho ** np.abs(idx[:, None] - idx[None, :])
    return cov


@pytest.mark.parametrize("svdcut,eps", [(0.05, None), (None, 0.02)], ids=["svdcut", "eps"])
def test_svd_noise_covariance_matches_oracle_correction(svdcut, eps):
    _need_gpu()
    from lsqfit_b200.whiten import PDF
    from oracle.whiten import PDF as OPDF
    cov = _problem()
    n = cov.shape[0]
    # three uncorrelated entries next to the block: the regulator does not touch them -> no noise there
    full = np.zeros((n + 3, n + 3))
    full[:n, :n] = cov
    full[n:, n:] = np.diag([0.1, 0.2, 0.3]) ** 2
    mean = np.linspace(1.0, 2.0, n + 3)
    pdf = PDF(mean, full, svdcut=svdcut, eps=eps)
    opdf = OPDF(mean, full, svdcut=svdcut, eps=eps)
    target = opdf.correction_cov
    assert np.max(np.abs(np.diag(target)[:n])) > 0.0          # the regulator did something
    ns = 40000
    z = pdf.svd_noise(ns, seed=11)
    assert z.shape == (ns, n + 3)
    assert np.all(z[:, n:] == 0.0)
    emp = z.T @ z / ns
    scale = np.sqrt(np.outer(np.diag(target), np.diag(target)))[:n, :n]
    # entries of an empirical covariance of ns samples scatter by ~ sqrt(2 / ns) = 0.7 % of the scale; bar: 5 sigma
    assert np.max(np.abs(emp[:n, :n] - target[:n, :n]) / scale) < 5.0 * np.sqrt(2.0 / ns)
    assert abs(np.trace(emp) / np.trace(target) - 1.0) < 0.02
    # a tiny svd cut modifies nothing: no noise at all
    quiet = PDF(mean, full, svdcut=1e-15)
    assert np.all(quiet.svd_noise(10, seed=1) == 0.0)
    # PDF(noise=True) adds exactly the first sample of the stream of its seed
    noisy = PDF(mean, full, svdcut=svdcut, eps=eps, noise=True, noise_seed=11)
    np.testing.assert_allclose(noisy.mean, mean + z[0], rtol=0, atol=1e-15)


def test_nonlinear_fit_noise_flags():
    _need_gpu()
    import lsqfit_b200 as lb
    rng = np.random.default_rng(5)
    t = 0.25 * np.arange(1, 25)
    ptrue = np.array([0.5, 0.4, 0.6, 1.3])
    f = ptrue[0] * np.exp(-ptrue[2] * t) + ptrue[1] * np.exp(-ptrue[3] * t)
    sig = 0.01 * f
    idx = np.arange(t.size)
    ycov = sig[:, None] * sig[None, :] * 0.98 ** np.abs(idx[:, None] - idx[None, :])
    y = f + np.linalg.cholesky(ycov) @ rng.standard_normal(t.size)
    prior = (np.array([0.5, 0.5, 0.5, 1.5]), np.array([0.5, 0.5, 0.3, 0.5]))
    kw = dict(data=(t[:, None], y, ycov), fcn="multiexp", prior=prior, svdcut=1e-3)
    base = lb.nonlinear_fit(**kw)
    assert base.svdn > 0 and base.noise == (False, False)
    a = lb.nonlinear_fit(noise=True, noise_seed=7, **kw)
    b = lb.nonlinear_fit(noise=True, noise_seed=7, **kw)
    c = lb.nonlinear_fit(noise=True, noise_seed=8, **kw)
    assert a.noise == (True, True)
    np.testing.assert_array_equal(a.y, b.y)
    np.testing.assert_array_equal(a.pmean, b.pmean)
    assert np.any(a.y != base.y) and np.any(a.prior[0] != base.prior[0])
    assert np.any(a.y != c.y) and np.any(a.prior[0] != c.prior[0])
    # the noise is of the size of the added uncertainty: the fit moves, but by O(1) standard deviations
    assert 0.0 < np.max(np.abs(a.pmean - base.pmean) / base.psdev) < 8.0
    # svd noise only: prior means untouched; prior noise only: data untouched
    s = lb.nonlinear_fit(noise=(True, False), noise_seed=7, **kw)
    np.testing.assert_array_equal(s.prior[0], base.prior[0])
    np.testing.assert_array_equal(s.y, a.y)
    p = lb.nonlinear_fit(noise=(False, True), noise_seed=7, **kw)
    np.testing.assert_array_equal(p.y, base.y)
    np.testing.assert_array_equal(p.prior[0], a.prior[0])
    # the same whitening either way (noise moves means, not covariances)
    assert a.svdn == base.svdn
    np.testing.assert_allclose(a.yp_pdf.cov, base.yp_pdf.cov, rtol=0, atol=0)
