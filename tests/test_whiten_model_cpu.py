"""CPU check of the device's large-block whitening ALGORITHM (tests/whiten_model.py) against the oracle."""
import numpy as np
import pytest

import whiten_model
from oracle.whiten import PDF as OPDF


def _sample_cov(n, ns, seed):
    rng = np.random.default_rng(seed)
    idx = np.arange(n)
    base = np.exp(-np.abs(idx[:, None] - idx[None, :]) / 50.0)
    L = np.linalg.cholesky(base + 1e-10 * np.eye(n))
    sig = rng.uniform(0.5, 2.0, size=n) * 1e-3
    samples = (L @ rng.standard_normal((n, ns))).T * sig[None, :]
    return np.cov(samples.T)


@pytest.mark.parametrize("n,ns,cut", [(200, 100, 1e-8), (160, 320, 1e-6), (130, 40, -1e-6)])
def test_pivoted_cholesky_onesided_jacobi_vs_oracle(n, ns, cut):
    """Pivoted Cholesky + one-sided block Jacobi + null-space completion reproduce the svd-cut whitening
    of the oracle (eigen-decomposition by LAPACK): nmod, logdet, inverse covariance, corrected covariance."""
    cov = _sample_cov(n, ns, seed=n)
    o = OPDF(np.zeros(n), cov, svdcut=cut)
    D = np.diag(cov) ** -0.5
    val, U, sweeps = whiten_model.eigen(cov * D[:, None] * D[None, :])
    assert sweeps <= 12
    assert np.max(np.abs(U.T @ U - np.eye(n))) < 1e-11
    vmin = abs(cut) * val[0]
    nmod = int(np.sum(val < vmin))
    assert nmod == o.nmod
    if cut > 0:
        used = np.maximum(val, vmin)
        W = (U / np.sqrt(used)).T * D[None, :]
        corr_cov = cov + (U * np.where(val < vmin, vmin - val, 0.0)) @ U.T / np.outer(D, D)
    else:
        kept = val >= vmin
        used = val[kept]
        W = (U[:, kept] / np.sqrt(used)).T * D[None, :]
        corr_cov = (U[:, kept] * used) @ U[:, kept].T / np.outer(D, D)
    logdet = np.sum(np.log(used)) - 2 * np.sum(np.log(D))
    np.testing.assert_allclose(logdet, o.logdet, rtol=1e-7, atol=1e-5)
    Wo = o.i_invwgts[1][1]
    assert W.shape == Wo.shape
    ic, ico = W.T @ W, Wo.T @ Wo
    sc = np.sqrt(np.diag(ico))
    tol = max(1e-8, 500 * whiten_model.EPS / abs(cut))
    assert np.max(np.abs(ic - ico) / np.outer(sc, sc)) < tol
    sc = np.sqrt(np.diag(o.cov))
    assert np.max(np.abs(corr_cov - o.cov) / np.outer(sc, sc)) < 1e-10
