"""Synthetic workloads of the BASELINE.json configs.

All inputs are generated on the host with ``numpy.random.default_rng(seed)`` so that the
same arrays can be fed to the device engine and to the CPU oracle.
"""
import numpy as np


def correlator(K, ny=64, dt=0.125, rho=0.9, rel_err=1e-3):
    """Lattice of the C3/C4 configs: t_i = dt*i (i = 1..ny), truth a_k = 0.5, E_k = 0.5 k,
    priors a_k = 0.5(4), E_k = 0.5k(2), data covariance sigma_i sigma_j rho^|i-j| with
    sigma_i = rel_err * f_i."""
    t = dt * np.arange(1, ny + 1)
    a = np.full(K, 0.5)
    E = 0.5 * np.arange(1, K + 1)
    ptrue = np.concatenate([a, E])
    f = (a[None, :] * np.exp(-E[None, :] * t[:, None])).sum(axis=1)
    sig = rel_err * f
    idx = np.arange(ny)
    ycov = sig[:, None] * sig[None, :] * rho ** np.abs(idx[:, None] - idx[None, :])
    prior_mean = ptrue.copy()
    prior_sdev = np.concatenate([np.full(K, 0.4), np.full(K, 0.2)])
    return dict(K=K, ny=ny, np=2 * K, x=t[:, None], f=f, ycov=ycov, ptrue=ptrue,
                prior_mean=prior_mean, prior_sdev=prior_sdev)


def bootstrap_means(cfg, B, seed, cov=None, vary_prior=True):
    """mean_b = [f + L z_b ; prior_mean + sigma_prior * z'_b]   (both fluctuate, reference
    src/lsqfit/__init__.py:1619-1623).  ``cov`` = the (svd-corrected) data covariance."""
    rng = np.random.default_rng(seed)
    ny, npar = cfg["ny"], cfg["np"]
    C = cfg["ycov"] if cov is None else cov
    val, vec = np.linalg.eigh(C)
    L = vec * np.sqrt(np.clip(val, 0.0, None))
    z = rng.standard_normal((B, ny))
    means = np.empty((B, ny + npar))
    means[:, :ny] = cfg["f"][None, :] + z @ L.T
    if vary_prior:
        means[:, ny:] = cfg["prior_mean"][None, :] + cfg["prior_sdev"][None, :] * rng.standard_normal((B, npar))
    else:
        means[:, ny:] = cfg["prior_mean"][None, :]
    return means


def c3(B=10000, seed=12345):
    """Config 3: 8-exponential correlator, 64 time slices, correlated data, B bootstrap copies;
    p0 = prior mean for every copy, tol = (1e-8, 1e-10, 1e-10), maxit = 1000, svdcut = 1e-12."""
    cfg = correlator(8)
    cfg.update(B=B, seed=seed, svdcut=1e-12, tol=(1e-8, 1e-10, 1e-10), maxit=1000,
               p0=cfg["prior_mean"].copy(), name="C3 8-exp correlator bootstrap")
    return cfg


def c4(B=1000000, seed=777):
    """Config 4: 3-exponential correlator, simulated_fit_iter semantics: data means
    f(pexact) + L z, prior means fixed, whitening shared, p0 = pexact."""
    cfg = correlator(3)
    cfg.update(B=B, seed=seed, svdcut=1e-12, tol=(1e-8, 1e-10, 1e-10), maxit=1000,
               p0=cfg["ptrue"].copy(), name="C4 3-exp correlator simulated fits")
    return cfg


def eval_flops(ny, npar, K, nchiv=None, dense=True):
    """Algorithmic flop count of one evaluation:
    whitening apply + J^T J, J^T r + Cholesky and solves + functor."""
    nchiv = ny + npar if nchiv is None else nchiv
    first = 2.0 * ny * ny * (npar + 1) if dense else 2.0 * ny * (npar + 1)
    return (first + 2.0 * nchiv * npar ** 2 + 2.0 * nchiv * npar
            + npar ** 3 / 3.0 + 2.0 * npar ** 2 + ny * K * 6.0)


def c5(ny=5000, K=1000, Ns=None, seed=5000, corr_len=50.0, rel_err=1e-3):
    """Config 5: one large dense fit.  Multi-exponential with K terms on
    t_i = 8 i / ny, priors a_k = 0.5(4)/sqrt(K), E_k = 0.01(k+1) +- 0.004; the data covariance is the
    SAMPLE covariance of Ns = ny/2 draws (rank deficient => more than ny/2 modes clamped by the svd
    cut) from sigma_i sigma_j exp(-|i-j|/corr_len), sigma_i = rel_err |f_i|; svdcut = 1e-8; truth
    drawn from the prior, data mean = f(truth) + noise with that sample covariance, start = prior mean."""
    rng = np.random.default_rng(seed)
    Ns = ny // 2 if Ns is None else Ns
    t = 8.0 * np.arange(1, ny + 1) / ny
    prior_mean = np.concatenate([np.full(K, 0.5 / np.sqrt(K)), 0.01 * (np.arange(K) + 1.0)])
    prior_sdev = np.concatenate([np.full(K, 0.4 / np.sqrt(K)), np.full(K, 0.004)])
    ptrue = prior_mean + prior_sdev * rng.standard_normal(2 * K)
    f = (ptrue[None, :K] * np.exp(-ptrue[None, K:] * t[:, None])).sum(axis=1)
    sig = rel_err * np.abs(f)
    idx = np.arange(ny)
    base = sig[:, None] * sig[None, :] * np.exp(-np.abs(idx[:, None] - idx[None, :]) / corr_len)
    L = np.linalg.cholesky(base + 1e-14 * np.diag(sig ** 2))
    draws = rng.standard_normal((Ns, ny)) @ L.T
    draws -= draws.mean(axis=0, keepdims=True)
    ycov = draws.T @ draws / (Ns - 1)
    # noise of the mean drawn from the sample covariance itself (lies in the span of the draws)
    ymean = f + draws.T @ rng.standard_normal(Ns) / np.sqrt(Ns - 1)
    return dict(K=K, ny=ny, np=2 * K, x=t[:, None], t=t, f=f, ymean=ymean, ycov=ycov, ptrue=ptrue,
                prior_mean=prior_mean, prior_sdev=prior_sdev, p0=prior_mean.copy(), svdcut=1e-8,
                tol=(1e-8, 1e-10, 1e-10), maxit=1000, name="C5 large dense fit %dx%d" % (2 * K, ny))
