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
