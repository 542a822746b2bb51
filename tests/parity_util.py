"""Helpers shared by the GPU parity tests (test infrastructure)."""
import numpy as np

from oracle import dual as D

# Both solvers accept a step only if the cost decreases, and cost differences below
# eps*cost cannot be resolved in fp64.  The reference's own solver therefore stalls up to
# ~sqrt(eps*chi2) standard deviations from the stationary point.  TIGHT switches the ftol
# test off (it would stop even earlier) and lets xtol decide.
TIGHT = (1e-15, 0.0, 0.0)


def exact_minimum(ofit, iters=30):
    """Stationary point of the oracle's chi2 next to ofit.pmean, by undamped Gauss-Newton
    steps in numpy (QR least squares), to the gradient's rounding level.  Returns
    (x, f, J, cov) with cov computed like the reference (src/lsqfit/_scipy.py:170-175)."""
    chiv = ofit._chiv
    x = np.array(ofit.pmean, dtype=float)
    sd = ofit.psdev
    for _ in range(iters):
        f = np.asarray(chiv(x))
        J = D.deriv(chiv(D.Dual.variables(x)), x.size)
        dx = np.linalg.lstsq(J, -f, rcond=None)[0]
        x = x + dx
        if np.max(np.abs(dx) / sd) < 1e-14:
            break
    f = np.asarray(chiv(x))
    J = D.deriv(chiv(D.Dual.variables(x)), x.size)
    _, s, VT = np.linalg.svd(J, full_matrices=False)
    cov = (VT.T / s ** 2) @ VT
    return x, f, J, cov


def rel_cov(cov, ref):
    s = np.sqrt(np.diag(ref))
    return np.max(np.abs(cov - ref) / (s[:, None] * s[None, :]))


def _y_noerr_problem(n, NT=100):
    """examples/y-noerr.py:24-46, 63-89 in array form: exact data, a prior of 100 exponentials of which the
    last 100-n are marginalised into the data (ymod = y - fcn(x, ymod_prior)).  Linear error propagation from
    the primary variables (a_k = 0.5(5), dE_k = 1.0(1), E = cumsum(dE)) gives the joint covariance of
    [ymod, a_<n, E_<n] -- data and prior are correlated through the shared dE's."""
    x = np.array([1., 1.2, 1.4, 1.6, 1.8, 2., 2.2, 2.4, 2.6])
    y = np.array([0.2740471001620033, 0.2056894154005132, 0.158389402324004, 0.1241967645280511,
                  0.0986901274726867, 0.0792134506060024, 0.0640743982173861, 0.052143504367789,
                  0.0426383022456816])
    a0, E0 = np.full(NT, 0.5), np.cumsum(np.full(NT, 1.0))
    ex = np.exp(-np.outer(x, E0))
    ymod = y - (a0[n:] * ex[:, n:]).sum(axis=1)
    J = np.zeros((9 + 2 * n, 2 * NT))
    J[:9, n:NT] = -ex[:, n:]
    dydE = a0[n:] * x[:, None] * ex[:, n:]
    for j in range(NT):
        ks = np.arange(max(j, n), NT)
        J[:9, NT + j] = dydE[:, ks - n].sum(axis=1)
    for k in range(n):
        J[9 + k, k] = 1.0
        J[9 + n + k, NT:NT + k + 1] = 1.0
    var = np.concatenate([np.full(NT, 0.25), np.full(NT, 0.01)])
    return x, ymod, (J * var) @ J.T, np.concatenate([a0[:n], E0[:n]])


Y_NOERR_OUT = {      # examples/y-noerr.out: chi2/dof, dof, Q, logGBF, svdcut/n, parameters
    1: ("0.19", 9, "0.99", "79.803", 2, ["0.4067(32)"], ["0.9030(16)"]),
    2: ("0.19", 9, "1", "81.799", 2, ["0.4015(23)", "0.435(24)"], ["0.9007(11)", "1.830(28)"]),
    3: ("0.2", 9, "0.99", "83.077", 3, ["0.4011(18)", "0.426(28)", "0.468(56)"],
        ["0.90045(77)", "1.822(27)", "2.84(12)"]),
    4: ("0.21", 9, "0.99", "83.212", 3, ["0.4009(10)", "0.424(22)", "0.469(61)", "0.426(94)"],
        ["0.90036(44)", "1.819(19)", "2.83(11)", "3.83(15)"]),
}


# ------------------------------------------------------------------------------------------------
# Oracle fits in a process pool (spawned: the parent may hold a CUDA context).  Every worker rebuilds the
# problem from the same seeded recipe, so inputs are identical to the device's.
# ------------------------------------------------------------------------------------------------
_W = {}


def _load_configs():
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_b200lm_configs", os.path.join(root, "lsqfit_b200", "configs.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def correlator_problem(K, svdcut=1e-12):
    """(cfg, oracle PDF) of the C3 / C4 lattice with K exponentials (lsqfit_b200/configs.py::correlator)."""
    from oracle.whiten import PDF as OPDF
    cfg = _load_configs().correlator(K)
    ny, npar = cfg["ny"], cfg["np"]
    N = ny + npar
    full = np.zeros((N, N))
    full[:ny, :ny] = cfg["ycov"]
    full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
    pdf = OPDF(np.concatenate([cfg["f"], cfg["prior_mean"]]), full, svdcut=svdcut)
    return cfg, pdf


def _corr_init(K):
    import os
    import warnings
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    warnings.simplefilter("ignore")
    _W["cfg"], _W["pdf"] = correlator_problem(K)


def _corr_job(arg):
    """One oracle fit of a correlator copy: arg = (mean, p0, tol, maxit, refine)."""
    from oracle.fit import nonlinear_fit
    mean, p0, tol, maxit, refine = arg
    cfg, pdf = _W["cfg"], _W["pdf"]
    ny = cfg["ny"]
    fo = nonlinear_fit("multiexp", cfg["x"], mean[:ny], prior_mean=mean[ny:], _yp_pdf=pdf, p0=p0, tol=tol,
                       maxit=maxit, x_scale="jac")
    out = dict(x=fo.pmean, cov=fo.cov, chi2=fo.chi2, nit=fo.nit, crit=fo.stopping_criterion)
    if refine and fo.stopping_criterion != 0:
        xe, fe, Je, cove = exact_minimum(fo, iters=3000)
        # undamped Gauss-Newton only converges from a point that is already close: keep the refined point when it
        # is finite and within 1e-3 sdev of where the reference's solver stopped, otherwise the copy is skipped
        gap = np.max(np.abs(fo.pmean - xe) / np.sqrt(np.abs(np.diag(cove)))) if np.all(np.isfinite(xe)) else np.inf
        if np.isfinite(gap) and gap <= 1e-3:
            # how converged the refinement itself is: one more Gauss-Newton step, in sdev
            last = np.max(np.abs(np.linalg.lstsq(Je, -fe, rcond=None)[0]) / np.sqrt(np.diag(cove)))
            out.update(xe=xe, cove=cove, chi2e=float(fe @ fe), logdete=float(np.linalg.slogdet(Je.T @ Je)[1]),
                       refine_last_step=float(last))
    return out


def oracle_correlator_fits(K, means, p0, tol, maxit=1000, refine=False, nproc=None):
    import multiprocessing as mp
    import os
    nproc = nproc or min(os.cpu_count() or 1, 32)
    jobs = [(m, p0, tol, maxit, refine) for m in means]
    with mp.get_context("spawn").Pool(nproc, initializer=_corr_init, initargs=(K,)) as pool:
        return pool.map(_corr_job, jobs, chunksize=max(1, len(jobs) // (nproc * 8)))


def _nist_init():
    import json
    import os
    import warnings
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    warnings.simplefilter("ignore")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "nist.json")) as f:
        _W["nist"] = json.load(f)["problems"]


def _nist_job(arg):
    from oracle.fit import nonlinear_fit
    k, p0, tol, maxit = arg
    pr = _W["nist"][k]
    fo = nonlinear_fit(pr["form"], np.array(pr["x"]), pr["y"], pr["ysdev"], prior_mean=pr["prior_mean"],
                       prior_cov=pr["prior_sdev"], p0=p0, tol=tol, maxit=maxit, x_scale="jac")
    return dict(x=fo.pmean, chi2=fo.chi2, nit=fo.nit, crit=fo.stopping_criterion, sd=fo.psdev)


def oracle_nist_fits(jobs, nproc=None):
    """jobs = [(problem index, p0, tol, maxit)]"""
    import multiprocessing as mp
    import os
    nproc = nproc or min(os.cpu_count() or 1, 32)
    with mp.get_context("spawn").Pool(nproc, initializer=_nist_init) as pool:
        return pool.map(_nist_job, jobs, chunksize=max(1, len(jobs) // (nproc * 8)))
