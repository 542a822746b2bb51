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
