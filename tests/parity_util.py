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
