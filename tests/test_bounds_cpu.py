"""Host side of the bounded single-fit path (lsqfit_b200/dense.py): the restated pieces of scipy's trf with bounds
(Coleman-Li scaling vector, step to the bound, trust-region intersection, 1-d quadratic minimiser, strict feasibility)
against scipy's own functions (scipy.optimize._lsq.common -- the third-party code behind lsqfit.scipy_least_squares
with ``bounds=``, reference src/lsqfit/_scipy.py:77) on random inputs.  No GPU needed."""
import numpy as np
import pytest


def test_bound_helpers_match_scipy():
    sc = pytest.importorskip("scipy.optimize._lsq.common")
    from lsqfit_b200 import dense as D
    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.integers(1, 9))
        lb = rng.uniform(-2, 0, n)
        ub = lb + rng.uniform(0.1, 3, n)
        lb[rng.random(n) < 0.2] = -np.inf
        ub[rng.random(n) < 0.2] = np.inf
        x = np.clip(rng.uniform(-1.5, 1.5, n), np.where(np.isfinite(lb), lb, -9) + 1e-3, np.where(np.isfinite(ub), ub, 9) - 1e-3)
        if trial % 5 == 0:                       # some points on / next to a bound
            k = int(rng.integers(0, n))
            if np.isfinite(ub[k]):
                x[k] = ub[k] - (1e-12 if trial % 10 == 0 else 0.0)
        g = rng.standard_normal(n)
        v0, dv0 = sc.CL_scaling_vector(x, g, lb, ub)
        v1, dv1 = D._cl_scaling_vector(x, g, lb, ub)
        assert np.array_equal(v0, v1) and np.array_equal(dv0, dv1)
        s = rng.standard_normal(n) * (rng.random(n) > 0.2)
        if np.any(s != 0):
            m0, h0 = sc.step_size_to_bound(x, s, lb, ub)
            m1, h1 = D._step_size_to_bound(x, s, lb, ub)
            assert m0 == m1 and np.array_equal(h0, h1)
        assert sc.in_bounds(x, lb, ub) == D._in_bounds(x, lb, ub)
        for rstep in (1e-10, 0):
            np.testing.assert_array_equal(sc.make_strictly_feasible(x, lb, ub, rstep=rstep),
                                          D._make_strictly_feasible(x, lb, ub, rstep=rstep))
        Delta = float(rng.uniform(0.5, 2))
        xin = rng.standard_normal(n)
        xin *= rng.uniform(0, 1) * Delta / np.linalg.norm(xin)
        sdir = rng.standard_normal(n)
        t0 = sc.intersect_trust_region(xin, sdir, Delta)
        t1 = D._intersect_trust_region(xin, sdir, Delta)
        np.testing.assert_allclose(t0, t1, rtol=1e-14)
        a, b, c = rng.standard_normal(3)
        lo, hi = sorted(rng.uniform(-2, 2, 2))
        q0 = sc.minimize_quadratic_1d(a, b, lo, hi, c=c)
        q1 = D._minimize_quadratic_1d(a, b, lo, hi, c=c)
        np.testing.assert_allclose(q0, q1, rtol=1e-14)
