"""Host model (numpy) of the DEVICE trust-region algorithm -- test infrastructure.

The CUDA engine does not port scipy or GSL.  It runs a per-fit trust-region
Levenberg-Marquardt iteration whose *decisions* (radius update, acceptance,
termination, Levenberg parameter search) follow scipy's unbounded TRF
(scipy/optimize/_lsq/trf.py: trf_no_bounds; common.py: solve_lsq_trust_region,
update_tr_radius, check_termination -- the solver behind the reference's
src/lsqfit/_scipy.py:156-161), but evaluates the secular equation with a
Cholesky factorisation of the scaled normal matrix  d.J^T J.d + alpha I
instead of an SVD of J.  This file restates that device algorithm in numpy so
that (a) the design can be validated on CPU against the oracle and (b) the
kernel has a line-by-line model.  It is never imported by lsqfit_b200.
"""
import numpy as np

EPS = np.finfo(float).eps


def _chol_solve(A, alpha, g):
    """Factor A + alpha I = L L^T; return (ok, L, p = -(A+alpha I)^-1 g)."""
    n = A.shape[0]
    M = A + alpha * np.eye(n)
    L = np.zeros_like(M)
    for j in range(n):
        s = M[j, j] - L[j, :j] @ L[j, :j]
        if not (s > 0.0) or not np.isfinite(s):
            return False, None, None
        L[j, j] = np.sqrt(s)
        L[j + 1:, j] = (M[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    y = np.linalg.solve(L, -g)
    p = np.linalg.solve(L.T, y)
    return True, L, p


def solve_tr(A, g, Delta, alpha0, rtol=0.01, max_iter=10):
    """min 1/2 p^T A p + g^T p, |p| <= Delta  (cf. common.py: solve_lsq_trust_region).

    Returns (p, alpha, n_factorizations)."""
    nfac = 1
    ok0, L0, p0 = _chol_solve(A, 0.0, g)
    full_rank = ok0
    if full_rank:
        pn = np.linalg.norm(p0)
        if pn <= Delta:
            return p0, 0.0, nfac
    alpha_upper = np.linalg.norm(g) / Delta
    if full_rank:
        phi = pn - Delta
        w = np.linalg.solve(L0, p0)
        phi_prime = -(w @ w) / pn
        alpha_lower = -phi / phi_prime
    else:
        alpha_lower = 0.0
    if alpha0 is None or (not full_rank and alpha0 == 0):
        alpha = max(0.001 * alpha_upper, (alpha_lower * alpha_upper) ** 0.5)
    else:
        alpha = alpha0
    p = None
    for it in range(max_iter):
        if alpha < alpha_lower or alpha > alpha_upper:
            alpha = max(0.001 * alpha_upper, (alpha_lower * alpha_upper) ** 0.5)
        ok, L, p = _chol_solve(A, alpha, g)
        nfac += 1
        if not ok:                       # numerically indefinite: push alpha up
            alpha_lower = max(alpha_lower, alpha)
            alpha = max(2 * alpha, 0.001 * alpha_upper)
            p = None
            continue
        pn = np.linalg.norm(p)
        phi = pn - Delta
        w = np.linalg.solve(L, p)
        phi_prime = -(w @ w) / pn
        if phi < 0:
            alpha_upper = alpha
        ratio = phi / phi_prime
        alpha_lower = max(alpha_lower, alpha - ratio)
        alpha -= (phi + Delta) * ratio / Delta
        if abs(phi) < rtol * Delta:
            break
    # scipy recomputes p at the updated alpha; one more factorisation
    ok, L, p2 = _chol_solve(A, alpha, g)
    nfac += 1
    if ok:
        p = p2
    p = p * (Delta / np.linalg.norm(p))
    return p, alpha, nfac


def solve_tr_device(A, g, Delta, alpha, gn, stats=None):
    """The kernel's variant (csrc/lm_kernel.cuh: solve_tr): same safeguarded Newton iteration, but
      * the Gauss-Newton factorisation (alpha = 0) is cached in ``gn`` across rejected trials and skipped
        when the warm-started alpha > 0 already gives a step longer than Delta;
      * the iteration stops at |phi| < 0.1 Delta (MINPACK's lmpar tolerance) and the last iterate's step is
        rescaled to the boundary instead of being recomputed at the updated alpha.
    Returns (p, alpha, number of factorisations)."""
    nfac = 0
    tried_warm = warm_ok = False
    warm = None
    if not gn.get("valid") and alpha > 0.0:
        nfac += 1
        tried_warm = True
        ok, L, pw = _chol_solve(A, alpha, g)
        if ok:
            w = np.linalg.solve(L, pw)
            warm = (pw, np.linalg.norm(pw), w @ w)
            warm_ok = True
    need_gn = not (tried_warm and warm_ok and warm[1] - Delta > 0.0)
    if not gn.get("valid") and need_gn:
        nfac += 1
        ok, L, p0 = _chol_solve(A, 0.0, g)
        gn.update(valid=True, full_rank=ok)
        if ok:
            w = np.linalg.solve(L, p0)
            gn.update(p=p0, pn=np.linalg.norm(p0), w2=w @ w)
    full_rank = gn.get("valid") and gn.get("full_rank")
    if full_rank and gn["pn"] <= Delta:
        if stats is not None:
            stats.append(("gn", nfac))
        return gn["p"], 0.0, nfac
    alpha_upper = np.linalg.norm(g) / Delta
    alpha_lower = 0.0
    p = None
    if full_rank:
        alpha_lower = (gn["pn"] - Delta) * gn["pn"] / gn["w2"]
        p = gn["p"]
    if not full_rank and alpha == 0.0:
        alpha = max(0.001 * alpha_upper, (alpha_lower * alpha_upper) ** 0.5)
    rounds = 0
    for it in range(10):
        if it == 0 and tried_warm:
            ok = warm_ok
            if ok:
                pt, pn, w2 = warm
        else:
            if alpha < alpha_lower or alpha > alpha_upper:
                alpha = max(0.001 * alpha_upper, (alpha_lower * alpha_upper) ** 0.5)
            nfac += 1
            ok, L, pt = _chol_solve(A, alpha, g)
            if ok:
                w = np.linalg.solve(L, pt)
                pn, w2 = np.linalg.norm(pt), w @ w
        rounds += 1
        if not ok:
            alpha_lower = max(alpha_lower, alpha)
            alpha = max(2 * alpha, 0.001 * alpha_upper)
            if alpha > alpha_upper:
                alpha_upper = 2 * alpha
            continue
        p = pt
        phi = pn - Delta
        if phi < 0:
            alpha_upper = alpha
        ratio = -phi * pn / w2
        alpha_lower = max(alpha_lower, alpha - ratio)
        alpha -= (phi + Delta) * ratio / Delta
        if abs(phi) < 0.1 * Delta:
            break
    if p is None:
        p = -g
    p = p * (Delta / np.linalg.norm(p))
    if stats is not None:
        stats.append(("boundary", nfac))
    return p, alpha, nfac


def _fac(A, alpha, g):
    ok, L, p = _chol_solve(A, alpha, g)
    if not ok:
        return dict(ok=False, a=alpha, pn=0.0, w2=0.0, p=None)
    w = np.linalg.solve(L, p)
    return dict(ok=True, a=alpha, pn=np.linalg.norm(p), w2=w @ w, p=p)


def solve_tr_dual(A, g, Delta, alpha, gn, lm, stats=None):
    """The kernel's two-shifts-per-round variant (csrc/lm_kernel.cuh: solve_tr_dual, used for np = 12..16):
    every factorisation round evaluates TWO shifts -- the warm-started alpha and the prediction of a linear
    model of 1/|p(alpha)| carried over from the previous trial (``lm``), or the Gauss-Newton shift 0 while
    that step is still unknown; later rounds pair the Newton and the secant iterate.  Same acceptance rule
    as solve_tr_device.  Returns (p, alpha, rounds)."""
    best = second = None
    rounds = 0

    def insert(t):
        nonlocal best, second
        if not t["ok"] or not t["a"] > 0.0:
            return
        d = abs(t["pn"] - Delta)
        if best is None or d < abs(best["pn"] - Delta):
            second, best = best, t
        elif second is None or d < abs(second["pn"] - Delta):
            second = t

    def take_gn(t):
        gn.update(valid=True, full_rank=t["ok"], p=t["p"], pn=t["pn"], w2=t["w2"])

    alpha_upper = np.linalg.norm(g) / Delta
    alpha_lower = 0.0

    def predict():
        if not lm.get("valid"):
            return -1.0
        pr = (1.0 / Delta - lm["a"]) / lm["b"]
        return pr if 0.0 < pr < 1e300 else -1.0

    if not gn.get("valid") and not alpha > 0.0:
        take_gn(_fac(A, 0.0, g))
        rounds += 1
    else:
        x1 = alpha if alpha > 0.0 else -1.0
        x2 = predict()
        if x1 < 0.0:
            nl = (gn["pn"] - Delta) * gn["pn"] / gn["w2"] if gn.get("valid") and gn.get("full_rank") else 0.0
            x1 = x2 if x2 > 0.0 else max(nl, 0.001 * alpha_upper)
            x2 = -1.0
        if x2 < 0.0 or abs(x2 - x1) < 1e-3 * x1:
            x2 = 1.5 * x1 if gn.get("valid") else 0.0
        lo, hi = _fac(A, x1, g), _fac(A, x2, g)
        rounds += 1
        if x2 == 0.0:
            take_gn(hi)
        else:
            insert(hi)
        insert(lo)
    if gn.get("valid") and gn.get("full_rank") and gn["pn"] <= Delta:
        if stats is not None:
            stats.append(rounds)
        return gn["p"], 0.0, rounds
    for it in range(5):
        if gn.get("valid") and gn.get("full_rank"):
            alpha_lower = max(alpha_lower, (gn["pn"] - Delta) * gn["pn"] / gn["w2"])
        for t in (best, second):
            if t is not None:
                if t["pn"] < Delta:
                    alpha_upper = min(alpha_upper, t["a"])
                else:
                    alpha_lower = max(alpha_lower, t["a"])
        if best is not None and abs(best["pn"] - Delta) < 0.1 * Delta and (gn.get("valid") or best["pn"] >= Delta):
            break
        if best is not None:
            phi = best["pn"] - Delta
            an = best["a"] - (phi + Delta) * (-phi * best["pn"] / best["w2"]) / Delta
        else:
            an = 2.0 * alpha if alpha > 0.0 else 0.0
        if not alpha_lower < an < alpha_upper:
            an = max(0.001 * alpha_upper, (alpha_lower * alpha_upper) ** 0.5)
        if not gn.get("valid") and best is not None and best["pn"] < Delta:
            a2 = 0.0
        else:
            a2 = -1.0
            if best is not None and second is not None and second["a"] != best["a"]:
                y0, y1 = 1.0 / best["pn"], 1.0 / second["pn"]
                if y1 != y0:
                    a2 = best["a"] + (1.0 / Delta - y0) * (second["a"] - best["a"]) / (y1 - y0)
            if not alpha_lower < a2 < alpha_upper or abs(a2 - an) < 1e-3 * an:
                a2 = 0.5 * (an + alpha_lower) if best is not None and best["pn"] < Delta else min(1.5 * an, 0.5 * (an + alpha_upper))
            if not a2 > 0.0:
                a2 = 1.5 * an
        lo, hi = _fac(A, an, g), _fac(A, a2, g)
        rounds += 1
        if a2 == 0.0:
            take_gn(hi)
            if gn["full_rank"] and gn["pn"] <= Delta:
                if stats is not None:
                    stats.append(rounds)
                return gn["p"], 0.0, rounds
        else:
            insert(hi)
        insert(lo)
        if not lo["ok"] and (a2 == 0.0 or not hi["ok"]):
            alpha_lower = max(alpha_lower, an, a2)
            alpha = max(2.0 * max(an, a2), 0.001 * alpha_upper)
            if alpha > alpha_upper:
                alpha_upper = 2.0 * alpha
    if best is None:
        p, alpha, k = solve_tr_device(A, g, Delta, alpha, gn)
        if stats is not None:
            stats.append(rounds + k)
        return p, alpha, rounds + k
    if second is not None and second["a"] != best["a"]:
        b = (1.0 / second["pn"] - 1.0 / best["pn"]) / (second["a"] - best["a"])
    else:
        b = best["w2"] / best["pn"] ** 3
    lm.update(valid=b > 0.0, b=b, a=1.0 / best["pn"] - b * best["a"])
    phi = best["pn"] - Delta
    an = best["a"] - (phi + Delta) * (-phi * best["pn"] / best["w2"]) / Delta
    alpha = an if an > 0.0 else best["a"]
    if stats is not None:
        stats.append(rounds)
    return best["p"] * (Delta / best["pn"]), alpha, rounds


def lm_fit(fun, jac, x0, xtol=1e-8, gtol=1e-10, ftol=1e-10, maxit=1000, scaler="more", device_solver=False,
           stats=None, dual=False):
    """Device algorithm model.  Returns dict(x, f, J, nfev, status, nfac).  ``device_solver=True`` uses the
    kernel's variant of the sub-problem solver (solve_tr_device) instead of the literal scipy iteration."""
    x = np.array(x0, dtype=float)
    f = fun(x)
    nfev = 1
    J = jac(x)
    m, n = J.shape
    cost = 0.5 * f @ f
    g = J.T @ f
    if scaler == "more":
        scale_inv = np.sqrt(np.sum(J ** 2, axis=0))
        scale_inv[scale_inv == 0] = 1
    else:
        scale_inv = np.ones(n)
    Delta = np.linalg.norm(x * scale_inv)
    if Delta == 0:
        Delta = 1.0
    alpha = 0.0
    status = None
    nfac = 0
    lm_state = {}
    while True:
        g_norm = np.max(np.abs(g))
        if g_norm < gtol:
            status = 1
        if status is not None or nfev >= maxit:
            break
        d = 1.0 / scale_inv
        A = (J.T @ J) * d[:, None] * d[None, :]
        g_h = d * g
        actual_reduction = -1.0
        gn = {}
        while actual_reduction <= 0 and nfev < maxit:
            if dual:
                step_h, alpha, k = solve_tr_dual(A, g_h, Delta, alpha, gn, lm_state, stats)
            elif device_solver:
                step_h, alpha, k = solve_tr_device(A, g_h, Delta, alpha, gn, stats)
            else:
                step_h, alpha, k = solve_tr(A, g_h, Delta, alpha)
            nfac += k
            predicted_reduction = -(0.5 * step_h @ A @ step_h + g_h @ step_h)
            step = d * step_h
            x_new = x + step
            f_new = fun(x_new)
            nfev += 1
            step_h_norm = np.linalg.norm(step_h)
            if not np.all(np.isfinite(f_new)):
                Delta = 0.25 * step_h_norm
                continue
            cost_new = 0.5 * f_new @ f_new
            actual_reduction = cost - cost_new
            # update_tr_radius
            if predicted_reduction > 0:
                ratio = actual_reduction / predicted_reduction
            elif predicted_reduction == actual_reduction == 0:
                ratio = 1
            else:
                ratio = 0
            Delta_new = Delta
            if ratio < 0.25:
                Delta_new = 0.25 * step_h_norm
            elif ratio > 0.75 and step_h_norm > 0.95 * Delta:
                Delta_new = Delta * 2.0
            # check_termination
            step_norm = np.linalg.norm(step)
            ft = actual_reduction < ftol * cost and ratio > 0.25
            xt = step_norm < xtol * (xtol + np.linalg.norm(x))
            if ft and xt:
                status = 4
            elif ft:
                status = 2
            elif xt:
                status = 3
            if status is not None:
                break
            alpha *= Delta / Delta_new
            Delta = Delta_new
        if actual_reduction > 0:
            x, f, cost = x_new, f_new, cost_new
            J = jac(x)
            g = J.T @ f
            if scaler == "more":
                scale_inv = np.maximum(scale_inv, np.sqrt(np.sum(J ** 2, axis=0)))
    if status is None:
        status = 0
    return dict(x=x, f=f, J=J, nfev=nfev, status=status, nfac=nfac)
