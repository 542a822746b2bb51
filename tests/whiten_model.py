"""Host model (numpy) of the DEVICE whitening algorithm for large blocks -- test infrastructure.

csrc/whiten_large.cu does not call LAPACK: it factors the correlation matrix with a diagonally pivoted
Cholesky, orthogonalises the columns of the factor with a one-sided block Jacobi iteration and completes
the null space by projection + Cholesky-QR.  This file restates those three steps in numpy so that the
algorithm can be checked on the CPU against the oracle's eigen-decomposition (tests/test_whiten_model_cpu.py).
"""
import numpy as np

EPS = 2.220446049250313e-16


def pivoted_cholesky(A):
    """A = F F^T with diagonal pivoting, stopped when the remaining trace is at rounding level
    (1024 eps lambda_max, lambda_max bounded below by max_i |A_i.|^2 / A_ii)."""
    n = A.shape[0]
    d = np.diag(A).copy()
    used = np.zeros(n, bool)
    lam_lo = np.max(np.sum(A * A, axis=1) / np.where(d > 0, d, np.inf))
    F = np.zeros((n, n))
    r = 0
    for j in range(n):
        cand = np.where(used, -np.inf, d)
        p = int(np.argmax(cand))
        rem = d[~used]
        if not (cand[p] > 0) or rem[rem > 0].sum() <= 1024 * EPS * lam_lo:
            break
        col = (A[:, p] - F[:, :j] @ F[p, :j]) / np.sqrt(d[p])
        col[used] = 0.0
        col[p] = np.sqrt(d[p])
        F[:, j] = col
        d = d - col ** 2
        d[p] = 0.0
        used[p] = True
        r += 1
    return F[:, :r]


def _round_robin(nbe, rnd, k):
    m = nbe - 1
    a, b = (m, rnd) if k == 0 else ((rnd + k) % m, (rnd - k + m) % m)
    return min(a, b), max(a, b)


def onesided_block_jacobi(F, b=32, max_sweeps=30):
    """Orthogonalise the columns of F by rotations F <- F Q on pairs of b-column blocks (F F^T invariant)."""
    F = F.copy()
    n, r = F.shape
    nb = (r + b - 1) // b
    nbe = nb + (nb & 1)
    off_prev = np.inf
    for sweep in range(max_sweeps):
        off = 0.0
        for rnd in range(max(nbe - 1, 1)):
            for k in range(max(nbe // 2, 1)):
                I, J = _round_robin(nbe, rnd, k) if nbe > 1 else (0, 1)
                if J >= nb:
                    continue
                cols = np.r_[I * b:min((I + 1) * b, r), J * b:min((J + 1) * b, r)]
                nI = min((I + 1) * b, r) - I * b
                S = F[:, cols].T @ F[:, cols]
                dg = np.diag(S)
                with np.errstate(divide="ignore", invalid="ignore"):
                    C = np.where(np.outer(dg, dg) > 0, S * S / np.outer(dg, dg), 0.0)
                off += C[:nI, nI:].sum()
                w, Q = np.linalg.eigh(S)
                F[:, cols] = F[:, cols] @ Q[:, ::-1]            # descending: keeps the graded columns ordered
        if off < 1e-26 * r * r or (off < 1e-12 and off > 0.1 * off_prev):
            return F, sweep + 1
        off_prev = off
    raise RuntimeError("one-sided Jacobi did not converge")


def eigen(A, seed=0):
    """Eigenvalues (descending, exact zeros for the completed null space) and orthonormal eigenvectors."""
    n = A.shape[0]
    F = pivoted_cholesky(A)
    F, sweeps = onesided_block_jacobi(F)
    lam = np.sum(F * F, axis=0)
    keep = lam > 1e-13 * lam.max()
    U = F[:, keep] / np.sqrt(lam[keep])
    val = lam[keep]
    m = n - U.shape[1]
    if m > 0:
        N = np.random.default_rng(seed).standard_normal((n, m))
        for _ in range(2):
            N -= U @ (U.T @ N)
        for _ in range(2):                                       # Cholesky-QR, twice, with a re-projection
            L = np.linalg.cholesky(N.T @ N)
            N = np.linalg.solve(L, N.T).T
            N -= U @ (U.T @ N)
        L = np.linalg.cholesky(N.T @ N)
        N = np.linalg.solve(L, N.T).T
        U = np.hstack([U, N])
        val = np.concatenate([val, np.zeros(m)])
    order = np.argsort(-val, kind="stable")
    return val[order], U[:, order], sweeps
