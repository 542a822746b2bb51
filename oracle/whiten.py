"""Restatement of gvar.PDF / gvar.regulate / gvar.svd whitening (oracle only).

The reference builds ``yp_pdf = gvar.PDF(concat(y, prior), svdcut=, eps=)``
(src/lsqfit/__init__.py:1895,1898) and consumes ``.mean .nchiv .i_invwgts
.logdet .nmod .nblocks .correction`` (src/lsqfit/__init__.py:549-561,574,723;
src/lsqfit/_utilities.pyx:59-61,104-106).  gvar (pin >=13.1.5, setup.cfg:21) is
third-party and absent from the reference tree, so its published algorithm is
restated here from doc/source/overview.rst:1546-1611 and from the layout that
tests/test_lsqfit.py:923-931 (``make_mat``) and :932-943 (``test_logdet``)
prove:

* the covariance of y(+)prior is split into its diagonal blocks (connected
  components of the non-zero pattern);
* all 1x1 blocks go to ``i_invwgts[0] = (idx, 1/sdev)``;
* for every larger block: ``D = diag(cov)^-1/2``, ``corr = D cov D`` is
  eigen-decomposed; ``svdcut > 0`` replaces eigenvalues ``< svdcut*max`` by
  ``svdcut*max`` (count -> nmod; the added covariance is the "correction");
  ``svdcut < 0`` drops those modes; rows ``W[i] = val_i^-1/2 * vec_i * D`` so
  that ``sum_i outer(W[i], W[i]) = inv(cov_block)``;
* ``logdet = sum log(val) - 2 sum log(D)  (+ 2 sum log(sdev) for 1x1 blocks)``.
* ``eps``: ``corr += eps*norm_inf(corr)*I`` and Cholesky instead of the
  eigen-decomposition (gvar.regulate documentation).  PARITY UNPINNED: the
  reference tree holds no numeric fixture for this branch.

TEST INFRASTRUCTURE ONLY -- never imported by lsqfit_b200.
"""
import numpy as np
import scipy.linalg


def cov_blocks(cov):
    """Connected components of the non-zero pattern of ``cov``.

    Returns ``(idx_diag, [idx_block, ...])`` -- indices of all 1x1 blocks and
    the index arrays of the larger blocks (each sorted, ordered by first index).
    Mirrors gvar.evalcov_blocks as used by gvar.svd/regulate.
    """
    cov = np.asarray(cov)
    n = cov.shape[0]
    label = -np.ones(n, dtype=int)
    nlab = 0
    nz = cov != 0
    for i in range(n):
        if label[i] >= 0:
            continue
        stack = [i]
        label[i] = nlab
        while stack:
            j = stack.pop()
            for k in np.nonzero(nz[j])[0]:
                if label[k] < 0:
                    label[k] = nlab
                    stack.append(k)
        nlab += 1
    diag, blocks = [], []
    for l in range(nlab):
        idx = np.nonzero(label == l)[0]
        if idx.size == 1:
            diag.append(idx[0])
        else:
            blocks.append(idx)
    return np.array(diag, dtype=np.intp), blocks


class PDF(object):
    """Whitened description of the Gaussian y(+)prior distribution."""

    def __init__(self, mean, cov, svdcut=1e-12, eps=None, noise=False, rng=None):
        mean = np.array(mean, dtype=float).reshape(-1)
        cov = np.array(cov, dtype=float)
        if cov.ndim == 1:
            cov = np.diag(cov ** 2)         # vector of sdevs
        N = mean.size
        assert cov.shape == (N, N)
        if svdcut is not None:
            eps = None                      # __init__.py:240-245: eps ignored if svdcut given
        self.svdcut, self.eps = svdcut, eps
        self.mean = mean
        self.meanflat = mean
        self.size = N
        self.cov_in = cov
        self.cov = cov.copy()               # corrected covariance (the "distribution")
        self.logdet = 0.0
        self.nmod = 0
        self.nblocks = {}
        idx0, blocks = cov_blocks(cov)
        sd0 = np.sqrt(cov[idx0, idx0])
        self.i_invwgts = [(idx0, 1.0 / sd0)]
        if idx0.size:
            self.nblocks[1] = idx0.size
        self.logdet += 2.0 * np.sum(np.log(sd0))
        nchiv = idx0.size
        for idx in blocks:
            self.nblocks[idx.size] = self.nblocks.get(idx.size, 0) + 1
            blk = cov[np.ix_(idx, idx)]
            W, logdet, nmod, newblk = (
                whiten_block_eps(blk, eps) if eps is not None else
                whiten_block_svd(blk, svdcut)
                )
            self.i_invwgts.append((idx, W))
            self.logdet += logdet
            self.nmod += nmod
            self.cov[np.ix_(idx, idx)] = newblk
            nchiv += W.shape[0]
        self.nchiv = nchiv
        self.correction_cov = self.cov - self.cov_in
        self.noise = bool(noise)
        if self.noise:
            # gvar.PDF(..., noise=True), call sites src/lsqfit/__init__.py:1895,1898: the means get one random sample of
            # the uncertainty the regulator added (covariance: corrected - input)
            self.mean += self.noise_samples(1, rng)[0]

    def noise_samples(self, n, rng=None):
        """[n, N] samples of N(0, corrected - input covariance) (eigen-decomposition of the correction)."""
        rng = np.random.default_rng() if rng is None else rng
        val, vec = np.linalg.eigh(0.5 * (self.correction_cov + self.correction_cov.T))
        L = vec * np.sqrt(np.clip(val, 0.0, None))[None, :]
        return rng.standard_normal((n, self.size)) @ L.T

    # dense helpers used by the oracle fit ----------------------------------
    def icov(self):
        """inv(cov) assembled exactly like tests/test_lsqfit.py:923-931."""
        n = self.size
        ans = np.zeros((n, n))
        i, w = self.i_invwgts[0]
        ans[i, i] = np.asarray(w) ** 2
        for i, W in self.i_invwgts[1:]:
            ans[np.ix_(i, i)] += W.T @ W
        return ans

    def copy_with_mean(self, mean):
        """__init__.py:545-552 -- simulated fits swap only the mean."""
        import copy
        new = copy.copy(self)
        new.mean = np.array(mean, dtype=float).reshape(-1)
        new.meanflat = new.mean
        return new


def whiten_block_svd(blk, svdcut):
    """One correlated block, svd-cut branch.  Returns (W, logdet, nmod, newcov)."""
    n = blk.shape[0]
    diag = np.fabs(np.diag(blk)).copy()
    diag[diag == 0.0] = 1.0
    D = diag ** -0.5
    corr = blk * D[:, None] * D[None, :]
    val, vec = np.linalg.eigh(corr)             # signed eigenvalues, ascending
    vec = vec.T                                 # rows = eigenvectors
    nmod = 0
    newblk = blk
    if svdcut is not None and svdcut != 0:
        valmin = abs(svdcut) * val[-1]
        small = val < valmin
        nmod = int(np.sum(small))
        if svdcut > 0:
            if nmod:
                # the correction uses the SIGNED eigenvalue, so that the corrected
                # covariance has exactly valmin in the modified modes.  Pinned by
                # examples/y-vs-x.out nexp=1 (lambda_min/lambda_max = -3.7e-13):
                # only then does fit.p's sdev equal the printed 0.00735(59).
                dval = valmin - val[small]
                dcorr = (vec[small].T * dval) @ vec[small]
                newblk = blk + dcorr / D[:, None] / D[None, :]
                val = np.where(small, valmin, val)
        else:
            val, vec = val[~small], vec[~small]
            # dropped modes get infinite variance; reported covariance keeps
            # only the retained modes (gvar.svd: newg)
            newblk = ((vec.T * val) @ vec) / D[:, None] / D[None, :]
    val = np.fabs(val)      # unregulated negative roundoff eigenvalues: |val| (SVD convention)
    W =(vec * D[None, :]) / np.sqrt(val)[:, None]
    W = W[::-1].copy()                            # largest eigenvalue first (gvar: decomp()[::-1])
    logdet = np.sum(np.log(val)) - 2.0 * np.sum(np.log(D))
    return W, logdet, nmod, newblk


def whiten_block_eps(blk, eps):
    """One correlated block, eps (Cholesky) branch.  PARITY UNPINNED."""
    n = blk.shape[0]
    sd = np.sqrt(np.diag(blk))
    corr = blk / sd[:, None] / sd[None, :]
    nmod = 0
    newblk = blk
    if eps is not None and eps > 0:
        shift = eps * np.linalg.norm(corr, np.inf)
        corr = corr + shift * np.eye(n)
        newblk = blk + shift * np.diag(sd ** 2)
        nmod = n
    L = np.linalg.cholesky(corr)
    # W = inv(L) diag(1/sd):  W^T W = diag(1/sd) inv(corr) diag(1/sd) = inv(cov)
    W = scipy.linalg.solve_triangular(L, np.diag(1.0 / sd), lower=True)
    logdet = 2.0 * np.sum(np.log(np.diag(L))) + 2.0 * np.sum(np.log(sd))
    return W, logdet, nmod, newblk
