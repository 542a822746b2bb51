"""Restatement of lsqfit.nonlinear_fit's arithmetic (oracle only), gvar-free.

Follows src/lsqfit/__init__.py:
  :539-575   yp_pdf / p0 / chiv set-up          (``_yp_pdf`` mean swap :545-552)
  :657-682   fitter dispatch + result contract
  :706-725   logGBF
  :897-922   ``_getp``  ->  D = cov.G^T.C^-1,  cov(p) = D.C.D^T
  :1947-1948 default p0
Inputs are plain arrays: data means + covariance (matrix or sdev vector), prior
means + covariance, or one joint covariance ``yp_cov`` when data and prior are
correlated.

TEST INFRASTRUCTURE ONLY -- never imported by lsqfit_b200.
"""
import numpy as np

from . import models as M
from .whiten import PDF
from .chiv import build_chiv_chivw
from .fitter import scipy_least_squares, gammaQ
from . import dual as D


def _as_cov(c, n):
    c = np.asarray(c, dtype=float)
    if c.ndim == 0:
        c = np.full(n, float(c))
    if c.ndim == 1:
        return np.diag(c ** 2)
    return c


def default_p0(prior_mean, prior_sdev):
    """__init__.py:1947-1948"""
    prior_mean = np.asarray(prior_mean, dtype=float)
    return np.where(prior_mean != 0.0, prior_mean, prior_mean + 0.1 * np.asarray(prior_sdev))


class nonlinear_fit(object):
    def __init__(self, fcn, x, ymean, ycov=None, prior_mean=None, prior_cov=None,
                 yp_cov=None, p0=None, svdcut=1e-12, eps=None, tol=1e-8, maxit=1000,
                 fitter=scipy_least_squares, _yp_pdf=None, **fitterargs):
        if isinstance(fcn, str):
            fcn = M.MODELS[fcn]
        ymean = np.asarray(ymean, dtype=float).reshape(-1)
        ny = ymean.size
        self.noprior = prior_mean is None
        if not self.noprior:
            prior_mean = np.asarray(prior_mean, dtype=float).reshape(-1)
        npar = (np.size(p0) if self.noprior else prior_mean.size)
        N = ny if self.noprior else ny + npar
        # ---- whitening (__init__.py:539-561) ---------------------------------
        if _yp_pdf is None:
            if yp_cov is None:
                yp_cov = np.zeros((N, N))
                yp_cov[:ny, :ny] = _as_cov(ycov, ny)
                if not self.noprior:
                    yp_cov[ny:, ny:] = _as_cov(prior_cov, npar)
            mean = ymean if self.noprior else np.concatenate([ymean, prior_mean])
            pdf = PDF(mean, yp_cov, svdcut=svdcut, eps=eps)
        else:
            mean = ymean if self.noprior else np.concatenate([ymean, prior_mean])
            pdf = _yp_pdf.copy_with_mean(mean)
        self.yp_pdf = pdf
        self.svdn = pdf.nmod
        self.nblocks = pdf.nblocks
        self.x = x
        self.ny, self.np = ny, npar
        # ---- p0 (__init__.py:562-565, 1912-1990) -----------------------------
        if p0 is None:
            psd = np.sqrt(np.diag(pdf.cov_in)[ny:])
            p0 = default_p0(prior_mean, psd)
        self.p0 = np.array(p0, dtype=float).reshape(-1)
        flatfcn = lambda p: fcn(x, p)
        self.flatfcn = flatfcn
        self._chiv, self._chivw = build_chiv_chivw(pdf, flatfcn, self.noprior)
        nf = pdf.nchiv
        self.dof = nf - self.p0.size
        # ---- fit (__init__.py:657-682) ---------------------------------------
        fit = fitter(self.p0, nf, self._chiv, tol=tol, maxit=maxit, **fitterargs)
        self.fitter_results = fit
        self.error = fit.error
        self.cov = fit.cov
        self.chi2 = np.sum(fit.f ** 2)
        self.J = fit.J
        self.residuals = np.array(fit.f)
        self.Q = gammaQ(self.dof / 2., self.chi2 / 2.)
        self.nit = fit.nit
        self.tol = fit.tol
        self.stopping_criterion = fit.stopping_criterion
        self.pmean = np.array(fit.x)
        self.psdev = np.sqrt(np.diag(fit.cov))
        # ---- logGBF (__init__.py:706-725) ------------------------------------
        if self.noprior:
            self.logGBF = None
        else:
            sign, ld = np.linalg.slogdet(fit.J.T.dot(fit.J))
            logdet_cov = -ld
            self.logGBF = 0.5 * (
                logdet_cov - pdf.logdet - self.chi2 - self.dof * np.log(2. * np.pi)
                )
        self._D = None

    # ---- _getp (__init__.py:897-922) ---------------------------------------
    @property
    def D(self):
        """D[a, i] = d p[a] / d buf[i],  buf = (y, prior)."""
        if self._D is None:
            cw = self._chivw(D.Dual.variables(self.pmean))
            dcw = D.deriv(cw, self.pmean.size)          # N x np  = C^-1 [G; I]
            self._D = self.cov @ dcw.T                  # :907-911  (mdotder)
        return self._D

    @property
    def p_cov(self):
        """cov(fit.p) = D C D^T with C = corrected cov(y (+) prior)   (:913-918)."""
        Dm = self.D
        return Dm @ self.yp_pdf.cov @ Dm.T

    @property
    def p_sdev(self):
        return np.sqrt(np.diag(self.p_cov))

    def fcn_values_cov(self, p=None):
        """mean and covariance of f(p) propagated from fit.p (for the 'Fit:' tables)."""
        p = self.pmean if p is None else p
        out = self.flatfcn(D.Dual.variables(p))
        G = D.deriv(out, self.pmean.size)
        return D.value(out), G @ self.p_cov @ G.T


def wavg(means, cov, index=None, svdcut=1e-12, eps=None, **fitterargs):
    """lsqfit.wavg in array form (reference src/lsqfit/_extras.py:358-516): fit of the data to
    f_i = p[index[i]] with no prior, p0 = plain average of the inputs (:478-494)."""
    means = np.asarray(means, dtype=float)
    if index is None:
        if means.ndim != 2:
            means = means.reshape(-1, 1)
        M, npar = means.shape
        index = np.tile(np.arange(npar), M)
    index = np.asarray(index, dtype=int).reshape(-1)
    y = means.reshape(-1)
    npar = int(index.max()) + 1
    p0 = np.zeros(npar)
    cnt = np.zeros(npar)
    np.add.at(p0, index, y)
    np.add.at(cnt, index, 1.0)
    return nonlinear_fit("gather", index.astype(float)[:, None], y, cov, p0=p0 / cnt, svdcut=svdcut, eps=eps, **fitterargs)
