"""A TEST DOUBLE of the `lsqfit` module: the exact call protocol of the reference around its fitter seam,
without gvar (which cannot be installed here).  TEST INFRASTRUCTURE ONLY.

What is reproduced, line for line in structure (reference src/lsqfit/__init__.py):
  :110, 453        ``nonlinear_fit.FITTERS`` registry; unknown names raise ``ValueError('unknown fitter: ...')`` (:529-530)
  :539-561         ``_unpack_data`` -> ``yp_pdf``  (here: oracle.whiten.PDF stands in for gvar.PDF);
                   ``_yp_pdf`` given (simulated fits): deep copy + mean swap (:545-552)
  :562-569, 1997-2042  ``_unpack_p0`` / ``_unpack_fcn``: flatfcn = functools.partial(flatfcn_aa, x=x, fcn=fcn, pshape=...)
  :571             ``self._chiv, self._chivw = _build_chiv_chivw(yp_pdf=, fcn=flatfcn, prior=)`` -- resolved as a MODULE
                   GLOBAL at call time (that is the seam lsqfit_b200._lsqfit_hook wraps)
  :662-682         ``fit = nonlinear_fit.FITTERS[self.fitter](p0, nf, self._chiv, tol=tol, maxit=maxit, **self.fitterargs)``
                   and the attributes read from the result
  :1391-1469       ``simulated_fit_iter``: one ``nonlinear_fit`` per copy, ``_yp_pdf=self.yp_pdf``, p0 = pexact
  :1548-1642       ``bootstrapped_fit_iter``: one ``nonlinear_fit`` per copy with ``fitter=self.fitter, fcn=self.fcn,
                   p0=self.pmean, **self.fitterargs`` (:1603-1624); the whitening is recomputed by every copy
  _extras.py:1816-1829, 1164-1212   MultiFitter: the fit function is a ``_multifitfcn(flatmodels)`` object called with
                   a parameter dictionary and returning a dictionary of arrays keyed by ``datatag``; data and
                   parameters travel as flat buffers (flatfcn_dd)

GVars are replaced by plain arrays: data = (x, ymean, ycov), prior = (pmean, pcov); dictionaries of arrays stand in
for BufferDicts.  The CPU arithmetic behind the double (PDF, chiv, the scipy plugin) is the oracle's.
"""
import collections
import copy
import functools

import numpy as np

from oracle.chiv import build_chiv_chivw as _oracle_build
from oracle.fitter import scipy_least_squares as _oracle_scipy, gammaQ
from oracle.whiten import PDF as _PDF
from oracle import dual as _D


# ---- module globals that nonlinear_fit resolves at call time (the seams) -------------------------------------
def _build_chiv_chivw(yp_pdf, fcn, prior):
    """src/lsqfit/_utilities.pyx:39-48"""
    return _oracle_build(yp_pdf, fcn, prior is None)


def flatfcn_aa(p, x, fcn, pshape):
    """src/lsqfit/__init__.py:2014-2020"""
    po = p if isinstance(p, _D.Dual) else p.reshape(pshape)       # (a Dual stands in for an array of GVars; 1-d here)
    ans = fcn(po) if x is False else fcn(x, po)
    if isinstance(ans, _D.Dual):
        return ans
    return ans.flat if hasattr(ans, "flat") else np.array(ans).flat


class _Flat(object):
    """dictionary of arrays <-> one flat buffer (the part of gvar.BufferDict the protocol needs)"""

    def __init__(self, d):
        self.keys_, self.slices, n = list(d.keys()), {}, 0
        for k in self.keys_:
            m = int(np.size(d[k]))
            self.slices[k] = (slice(n, n + m), np.shape(d[k]))
            n += m
        self.size = n
        self.shape = None                     # BufferDict.shape is None (tested at :2000-2011)

    def unflatten(self, buf):
        return collections.OrderedDict((k, buf[s]) for k, (s, shp) in self.slices.items())      # (1-d entries)

    def flatten(self, d):
        return _D.concatenate([np.ravel(d[k]) if not isinstance(d[k], _D.Dual) else d[k] for k in self.keys_])


def flatfcn_dd(p, x, fcn, po, yo):
    """src/lsqfit/__init__.py:2037-2042: dictionary parameters, dictionary outputs"""
    fxp = fcn(po.unflatten(p)) if x is False else fcn(x, po.unflatten(p))
    return yo.flatten(fxp)


def _unpack_fcn(fcn, p0, y, x):
    """src/lsqfit/__init__.py:1997-2012"""
    if isinstance(y, _Flat):
        return functools.partial(flatfcn_dd, x=x, fcn=fcn, po=p0, yo=y)
    return functools.partial(flatfcn_aa, x=x, fcn=fcn, pshape=np.shape(p0))


class _scipy_plugin(object):
    """lsqfit.scipy_least_squares with the plugin signature (src/lsqfit/_scipy.py:115-181; oracle restatement)."""

    def __init__(self, x0, n, f, tol=(1e-8, 1e-10, 1e-10), maxit=1000, **extra_args):
        fit = _oracle_scipy(x0, n, f, tol=tol, maxit=maxit, **extra_args)
        for k in ("x", "cov", "f", "J", "nit", "tol", "stopping_criterion", "error", "results", "description"):
            setattr(self, k, getattr(fit, k, None))


_FITTERS = {"scipy_least_squares": _scipy_plugin}


def _cov2(c, n):
    c = np.asarray(c, dtype=float)
    return np.diag(c ** 2) if c.ndim == 1 else c


class nonlinear_fit(object):
    FITTERS = _FITTERS                      # :453 (same object as lsqfit._FITTERS, :110)
    calls = collections.Counter()           # protocol counters, read by the tests

    def __init__(self, data=None, fcn=None, prior=None, p0=None, svdcut=1e-12, eps=None, tol=1e-8, maxit=1000,
                 fitter="scipy_least_squares", _yp_pdf=None, **fitterargs):
        if fitter not in nonlinear_fit.FITTERS:
            raise ValueError("unknown fitter: " + str(fitter))                   # :529-530
        self.fitter, self.fitterargs, self.fcn = fitter, fitterargs, fcn
        x, ymean, ycov = data
        yflat = None
        if isinstance(ymean, dict):                                               # MultiFitter: dictionaries of arrays
            yflat = _Flat(ymean)
            ymean = np.concatenate([np.ravel(ymean[k]) for k in yflat.keys_])
        ymean = np.asarray(ymean, dtype=float).reshape(-1)
        ny = ymean.size
        pflat = None
        if prior is not None:
            pmean, pcov = prior
            if isinstance(pmean, dict):
                pflat = _Flat(pmean)
                pmean = np.concatenate([np.ravel(pmean[k]) for k in pflat.keys_])
            pmean = np.asarray(pmean, dtype=float).reshape(-1)
            pcov = _cov2(pcov, pmean.size)
        # ---- :539-561
        if _yp_pdf is None:
            nonlinear_fit.calls["_unpack_data"] += 1
            mean = ymean if prior is None else np.concatenate([ymean, pmean])
            N = mean.size
            full = np.zeros((N, N))
            full[:ny, :ny] = _cov2(ycov, ny)
            if prior is not None:
                full[ny:, ny:] = pcov
            yp_pdf = _PDF(mean, full, svdcut=svdcut, eps=eps)                     # gvar.PDF stand-in (:1895, 1898)
        else:
            yp_pdf = copy.deepcopy(_yp_pdf)                                       # :545-552
            yp_pdf.mean[:ny] = ymean
            if prior is not None:
                yp_pdf.mean[ny:] = pmean
            yp_pdf.meanflat = yp_pdf.mean
        self.x, self.y, self.prior, self.yp_pdf = x, (ymean, _cov2(ycov, ny)), prior, yp_pdf
        self._yflat, self._pflat = yflat, pflat
        self.svdcut, self.eps, self.svdn, self.nblocks = yp_pdf.svdcut, yp_pdf.eps, yp_pdf.nmod, yp_pdf.nblocks
        # ---- :562-569
        if p0 is None:
            psd = np.sqrt(np.diag(pcov))
            p0 = np.where(pmean != 0.0, pmean, pmean + 0.1 * psd)                 # :1947-1948
        elif isinstance(p0, dict):
            p0 = np.concatenate([np.ravel(p0[k]) for k in pflat.keys_])
        self.p0 = np.array(p0, dtype=float)
        p0f = self.p0.flatten()
        flatfcn = _unpack_fcn(fcn=self.fcn, p0=pflat if pflat is not None else self.p0, y=yflat if yflat is not None else ymean, x=x)
        # ---- :571  (module global, looked up NOW)
        import lsqfit_double as _self_module
        self._chiv, self._chivw = _self_module._build_chiv_chivw(yp_pdf=self.yp_pdf, fcn=flatfcn, prior=self.prior)
        nf = self.yp_pdf.nchiv
        self.dof = nf - self.p0.size
        # ---- :657-682
        nonlinear_fit.calls["fitter:" + fitter] += 1
        fit = nonlinear_fit.FITTERS[self.fitter](p0f, nf, self._chiv, tol=tol, maxit=maxit, **self.fitterargs)
        self.error = fit.error
        self.cov = fit.cov
        self.chi2 = np.sum(fit.f ** 2)
        self.J = fit.J
        self.residuals = np.array(fit.f)
        self.Q = gammaQ(self.dof / 2., self.chi2 / 2.)
        self.nit = fit.nit
        self.tol = fit.tol
        self.maxit = maxit
        self.stopping_criterion = fit.stopping_criterion
        self.description = getattr(fit, "description", "")
        self.fitter_results = fit.results
        self.pmean = np.array(fit.x)
        self.psdev = np.sqrt(np.diag(fit.cov))

    # ---- :1548-1642 ------------------------------------------------------------------------------------------
    def bootstrapped_fit_iter(self, n=None, datalist=None, seed=0, **kargs):
        fargs = dict(fitter=self.fitter, fcn=self.fcn)                            # :1603-1609
        fargs.update(self.fitterargs)
        fargs["p0"] = self.pmean
        fargs["prior"] = self.prior
        fargs.update(kargs)
        prior = fargs.pop("prior")
        x, (ymean, ycov) = self.x, self.y
        # gvar.bootstrap_iter stand-in: mean + L z with L L^T the (svd-corrected) covariance of y (+) prior
        C = self.yp_pdf.cov
        val, vec = np.linalg.eigh(C)
        L = vec * np.sqrt(np.clip(val, 0.0, None))
        rng = np.random.default_rng(seed)
        ny = ymean.size
        for _ in range(n):
            gb = self.yp_pdf.mean + L @ rng.standard_normal(C.shape[0])
            yb = (x, gb[:ny], ycov)
            priorb = None if prior is None else (gb[ny:], prior[1])
            yield nonlinear_fit(data=yb, prior=priorb, **fargs)                   # :1612-1624: one fit object per copy

    # ---- :1391-1469 ------------------------------------------------------------------------------------------
    def simulated_fit_iter(self, n=None, pexact=None, seed=0, **kargs):
        pexact = self.pmean if pexact is None else np.asarray(pexact, dtype=float)
        fargs = dict(fcn=self.fcn, fitter=self.fitter, p0=pexact, prior=self.prior)
        fargs.update(self.fitterargs)
        fargs.update(kargs)
        x, (ymean, ycov) = self.x, self.y
        ny = ymean.size
        fexact = np.asarray(self.fcn(x, pexact) if x is not False else self.fcn(pexact), dtype=float).reshape(-1)
        val, vec = np.linalg.eigh(self.yp_pdf.cov[:ny, :ny])
        L = vec * np.sqrt(np.clip(val, 0.0, None))
        rng = np.random.default_rng(seed)
        for _ in range(n):
            ysim = fexact + L @ rng.standard_normal(ny)
            yield nonlinear_fit(data=(x, ysim, ycov), _yp_pdf=self.yp_pdf, **fargs)   # :1457-1463


# ---- MultiFitter (src/lsqfit/_extras.py) ------------------------------------------------------------------------
class _multifitfcn(object):
    """_extras.py:1816-1829"""

    def __init__(self, flatmodels):
        self.flatmodels = flatmodels

    def __call__(self, p):
        ans = collections.OrderedDict()
        for m in self.flatmodels:
            ans[m.datatag] = m.fitfcn(p)
        return ans


class MultiFitter(object):
    """The lsqfit step of MultiFitter.lsqfit (_extras.py:1164-1212): build data, prior and the fit function from the
    models, then ``nonlinear_fit(data=fitdata, prior=fitprior, fcn=fitfcn, p0=p0, **fitterargs)``; and its bootstrap
    iterator (:1540-1586), which calls the fitter again for every copy."""

    def __init__(self, models, **fitterargs):
        self.models, self.flatmodels, self.fitterargs = models, list(models), fitterargs

    def buildfitfcn(self):
        return _multifitfcn(self.flatmodels)                                      # :1028

    def lsqfit(self, data, prior, p0=None, **kargs):
        """data: {datatag: (mean, cov)} (independent data sets); prior: ({key: mean}, {key: sdev})"""
        ymean = collections.OrderedDict((m.datatag, np.asarray(data[m.datatag][0], dtype=float)) for m in self.flatmodels)
        ny = sum(v.size for v in ymean.values())
        ycov = np.zeros((ny, ny))
        o = 0
        for m in self.flatmodels:
            c = _cov2(data[m.datatag][1], ymean[m.datatag].size)
            ycov[o:o + c.shape[0], o:o + c.shape[0]] = c
            o += c.shape[0]
        pm, ps = prior
        psd = np.concatenate([np.ravel(ps[k]) for k in pm])
        args = dict(self.fitterargs)
        args.update(kargs)
        self._last = dict(ymean=ymean, ycov=ycov, prior=(pm, psd), args=args)
        self.fit = nonlinear_fit(data=(False, ymean, ycov), prior=(pm, psd), fcn=self.buildfitfcn(), p0=p0, **args)
        return self.fit

    def bootstrapped_fit_iter(self, n, seed=0):
        last = self._last
        keys = list(last["ymean"].keys())
        flat = np.concatenate([np.ravel(last["ymean"][k]) for k in keys])
        L = np.linalg.cholesky(last["ycov"] + 1e-30 * np.eye(flat.size))
        rng = np.random.default_rng(seed)
        for _ in range(n):
            yb = flat + L @ rng.standard_normal(flat.size)
            d, o = collections.OrderedDict(), 0
            for k in keys:
                m = last["ymean"][k].size
                d[k] = yb[o:o + m]
                o += m
            yield nonlinear_fit(data=(False, d, last["ycov"]), prior=last["prior"], fcn=self.buildfitfcn(),
                                p0=self.fit.pmean, **last["args"])
