"""Hook for a real (unmodified) lsqfit installation.  Needs lsqfit + gvar; neither is
available in the build container, so this module is exercised only by the
skip-if-missing integration test (tests/test_lsqfit_integration.py).

Seams used (reference src/lsqfit/__init__.py):
  :110-128, 453   ``nonlinear_fit.FITTERS`` registry        -> add 'b200_lm'
  :571, 2074      ``lsqfit._build_chiv_chivw`` module global -> wrap, attach ChivSpec
  :1997-2012      flatfcn is a functools.partial whose keywords hold fcn, x, ...
"""
import functools

import numpy as np

from .fitter import ChivSpec, b200_lm
from .functors import Functor


class _ChivProxy(object):
    """Behaves like the reference's chiv object, plus the ``b200`` attribute."""

    def __init__(self, chiv, spec):
        self._chiv, self.b200 = chiv, spec

    def __call__(self, *args, **kargs):
        return self._chiv(*args, **kargs)


def _find_functor(flatfcn):
    """(functor, x, pperm) behind the reference's flatfcn (a functools.partial whose keywords hold fcn, x and the
    parameter / data buffers, src/lsqfit/__init__.py:1997-2012), or (None, None, None).  A MultiFitter closure
    (``_multifitfcn``, _extras.py:1816-1829) whose models are all device-backed maps onto one composite functor."""
    from .multifit import composite
    f = flatfcn
    while isinstance(f, functools.partial):
        kw = f.keywords or {}
        fcn = kw.get("fcn")
        if isinstance(fcn, Functor):
            return fcn, kw.get("x", False), None
        if hasattr(fcn, "flatmodels") and "po" in kw and "yo" in kw:
            c = composite(fcn, kw["po"], kw["yo"])
            if c is not None:
                return c
        f = f.func
    return None, None, None


def _prior_size(prior):
    """number of prior entries: a GVar array / BufferDict (``.size`` / ``.flat``), or the (mean, cov) pair of the
    array-form test double"""
    if isinstance(prior, tuple):
        m = prior[0]
        if isinstance(m, dict):
            return int(sum(np.size(v) for v in m.values()))
        return int(np.size(m))
    if hasattr(prior, "size"):
        return int(prior.size)
    return int(np.size(getattr(prior, "flat", prior)[:]))


def install(lsqfit):
    if getattr(lsqfit, "_b200lm_installed", False):
        return lsqfit
    reference_build = lsqfit._build_chiv_chivw

    def _build_chiv_chivw(yp_pdf, fcn, prior):
        cv, cvw = reference_build(yp_pdf=yp_pdf, fcn=fcn, prior=prior)
        functor, x, pperm = _find_functor(fcn)
        if functor is None:
            return cv, cvw                      # ordinary Python fcn: CPU fitters only
        noprior = prior is None
        N = len(yp_pdf.mean)
        npar = 0 if noprior else _prior_size(prior)
        ny = N - npar
        if noprior:
            npar = None                         # taken from p0 by the fitter
        spec = ChivSpec(functor, x, yp_pdf, noprior, ny, npar if npar is not None else -1, pperm=pperm)
        return _ChivProxy(cv, spec), cvw

    class b200_lm_plugin(b200_lm):
        def __init__(self, x0, n, f, **kargs):
            spec = getattr(f, "b200", None)
            if spec is not None and spec.np < 0:
                spec.np = len(x0)
            b200_lm.__init__(self, x0, n, f, **kargs)

    from .dense import b200_dense
    lsqfit._build_chiv_chivw = _build_chiv_chivw
    lsqfit.nonlinear_fit.FITTERS["b200_lm"] = b200_lm_plugin
    lsqfit.nonlinear_fit.FITTERS["b200_dense"] = b200_dense
    lsqfit.b200_lm = b200_lm_plugin
    lsqfit._b200lm_installed = True
    return lsqfit
