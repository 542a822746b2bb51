"""Restatement of lsqfit._utilities.chiv / chivw (oracle only).

Follows src/lsqfit/_utilities.pyx:50-94 (``chiv.__call__``) and :96-139
(``chivw.__call__``); ``dot`` (:20-36) becomes ``_matvec`` on a Dual.

TEST INFRASTRUCTURE ONLY -- never imported by lsqfit_b200.
"""
import numpy as np

from . import dual as D


def _matvec(W, x):
    if isinstance(x, D.Dual):
        return D.Dual(W @ x.v, W @ x.d)
    return W @ x


def _delta(pdf, fcn, noprior, p):
    fp = fcn(p)
    if noprior:                                   # _utilities.pyx:74-75
        return fp - pdf.mean
    return D.concatenate([fp, p]) - pdf.mean      # :76-77


class Chiv(object):
    """chi**2 = sum(chiv(p)**2)   (_utilities.pyx:50-94)"""

    def __init__(self, pdf, fcn, noprior):
        self.pdf, self.fcn, self.noprior = pdf, fcn, noprior

    def __call__(self, p):
        pdf = self.pdf
        delta = _delta(pdf, self.fcn, self.noprior, p)
        parts = []
        iw, wgts = pdf.i_invwgts[0]
        if len(iw) > 0:
            parts.append(delta[iw] * wgts)                    # :85-89
        for iw, wgt in pdf.i_invwgts[1:]:
            parts.append(_matvec(wgt, delta[iw]))             # :90-93
        return D.concatenate(parts)


class Chivw(object):
    """inv(cov) . delta   (_utilities.pyx:96-139)"""

    def __init__(self, pdf, fcn, noprior):
        self.pdf, self.fcn, self.noprior = pdf, fcn, noprior

    def __call__(self, p):
        pdf = self.pdf
        delta = _delta(pdf, self.fcn, self.noprior, p)
        isdual = isinstance(delta, D.Dual)
        v = np.zeros(pdf.size)
        d = np.zeros((pdf.size, delta.n)) if isdual else None
        iw, wgts = pdf.i_invwgts[0]
        if len(iw) > 0:
            x = delta[iw] * wgts ** 2                         # :131-133
            v[iw] = D.value(x)
            if isdual:
                d[iw] = x.d
        for iw, wgt in pdf.i_invwgts[1:]:
            wgt2 = wgt.T @ wgt                                # sum_j outer(w_j, w_j), :134-137
            x = _matvec(wgt2, delta[iw])
            v[iw] = D.value(x)
            if isdual:
                d[iw] = x.d
        return D.Dual(v, d) if isdual else v


def build_chiv_chivw(pdf, fcn, prior_is_none):
    """_utilities.pyx:39-48"""
    return Chiv(pdf, fcn, prior_is_none), Chivw(pdf, fcn, prior_is_none)
