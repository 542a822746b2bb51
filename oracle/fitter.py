"""Restatement of lsqfit.scipy_least_squares (oracle only).

Follows src/lsqfit/_scipy.py:115-181 line by line; the only change is that
the Jacobian comes from ``dual.Dual`` instead of ``gvar.valder``
(:144, :149-154).  scipy (1.18.1 here) is the real third-party solver the
reference calls, so this leg *is* the reference's numerical engine.

``gammaQ`` follows src/lsqfit/_scipy.py:16-18.

TEST INFRASTRUCTURE ONLY -- never imported by lsqfit_b200.
"""
import numpy

from . import dual as D


def gammaQ(a, x):
    from scipy.special import gammaincc
    return gammaincc(a, x)


def normalize_tol(tol):
    """_scipy.py:124-132"""
    if numpy.shape(tol) == ():
        tol = (tol, 1e-10, 1e-10)
    elif numpy.shape(tol) == (1,):
        tol = (tol[0], 1e-10, 1e-10)
    elif numpy.shape(tol) == (2,):
        tol = (tol[0], tol[1], 1e-10)
    elif numpy.shape(tol) != (3,):
        raise ValueError("tol must be number or a 1-, 2-, or 3-tuple")
    return tuple(tol)


class scipy_least_squares(object):
    def __init__(self, x0, n, f, tol=(1e-8, 1e-8, 1e-8), maxit=1000, **extra_args):
        from scipy.optimize import least_squares
        from scipy.linalg import svd

        tol = normalize_tol(tol)
        self.tol = tol
        self.description = 'method = {}'.format(
            'trf'
            if 'method' not in extra_args or extra_args['method'] is None else
            extra_args['method']
            )
        self.maxit = maxit
        self.n = n
        self.error = None
        self.x0 = x0
        x0 = numpy.asarray(x0, dtype=float)

        def func(x):
            return numpy.asarray(f(x), float)

        def Dfun(x):
            fx = f(D.Dual.variables(x))
            return numpy.array(D.deriv(fx, len(x)), float)

        fit = least_squares(
            fun=func, jac=Dfun, x0=x0,
            xtol=tol[0], gtol=tol[1], ftol=tol[2],
            max_nfev=maxit,
            **extra_args
            )
        if fit.status > 4:
            raise RuntimeError('fit crashed -- ' + fit.message)
        self.x = fit.x
        self.f = func(self.x)
        self.J = Dfun(self.x)
        self.nit = fit.nfev
        self.results = fit

        # covariance (_scipy.py:170-175)
        _, _s, _VT = svd(fit.jac, full_matrices=False)
        _threshold = numpy.finfo(float).eps * max(fit.jac.shape) * _s[0]
        _s = _s[_s > _threshold]
        _VT = _VT[:_s.size]
        self.cov = numpy.dot(_VT.T / _s**2, _VT)

        self.error = None
        if fit.status < 0:
            self.stopping_criterion = 0
        else:
            self.stopping_criterion = {0: 0, 1: 2, 2: 3, 3: 1, 4: 1}[fit.status]
