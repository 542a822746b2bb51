"""Restatement of lsqfit.gsl_multifit (oracle only): GSL's ``gsl_multifit_nlinear`` trust-region driver with the
Levenberg-Marquardt sub-problem (``alg='lm'``), as the reference calls it (src/lsqfit/_gsl.pyx:563-723).

GSL is a third-party C library (the reference links ``libgsl`` >= 2.4; not vendored, not installed here), so its
published algorithm is restated from the GSL 2.x sources / manual ("Nonlinear Least-Squares Fitting"):

  multifit_nlinear/trust.c    trust_init: D from the scaler, delta = 0.3 max(1, |D x|), mu from nielsen_init;
                              trust_iterate: step -> f(x + dx) -> rho -> delta update (factor_up / factor_down) ->
                              accept iff rho > 0 (new J, g, D; nielsen_accept) else nielsen_reject, more than
                              15 consecutive rejections = GSL_ENOPROG
  multifit_nlinear/lm.c       lm_step: solve [J; sqrt(mu) D] dx = -[f; 0];  lm_preduction = quadratic model
  multifit_nlinear/nielsen.c  mu0 = 1e-3 max_j (|J_j| / D_j)^2, nu = 2; accept: mu *= max(1/3, 1 - (2 rho - 1)^3), nu = 2;
                              reject: mu *= nu, nu *= 2
  multifit_nlinear/scaling.c  more: D_j = max(D_j, |J_j|) (0 -> 1 at init); levenberg: D = 1; marquardt: D_j = |J_j|
  multifit_nlinear/convergence.c   info 1: |dx_i| < xtol^2 + xtol |x_i| for all i;  info 2: max_i |g_i max(x_i, 1)| <=
                              gtol max(phi, 1), phi = |f|^2 / 2;  the ftol test is compiled out
  multifit_nlinear/fdf.c      driver: iterate, ++iter, test, until converged or iter == maxit; ENOPROG in the first
                              iteration ends the fit (info = 27); niter counts calls of iterate
  multifit_nlinear/covar.c    covar = (J^T J)^-1 from a pivoted QR of J (epsrel = 0: full rank)

PINNED by the reference's own golden output: the iteration counts printed in examples/nist.out (27 fits made with
this fitter, ``itns`` = gsl_multifit_nlinear_niter) -- tests/test_oracle_golden.py::test_gsl_lm_iteration_counts.

TEST INFRASTRUCTURE ONLY -- never imported by lsqfit_b200.
"""
import numpy

from . import dual as D
from .fitter import normalize_tol

GSL_SUCCESS, GSL_CONTINUE, GSL_EMAXITER, GSL_ENOPROG = 0, -2, 11, 27


def _colnorms(J):
    return numpy.sqrt(numpy.sum(J * J, axis=0))


class gsl_multifit(object):
    def __init__(self, x0, n, f, tol=(1e-5, 0.0, 0.0), maxit=1000, alg='lm', solver='qr', scaler='more',
                 factor_up=3.0, factor_down=2.0, avmax=0.75):
        if alg != 'lm':
            raise ValueError('oracle restates alg="lm" only, not ' + str(alg))
        if scaler not in ('more', 'levenberg', 'marquardt'):
            raise ValueError('unkown scaler ' + str(scaler))
        if solver not in ('qr', 'cholesky', 'svd'):
            raise ValueError('unkown solver ' + str(solver))
        tol = normalize_tol(tol)
        self.tol, self.maxit, self.alg, self.solver, self.scaler = tol, maxit, alg, solver, scaler
        self.factor_up, self.factor_down, self.avmax = factor_up, factor_down, avmax
        self.x0, self.n, self.error = x0, n, None
        self.description = "methods = {}/{}/{}".format(alg, scaler, solver)

        def func(x):
            return numpy.asarray(f(x), float)

        def Dfun(x):
            return numpy.array(D.deriv(f(D.Dual.variables(x)), len(x)), float)

        # ---- trust_init ------------------------------------------------------------------------------------
        x = numpy.array(x0, dtype=float)
        p = x.size
        fx, J = func(x), Dfun(x)
        g = J.T @ fx
        if scaler == 'levenberg':
            diag = numpy.ones(p)
        else:
            diag = _colnorms(J)
            diag[diag == 0.0] = 1.0
        delta = 0.3 * max(1.0, numpy.linalg.norm(diag * x))
        mu = 1.0e-3 * numpy.max(_colnorms(J) / diag) ** 2
        nu = 2
        dx = numpy.zeros(p)
        self.nfev, self.njev = 1, 1

        def lm_step(mu):
            # [J; sqrt(mu) D] dx = -[f; 0]   (least squares; the 'qr', 'cholesky' and 'svd' solvers agree to rounding)
            Aug = numpy.concatenate([J, numpy.sqrt(mu) * numpy.diag(diag)], axis=0)
            rhs = numpy.concatenate([-fx, numpy.zeros(p)])
            return numpy.linalg.lstsq(Aug, rhs, rcond=None)[0]

        def iterate():
            nonlocal x, fx, J, g, diag, delta, mu, nu, dx
            bad_steps = 0
            while True:
                dx = lm_step(mu)
                x_trial = x + dx
                f_trial = func(x_trial)
                self.nfev += 1
                # trust_calc_rho
                normf, normf_trial = numpy.linalg.norm(fx), numpy.linalg.norm(f_trial)
                if not normf_trial < normf:                       # (also rejects NaN)
                    rho = -1.0
                else:
                    u = normf_trial / normf
                    actual = 1.0 - u * u
                    beta = (J @ dx) / normf
                    pred = -(beta @ beta) - 2.0 * ((fx / normf) @ beta)
                    rho = actual / pred if pred > 0.0 else -1.0
                if rho > 0.75:
                    delta *= factor_up
                elif rho < 0.25:
                    delta /= factor_down
                if rho > 0.0:
                    J = Dfun(x_trial)
                    self.njev += 1
                    x, fx = x_trial, f_trial
                    g = J.T @ fx
                    if scaler == 'more':
                        diag = numpy.maximum(diag, _colnorms(J))
                    elif scaler == 'marquardt':
                        diag = _colnorms(J)
                        diag[diag == 0.0] = 1.0
                    b = 2.0 * rho - 1.0
                    mu *= max(0.333333333333333, 1.0 - b * b * b)
                    nu = 2
                    return GSL_SUCCESS
                mu *= float(nu)
                nu *= 2
                bad_steps += 1
                if bad_steps > 15:
                    return GSL_ENOPROG

        def test():
            xtol, gtol, _ = tol
            if numpy.all((numpy.abs(dx) < xtol * xtol + xtol * numpy.abs(x)) | (dx == 0.0)):
                return GSL_SUCCESS, 1
            gnorm = numpy.max(numpy.abs(g * numpy.maximum(x, 1.0)))
            phi = 0.5 * (fx @ fx)
            if gnorm <= gtol * max(phi, 1.0):
                return GSL_SUCCESS, 2
            return GSL_CONTINUE, 0

        # ---- gsl_multifit_nlinear_driver -------------------------------------------------------------------
        it, info = 0, 0
        while True:
            status = iterate()
            if status == GSL_ENOPROG and it == 0:
                info, status = GSL_ENOPROG, GSL_EMAXITER
                it = 1                                            # niter counts the call
                break
            it += 1
            status, info = test()
            if not (status == GSL_CONTINUE and it < maxit):
                break
        if it >= maxit and status != GSL_SUCCESS:
            status = GSL_EMAXITER
        if status:
            self.error = (status, "exceeded max number of iterations" if status == GSL_EMAXITER else str(status))
        # ---- _gsl.pyx:689-723 ------------------------------------------------------------------------------
        if 0 <= info <= 3:
            self.stopping_criterion = info
        elif info == GSL_ENOPROG:
            self.stopping_criterion = 4
        else:
            self.stopping_criterion = 0
        _, R = numpy.linalg.qr(J)
        Rinv = numpy.linalg.inv(R)
        self.cov = Rinv @ Rinv.T
        self.x, self.f, self.J = x, fx, J
        self.nit = it
        if status == GSL_EMAXITER and self.nit < self.maxit:
            self.error = "gsl_multifit can't improve on starting value; may have converged already."
        if info == 0 and self.error is None:
            self.error = "gsl_multifit didn't converge in {} iterations".format(maxit)
        self.results = None
