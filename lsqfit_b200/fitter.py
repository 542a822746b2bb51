"""``b200_lm`` -- the new ``fitter=`` plugin for ``lsqfit.nonlinear_fit``.

Plugin contract (reference src/lsqfit/__init__.py:662-682, mirrored from
src/lsqfit/_scipy.py:115-181 and src/lsqfit/_gsl.pyx:563-723):

    fit = b200_lm(x0, n, f, tol=tol, maxit=maxit, **fitterargs)

with attributes ``x cov f J nit tol stopping_criterion error results description``.
Non-convergence is reported through ``error`` / ``stopping_criterion``, never raised.

The reference hands the fitter an opaque callable ``f`` (a ``chiv`` instance).  The
device engine needs the model identity, ``x``, the mean vector and the whitening, so ``f``
must carry a ``b200`` attribute (a ``ChivSpec``) -- attached by
``lsqfit_b200.fit.nonlinear_fit`` or, for real lsqfit, by the ``_build_chiv_chivw`` hook
installed with ``lsqfit_b200.register()``.
"""
import numpy as np

from .engine import Plan, STOPPING_CRITERION, normalize_tol


class ChivSpec(object):
    """Everything the device needs to evaluate chiv: functor, x, whitening, means."""

    def __init__(self, functor, x, pdf, noprior, ny, np_):
        self.functor, self.x, self.pdf, self.noprior = functor, x, pdf, bool(noprior)
        self.ny, self.np = int(ny), int(np_)
        self._plans = {}

    @property
    def mean(self):
        return np.asarray(self.pdf.mean, dtype=float)

    def plan(self, device=0):
        if device not in self._plans:
            self._plans[device] = Plan(self.functor, self.np, self.ny, self.x, self.pdf.i_invwgts,
                                       noprior=self.noprior, device=device)
        return self._plans[device]


class DeviceChiv(object):
    """Callable ``chiv(p)`` evaluated on the device (reference src/lsqfit/_utilities.pyx:50-94)."""

    def __init__(self, spec, device=0):
        self.b200 = spec
        self.device = device

    def __call__(self, p):
        plan = self.b200.plan(self.device)
        f, _, _ = plan.residual_jacobian(np.asarray(p, dtype=float).reshape(1, -1), self.b200.mean)
        return f[0].cpu().numpy()

    def jacobian(self, p):
        plan = self.b200.plan(self.device)
        _, J, _ = plan.residual_jacobian(np.asarray(p, dtype=float).reshape(1, -1), self.b200.mean)
        return J[0].cpu().numpy()


class b200_lm(object):
    """B200 batched Levenberg-Marquardt fitter (single-fit plugin face).

    Args mirror the reference plugins: ``x0`` start, ``n`` number of residuals, ``f`` the chiv
    callable, ``tol`` = xtol or (xtol, gtol, ftol), ``maxit`` = max function evaluations.
    Extra ``fitterargs``: ``scaler`` ('more' default | 'levenberg'), ``device`` (CUDA index), ``polish`` (max
    Gauss-Newton refinement steps after the trust-region loop; default 0 = stop where the
    reference's solver stops).
    """

    def __init__(self, x0, n, f, tol=(1e-8, 1e-10, 1e-10), maxit=1000, scaler="more", device=0,
                 polish=0, **extra_args):
        if extra_args:
            raise ValueError("b200_lm: unknown fitter arguments: " + ", ".join(sorted(extra_args)))
        spec = getattr(f, "b200", None)
        if spec is None:
            raise ValueError(
                "the b200_lm fitter needs a device functor: use lsqfit_b200.Functor(...) as fcn "
                "and install the hook with lsqfit_b200.register() (no CPU fallback exists)")
        self.tol = normalize_tol(tol)
        self.maxit = maxit
        self.n = n
        self.x0 = np.array(x0, dtype=float)
        self.description = "scaler = {}    device = cuda:{}".format(scaler, device)
        plan = spec.plan(device)
        if n != plan.nchiv:
            raise ValueError("b200_lm: n=%d does not match the whitening (%d residuals)" % (n, plan.nchiv))
        out = plan.fit_batch_host(spec.mean, self.x0.reshape(1, -1), tol=self.tol, maxit=maxit,
                                  scaler=scaler, want_cov=True, want_fJ=True, polish=polish)
        self.x = out["x"][0].copy()
        self.cov = out["cov"][0].copy()
        self.f = out["f"][0].copy()
        self.J = out["J"][0].copy()
        self.nit = int(out["nit"][0])
        self.logdet_JtJ = float(out["logdet"][0])
        status = int(out["status"][0])
        self.results = dict(status=status, nfev=self.nit, chi2=float(out["chi2"][0]),
                            logdet_JtJ=self.logdet_JtJ)
        self.stopping_criterion = STOPPING_CRITERION[status]
        self.error = None
        if status == -1:
            self.error = "b200_lm: residuals are not finite at the starting point"
        elif status == 0:
            self.error = "b200_lm: no convergence in {} function evaluations".format(maxit)
        elif not np.all(np.isfinite(self.cov)):
            self.error = "b200_lm: J^T J is singular at the solution"
