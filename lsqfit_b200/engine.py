"""Plan objects and batched launches (thin layer over the C ABI).

``Plan`` owns one ``b200lm_handle``: a functor, the constant ``x`` array and the
whitening (the reference's ``yp_pdf.i_invwgts``).  ``fit_batch`` fits B independent
problems that differ only in their mean vectors / starting points -- exactly what the
reference's ``bootstrapped_fit_iter`` and ``simulated_fit_iter`` loop over one
``nonlinear_fit`` at a time (src/lsqfit/__init__.py:1391-1469, 1548-1642).

PyTorch is used only to own device buffers and streams.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi
from .functors import Functor

# solver status -> lsqfit stopping_criterion (reference src/lsqfit/_scipy.py:178-181)
# (11, 12, 14: the GSL policy's info 1 / 2 / 27, reference src/lsqfit/_gsl.pyx:689-701)
STOPPING_CRITERION = {0: 0, 1: 2, 2: 3, 3: 1, 4: 1, -1: 0, 11: 1, 12: 2, 14: 4}
SCALER = {"more": 1, "jac": 1, 1: 1, "none": 0, "levenberg": 0, None: 0, 0: 0, "marquardt": 2, 2: 2}
POLICY = {"trf": 0, "scipy": 0, "scipy_least_squares": 0, 0: 0, None: 0, "gsl": 1, "gsl_lm": 1, "gsl_multifit": 1, "lm": 1, 1: 1}


def normalize_tol(tol):
    """Exactly the reference's normalisation (src/lsqfit/_scipy.py:124-132)."""
    if np.shape(tol) == ():
        tol = (tol, 1e-10, 1e-10)
    elif np.shape(tol) == (1,):
        tol = (tol[0], 1e-10, 1e-10)
    elif np.shape(tol) == (2,):
        tol = (tol[0], tol[1], 1e-10)
    elif np.shape(tol) != (3,):
        raise ValueError("tol must be number or a 1-, 2-, or 3-tuple")
    return tuple(float(t) for t in tol)


class BatchResult(object):
    """Device-resident results of one batch (torch tensors on the plan's device)."""

    def __init__(self, x, chi2, cov, logdet, nit, status, f=None, J=None):
        self.x, self.chi2, self.cov, self.logdet = x, chi2, cov, logdet
        self.nit, self.status, self.f, self.J = nit, status, f, J

    def numpy(self):
        return dict((k, getattr(self, k).cpu().numpy()) for k in
                    ("x", "chi2", "cov", "logdet", "nit", "status", "f", "J") if getattr(self, k) is not None)


class Plan(object):
    def __init__(self, functor, np_, ny, x, i_invwgts, noprior=False, device=0, team=None):
        if isinstance(functor, str):
            functor = Functor(functor)
        self.functor = functor
        self.np, self.ny, self.noprior = int(np_), int(ny), bool(noprior)
        self.N = self.ny if self.noprior else self.ny + self.np
        self.device = int(device)
        self._h = _cabi.handle_t()
        if not torch.cuda.is_available():
            raise RuntimeError("lsqfit_b200: no CUDA device visible; this engine has no CPU fallback")
        _cabi.check(_cabi.lib.b200lm_create(functor.family, self.ny, self.np, functor.nx,
                                            1 if noprior else 0, self.device, C.byref(self._h)))
        xr = functor.xrows(x, self.ny)
        _cabi.check(_cabi.lib.b200lm_set_const(self._h, xr.ctypes.data, xr.size), self._h)
        self.set_weights(i_invwgts)
        self.tdev = torch.device("cuda", self.device)
        if team is not None:
            self.set_team(team)

    # ---- whitening ---------------------------------------------------------------
    def set_weights(self, i_invwgts):
        idx0, w0 = i_invwgts[0]
        idx0 = np.ascontiguousarray(idx0, dtype=np.int32)
        w0 = np.ascontiguousarray(w0, dtype=np.float64)
        nin = np.array([len(i) for i, _ in i_invwgts[1:]], dtype=np.int32)
        nout = np.array([len(w) for _, w in i_invwgts[1:]], dtype=np.int32)
        bidx = (np.concatenate([np.asarray(i, dtype=np.int32) for i, _ in i_invwgts[1:]])
                if len(nin) else np.zeros(0, dtype=np.int32))
        bw = (np.concatenate([np.asarray(w, dtype=np.float64).reshape(-1) for _, w in i_invwgts[1:]])
              if len(nin) else np.zeros(0))
        bidx = np.ascontiguousarray(bidx)
        bw = np.ascontiguousarray(bw)
        _cabi.check(_cabi.lib.b200lm_set_weights(
            self._h, len(idx0), idx0.ctypes.data, w0.ctypes.data, len(nin),
            nin.ctypes.data, nout.ctypes.data, bidx.ctypes.data, bw.ctypes.data), self._h)
        self.nchiv = _cabi.lib.b200lm_nchiv(self._h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _cabi.lib.b200lm_destroy(self._h)
            self._h = _cabi.handle_t()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -----------------------------------------------------------------
    def _dev(self, a, cols):
        """-> (tensor on device, stride): 1-d input = shared across the batch."""
        if isinstance(a, torch.Tensor):
            t = a.to(self.tdev, dtype=torch.float64).contiguous()
        else:
            t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(self.tdev)
        if t.dim() == 1:
            assert t.numel() == cols, (t.shape, cols)
            return t, 0
        assert t.dim() == 2 and t.shape[1] == cols, (t.shape, cols)
        return t, cols

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    # ---- the batch fit -------------------------------------------------------------
    def _set_policy(self, policy, scaler):
        """policy: 'trf' (decisions of scipy's trf, the solver behind lsqfit.scipy_least_squares; nit = function
        evaluations) or 'gsl' (decisions of lsqfit.gsl_multifit with alg='lm'; nit = iterations, maxit limits
        iterations, scaler may also be 'marquardt')."""
        if policy not in POLICY:
            raise ValueError("unknown policy %r (use 'trf' or 'gsl')" % (policy,))
        if scaler not in SCALER:
            raise ValueError("unkown scaler " + str(scaler))
        _cabi.check(_cabi.lib.b200lm_set_policy(self._h, POLICY[policy]), self._h)
        return SCALER[scaler]

    def fit_batch(self, mean, p0, tol=1e-8, maxit=1000, scaler="more", B=None,
                  want_cov=True, want_fJ=False, out=None, polish=0, policy="trf"):
        """Fit B problems on the device; inputs may be numpy arrays or device tensors.

        mean: [B, N] or [N] (shared);  p0: [B, np] or [np] (shared).
        Returns a BatchResult of device tensors (no host synchronisation)."""
        xtol, gtol, ftol = normalize_tol(tol)
        sc = self._set_policy(policy, scaler)
        tm, sm = self._dev(mean, self.N)
        tp, sp = self._dev(p0, self.np)
        if B is None:
            B = tm.shape[0] if sm else (tp.shape[0] if sp else 1)
        if (sm and tm.shape[0] != B) or (sp and tp.shape[0] != B):
            raise ValueError("mean and p0 disagree on the batch size")
        if out is None:
            kw = dict(device=self.tdev, dtype=torch.float64)
            out = BatchResult(
                torch.empty((B, self.np), **kw), torch.empty(B, **kw),
                torch.empty((B, self.np, self.np), **kw) if want_cov else None,
                torch.empty(B, **kw),
                torch.empty(B, device=self.tdev, dtype=torch.int32),
                torch.empty(B, device=self.tdev, dtype=torch.int32),
                torch.empty((B, self.nchiv), **kw) if want_fJ else None,
                torch.empty((B, self.nchiv, self.np), **kw) if want_fJ else None)
        ptr = lambda t: t.data_ptr() if t is not None else None
        _cabi.check(_cabi.lib.b200lm_fit_batch(
            self._h, B, tm.data_ptr(), sm, tp.data_ptr(), sp, xtol, gtol, ftol, int(maxit), sc, int(polish),
            out.x.data_ptr(), out.chi2.data_ptr(), ptr(out.cov), out.logdet.data_ptr(),
            out.nit.data_ptr(), out.status.data_ptr(), ptr(out.f), ptr(out.J), self._stream()), self._h)
        out._keep = (tm, tp)        # keep inputs alive until the stream has consumed them
        return out

    def fit_batch_host(self, mean, p0, tol=1e-8, maxit=1000, scaler="more", want_cov=True,
                       want_fJ=False, out=None, polish=0, policy="trf"):
        """Same fit through the host-buffer C entry point: numpy in, numpy out.
        ``out`` may hold preallocated arrays (dict with keys x, chi2, cov, logdet, nit, status)."""
        xtol, gtol, ftol = normalize_tol(tol)
        sc = self._set_policy(policy, scaler)
        # the C entry point reads raw host pointers: dtype, contiguity and shapes are checked HERE
        # (np.ascontiguousarray is a no-op for conforming arrays)
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        if mean.ndim not in (1, 2) or mean.shape[-1] != self.N:
            raise ValueError("mean must be [N] or [B, N] with N = %d, got %s" % (self.N, mean.shape))
        if p0.ndim not in (1, 2) or p0.shape[-1] != self.np:
            raise ValueError("p0 must be [np] or [B, np] with np = %d, got %s" % (self.np, p0.shape))
        sm = 0 if mean.ndim == 1 else self.N
        sp = 0 if p0.ndim == 1 else self.np
        B = mean.shape[0] if sm else (p0.shape[0] if sp else 1)
        if (sm and mean.shape[0] != B) or (sp and p0.shape[0] != B):
            raise ValueError("mean and p0 disagree on the batch size")
        shapes = dict(x=((B, self.np), np.float64), chi2=((B,), np.float64), cov=((B, self.np, self.np), np.float64),
                      logdet=((B,), np.float64), nit=((B,), np.int32), status=((B,), np.int32),
                      f=((B, self.nchiv), np.float64), J=((B, self.nchiv, self.np), np.float64))
        if out is None:
            out = dict(
                x=np.empty((B, self.np)), chi2=np.empty(B),
                cov=np.empty((B, self.np, self.np)) if want_cov else None, logdet=np.empty(B),
                nit=np.empty(B, dtype=np.int32), status=np.empty(B, dtype=np.int32),
                f=np.empty((B, self.nchiv)) if want_fJ else None,
                J=np.empty((B, self.nchiv, self.np)) if want_fJ else None)
        else:
            for k in ("x", "chi2", "logdet", "nit", "status"):
                if out.get(k) is None:
                    raise ValueError("out[%r] is required" % k)
            for k, (shp, dt) in shapes.items():
                a = out.get(k)
                if a is None:
                    continue
                if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags["C_CONTIGUOUS"]
                        and a.flags["WRITEABLE"] and tuple(a.shape) == shp):
                    raise ValueError("out[%r] must be a writable C-contiguous %s array of shape %s"
                                     % (k, np.dtype(dt).name, shp))
        ptr = lambda a: a.ctypes.data if a is not None else None
        _cabi.check(_cabi.lib.b200lm_fit_batch_host(
            self._h, B, mean.ctypes.data, sm, p0.ctypes.data, sp, xtol, gtol, ftol, int(maxit), sc, int(polish),
            ptr(out["x"]), ptr(out["chi2"]), ptr(out.get("cov")), ptr(out["logdet"]),
            ptr(out["nit"]), ptr(out["status"]), ptr(out.get("f")), ptr(out.get("J"))), self._h)
        return out

    def last_stats(self):
        """(function evaluations, Jacobian evaluations, Cholesky factorisations) of the
        last batch; synchronises."""
        buf = (C.c_ulonglong * 3)()
        _cabi.check(_cabi.lib.b200lm_last_stats(self._h, buf), self._h)
        return tuple(int(v) for v in buf)

    def last_stats_ex(self, n=16):
        """All diagnostic counters of the last batch (see b200lm_last_stats_ex); synchronises."""
        buf = (C.c_ulonglong * 16)()
        _cabi.check(_cabi.lib.b200lm_last_stats_ex(self._h, buf, int(n)), self._h)
        return [int(v) for v in buf[:n]]

    def set_order(self, mode):
        """Order of the work queue: None / -1 = default policy, 0 = input order, 1 = largest start-point chi2 first
        (b200lm_set_order: scheduling only, results unchanged)."""
        _cabi.check(_cabi.lib.b200lm_set_order(self._h, -1 if mode is None else int(mode)), self._h)

    def last_order(self):
        return int(_cabi.lib.b200lm_last_order(self._h))

    def set_team(self, team):
        """Kernel choice: 0/None = default policy (by problem shape), 1 = one warp per fit (saturated batches),
        2 / 4 = team kernel (lowest latency per trial point)."""
        _cabi.check(_cabi.lib.b200lm_set_team(self._h, int(team or 0)), self._h)

    def last_team(self):
        """Warps per fit of the last fit_batch launch (1 = one warp per fit, 2 / 4 = team kernel)."""
        return int(_cabi.lib.b200lm_last_team(self._h))

    def launch_count(self):
        return int(_cabi.lib.b200lm_launch_count(self._h))

    # ---- chiv(p) and its Jacobian ------------------------------------------------------
    def residual_jacobian(self, p, mean):
        tp, sp = self._dev(p, self.np)
        tm, sm = self._dev(mean, self.N)
        B = tp.shape[0] if sp else (tm.shape[0] if sm else 1)
        kw = dict(device=self.tdev, dtype=torch.float64)
        f = torch.empty((B, self.nchiv), **kw)
        J = torch.empty((B, self.nchiv, self.np), **kw)
        chi2 = torch.empty(B, **kw)
        _cabi.check(_cabi.lib.b200lm_residual_jacobian(
            self._h, B, tp.data_ptr(), sp, tm.data_ptr(), sm, f.data_ptr(), J.data_ptr(),
            chi2.data_ptr(), self._stream()), self._h)
        return f, J, chi2

    # ---- fit.p propagation -------------------------------------------------------------
    def propagate(self, x, cov, C_yp=None):
        """D = d p / d (y, prior)  [B, np, N]  and  cov(p) = D C D^T  [B, np, np]."""
        tx, _ = self._dev(np.atleast_2d(x) if not isinstance(x, torch.Tensor) else x, self.np)
        B = tx.shape[0]
        tc = (cov if isinstance(cov, torch.Tensor) else torch.as_tensor(np.asarray(cov, dtype=np.float64)))
        tc = tc.to(self.tdev, dtype=torch.float64).reshape(B, self.np, self.np).contiguous()
        kw = dict(device=self.tdev, dtype=torch.float64)
        D = torch.empty((B, self.np, self.N), **kw)
        covp = tC = None
        if C_yp is not None:
            tC = (C_yp if isinstance(C_yp, torch.Tensor) else torch.as_tensor(np.asarray(C_yp, dtype=np.float64)))
            tC = tC.to(self.tdev, dtype=torch.float64).contiguous()
            assert tuple(tC.shape) == (self.N, self.N)
            covp = torch.empty((B, self.np, self.np), **kw)
        _cabi.check(_cabi.lib.b200lm_propagate(
            self._h, B, tx.data_ptr(), tc.data_ptr(), tC.data_ptr() if tC is not None else None,
            D.data_ptr(), covp.data_ptr() if covp is not None else None, self._stream()), self._h)
        return D, covp
