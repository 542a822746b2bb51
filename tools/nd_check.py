"""normal_diag (one-pass normal equations of an uncorrelated fit) at 2e6 rows: kernel time as bench.py measures it,
and where the host time of the whole DenseFit goes (cProfile).  Run under gpurun."""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsqfit_b200 import _cabi
from lsqfit_b200.dense import DenseFit

dev = torch.device("cuda", 0)
Nu = 2000000
rng = np.random.default_rng(12)
xu = np.linspace(0.2, 2.0, Nu)
yu = 0.5 + 0.4 * np.exp(-0.7 * xu) + 1e-3 * rng.standard_normal(Nu)


def make():
    return DenseFit((xu, yu, np.full(Nu, 1e-3)), (np.zeros(3), np.ones(3)), p0=[0.1, 0.1, 0.1], fcn="offset_exp",
                    tol=1e-10, device=0)


fit = make()
del fit
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
fit = make()
pr.disable()
wall = time.perf_counter() - t0
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25)
print(s.getvalue()[:4000])
# which kernels the GPU time of one fit goes to (CUPTI sees the ctypes-launched kernels too)
try:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        f2 = make()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70)[:7000])
    del f2
except Exception as e:  # noqa: BLE001
    print("profiler failed:", repr(e))


def nd():
    _cabi.check(_cabi.lib.b200lm_normal_diag(fit._h, fit.x.data_ptr(), fit.d_y.data_ptr(), fit.wdiag.data_ptr(),
                                             fit.nacc.data_ptr(), fit.la.stream()), fit._h)


flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ts = []
for _ in range(9):
    flush.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); nd(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
a = fit.nacc.cpu().numpy().copy()
nd(); torch.cuda.synchronize()
b = fit.nacc.cpu().numpy().copy()
ms = float(np.median(ts))
print(json.dumps(dict(normal_diag_ms=ms, all=ts, gbs=24.0 * Nu / (ms * 1e-3) / 1e9, deterministic=bool(np.array_equal(a, b)),
                      wall_s=wall, times=fit.times, nit=int(fit.nit), chi2_dof=float(fit.chi2 / fit.dof))))
