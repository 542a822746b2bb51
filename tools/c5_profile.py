"""Where the GPU time of the config-5 fit goes: kernel table (torch.profiler / CUPTI sees the ctypes-launched kernels)
of ONE warm DenseFit at full size, next to its wall-clock pieces.  Run under gpurun."""
import sys
import time

import torch

sys.path.insert(0, ".")
from lsqfit_b200 import configs                     # noqa: E402
from lsqfit_b200.dense import DenseFit              # noqa: E402
from torch.profiler import profile, ProfilerActivity   # noqa: E402

cfg = configs.c5()


def make():
    return DenseFit((cfg["t"], cfg["ymean"], cfg["ycov"]), (cfg["prior_mean"], cfg["prior_sdev"]),
                    svdcut=cfg["svdcut"], tol=cfg["tol"], maxit=cfg["maxit"])


f = make()
print("first", f.times, f.nit, f.nfev_jac, f.nfac)
del f
t0 = time.perf_counter()
f = make()
print("warm", f.times, "wall", time.perf_counter() - t0)
del f
import cProfile, pstats, io
pr = cProfile.Profile()
pr.enable()
f = make()
pr.disable()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(22)
print(st.getvalue()[:5000])
print("cprofiled", f.times)
del f
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    f = make()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
