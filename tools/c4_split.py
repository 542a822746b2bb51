"""C4 job (10^6 3-exp fits, wave kernel): GPU time per kernel (CUPTI through torch.profiler)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import lsqfit_b200 as lb
from lsqfit_b200 import configs
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = configs.c4(B=B) if K == 3 else configs.c3(B=B)
ny, npar = cfg["ny"], cfg["np"]; N = ny + npar
full = np.zeros((N, N)); full[:ny, :ny] = cfg["ycov"]; full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
pdf = lb.PDF(np.concatenate([cfg["f"], cfg["prior_mean"]]), full, svdcut=cfg["svdcut"])
means = torch.as_tensor(configs.bootstrap_means(cfg, B, cfg["seed"], cov=pdf.cov[:ny, :ny], vary_prior=(K != 3))).cuda()
plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
import os
plan.set_team(int(os.environ.get("C4_TEAM", "32")))
p0 = torch.as_tensor(cfg["p0"]).cuda()
out = plan.fit_batch(means, p0, tol=cfg["tol"], maxit=cfg["maxit"])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    plan.fit_batch(means, p0, tol=cfg["tol"], maxit=cfg["maxit"], out=out)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=6, max_name_column_width=90))
