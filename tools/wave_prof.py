"""one wave-kernel launch on the C3 workload for ncu (B from argv)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import lsqfit_b200 as lb
from lsqfit_b200 import configs
from team_check import problem
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
team = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg, pdf = problem(8, ny=64, kind="dense")
ny, npar = cfg["ny"], cfg["np"]
plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts, team=team)
p0 = torch.as_tensor(cfg["prior_mean"]).cuda()
means = torch.as_tensor(configs.bootstrap_means(cfg, B, 12345, cov=pdf.cov[:ny, :ny])).cuda()
for _ in range(2):
    out = plan.fit_batch(means, p0, want_cov=False)
    torch.cuda.synchronize()
print("ok", int((out.status > 0).sum()))
