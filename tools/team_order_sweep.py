"""C3 workload: ms per batch for B x kernel (one warp / team of 2 / team of 4 / wave) x queue order -- the data behind
the default kernel policy of b200lm_fit_batch."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import lsqfit_b200 as lb
from lsqfit_b200 import configs

Bmax = 40000
cfg = configs.c3(B=Bmax)
ny, npar = cfg["ny"], cfg["np"]; N = ny + npar
full = np.zeros((N, N)); full[:ny, :ny] = cfg["ycov"]; full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
pdf = lb.PDF(np.concatenate([cfg["f"], cfg["prior_mean"]]), full, svdcut=cfg["svdcut"])
means_all = torch.as_tensor(configs.bootstrap_means(cfg, Bmax, cfg["seed"], cov=pdf.cov[:ny, :ny])).cuda()
plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
p0 = torch.as_tensor(cfg["p0"]).cuda()
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
rows = []
for B in (1000, 2048, 5000, 10000, 20000, 40000):
    means = means_all[:B].contiguous()
    row = dict(B=B)
    for team in (1, 2, 4, 32):
        for order in (0, 1):
            plan.set_team(team); plan.set_order(order)
            out = plan.fit_batch(means, p0, tol=cfg["tol"], maxit=cfg["maxit"])
            ts = []
            for _ in range(5):
                flush.zero_(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); plan.fit_batch(means, p0, tol=cfg["tol"], maxit=cfg["maxit"], out=out); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            row["team%d_order%d" % (team, order)] = round(float(np.median(ts)), 3)
    row["max_nit"] = int(out.nit.max())
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open("gpurun_out/team_order_sweep.json", "w"), indent=1)
