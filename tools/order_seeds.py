"""Is the ordered queue worth its start-point pass on average?  C3, 10^4 copies, eight different batches (seeds), kernels
one warp / team of 2 / team of 4, input order vs start-point-chi2 order."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import lsqfit_b200 as lb
from lsqfit_b200 import configs

B = 10000
cfg = configs.c3(B=B)
ny, npar = cfg["ny"], cfg["np"]; N = ny + npar
full = np.zeros((N, N)); full[:ny, :ny] = cfg["ycov"]; full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
pdf = lb.PDF(np.concatenate([cfg["f"], cfg["prior_mean"]]), full, svdcut=cfg["svdcut"])
plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
p0 = torch.as_tensor(cfg["p0"]).cuda()
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
rows = []
for seed in range(12345, 12353):
    means = torch.as_tensor(configs.bootstrap_means(cfg, B, seed, cov=pdf.cov[:ny, :ny])).cuda()
    row = dict(seed=seed)
    for team in (1, 2, 4):
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
avg = {k: round(float(np.mean([r[k] for r in rows])), 3) for k in rows[0] if k.startswith("team")}
print(json.dumps(dict(mean_over_seeds=avg)))
json.dump(dict(rows=rows, mean_over_seeds=avg), open("gpurun_out/order_seeds.json", "w"), indent=1)
