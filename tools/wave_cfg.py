#!/usr/bin/env python
"""Timings of the wave kernel configurations (B200LM_WAVE_CFG) against the one-warp and team kernels, C3 workload."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import lsqfit_b200 as lb
from lsqfit_b200 import configs
from team_check import problem, run, compare

def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1000, 10000, 40000, 160000]
    cfg, pdf = problem(8, ny=64, kind="dense")
    ny, npar = cfg["ny"], cfg["np"]
    plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
    p0 = torch.as_tensor(cfg["prior_mean"]).cuda()
    tol, maxit = (1e-8, 1e-10, 1e-10), 1000
    means = torch.as_tensor(configs.bootstrap_means(cfg, 2000, 11, cov=pdf.cov[:ny, :ny])).cuda()
    a, _, sa = run(plan, means, p0, 1, tol, maxit)
    for wc in (0,):
        os.environ["B200LM_WAVE_CFG"] = str(wc)
        b, _, sb = run(plan, means, p0, 32, tol, maxit)
        r = compare(a, b)
        r.update(wave_cfg=wc, nfev_warp=sa[0] / 2000, nfev_wave=sb[0] / 2000)
        print(json.dumps(r), flush=True)
    for B in sizes:
        means = torch.as_tensor(configs.bootstrap_means(cfg, B, 12345, cov=pdf.cov[:ny, :ny])).cuda()
        for team, wc in ((1, 0), (4, 0), (32, 0)):
            os.environ["B200LM_WAVE_CFG"] = str(wc)
            _, ms, st = run(plan, means, p0, team, tol, maxit, reps=3)
            print(json.dumps(dict(B=B, team=plan.last_team(), wave_cfg=wc if team == 32 else None, ms=round(ms, 3),
                                  fits_per_s=round(B / ms * 1e3, 1), nfev=st[0] / B, nfac=st[2] / B)), flush=True)

if __name__ == "__main__":
    main()
