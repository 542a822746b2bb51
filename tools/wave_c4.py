#!/usr/bin/env python
"""C4 shape (3-exp correlator, np = 6): wave kernel vs one warp per fit at large batch sizes."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import lsqfit_b200 as lb
from lsqfit_b200 import configs
from team_check import problem, run, compare

def main():
    cfg, pdf = problem(3, ny=64, kind="dense")
    ny, npar = cfg["ny"], cfg["np"]
    plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
    p0 = torch.as_tensor(cfg["ptrue"]).cuda()
    tol, maxit = (1e-8, 1e-10, 1e-10), 1000
    for B in (100000, 1000000):
        means = torch.as_tensor(configs.bootstrap_means(cfg, B, 777, cov=pdf.cov[:ny, :ny], vary_prior=False)).cuda()
        res = {}
        for team in (1, 32):
            out, ms, st = run(plan, means, p0, team, tol, maxit, reps=3, want_cov=False)
            res[team] = out
            print(json.dumps(dict(K=3, B=B, team=plan.last_team(), ms=round(ms, 3), fits_per_s=round(B / ms * 1e3, 1), nfev=st[0] / B)), flush=True)
        ok = (res[1]["status"] > 0) & (res[32]["status"] > 0)
        print(json.dumps(dict(B=B, status_equal=float((res[1]["status"] == res[32]["status"]).mean()),
                              nit_equal=float((res[1]["nit"] == res[32]["nit"]).mean()),
                              dchi2=float(np.max(np.abs(res[1]["chi2"][ok] - res[32]["chi2"][ok]) / res[1]["chi2"][ok])))), flush=True)

if __name__ == "__main__":
    main()
