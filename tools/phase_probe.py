#!/usr/bin/env python
"""Phase cycle counters of the fit kernels on the C3 workload (diagnostics).
    python tools/phase_probe.py [B ...]   ; env B200LM_TEAM selects the kernel for --one"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lsqfit_b200 as lb
from lsqfit_b200 import configs

def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    sizes = [int(a) for a in args] or [1000, 10000]
    teams = [int(os.environ["B200LM_TEAM"])] if "--one" in sys.argv else [1, 2, 4]
    cfg = configs.c3()
    ny, npar = cfg["ny"], cfg["np"]
    N = ny + npar
    full = np.zeros((N, N)); full[:ny, :ny] = cfg["ycov"]; full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
    pdf = lb.PDF(np.concatenate([cfg["f"], cfg["prior_mean"]]), full, svdcut=cfg["svdcut"])
    plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
    p0 = torch.as_tensor(cfg["p0"]).cuda()
    for B in sizes:
        means = torch.as_tensor(configs.bootstrap_means(cfg, B, cfg["seed"], cov=pdf.cov[:ny, :ny])).cuda()
        for team in teams:
            os.environ["B200LM_TEAM"] = str(team)
            out = plan.fit_batch(means, p0, tol=cfg["tol"], maxit=cfg["maxit"])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.fit_batch(means, p0, tol=cfg["tol"], maxit=cfg["maxit"], out=out); e1.record()
            torch.cuda.synchronize()
            st = plan.last_stats_ex()
            nfev = st[0]
            r = dict(B=B, team=plan.last_team(), ms=e0.elapsed_time(e1), nfev_per_fit=nfev / B, nfac_per_fit=st[2] / B,
                     kclk_per_trial=dict(eval=st[3] / nfev / 1e3, solve=st[4] / nfev / 1e3, fact_in_solve=st[11] / nfev / 1e3, total=st[5] / nfev / 1e3,
                                         other=(st[5] - st[3] - st[4]) / nfev / 1e3),
                     team_eval_kclk=dict(prior_1x1=st[6] / nfev / 1e3, model_rows=st[7] / nfev / 1e3, wg=st[8] / nfev / 1e3,
                                         normal_eq=st[9] / nfev / 1e3, reduce=st[10] / nfev / 1e3))
            print(json.dumps(r), flush=True)

if __name__ == "__main__":
    main()
