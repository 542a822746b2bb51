#!/usr/bin/env python
"""Phase cycle counters of the wave kernel (build with B200LM_EXTRA_CFLAGS=-DB200LM_PHASE_TICKS)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lsqfit_b200 as lb
from lsqfit_b200 import configs
from team_check import problem

def main():
    sizes = [int(a) for a in sys.argv[1:]] or [148, 1000, 10000, 40000]
    cfg, pdf = problem(8, ny=64, kind="dense")
    ny, npar = cfg["ny"], cfg["np"]
    plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts, team=32)
    p0 = torch.as_tensor(cfg["prior_mean"]).cuda()
    for B in sizes:
        means = torch.as_tensor(configs.bootstrap_means(cfg, B, 12345, cov=pdf.cov[:ny, :ny])).cuda()
        out = plan.fit_batch(means, p0, want_cov=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.fit_batch(means, p0, out=out, want_cov=False); e1.record(); torch.cuda.synchronize()
        st = plan.last_stats_ex()
        ctas = min(148, (B + 7) // 8)
        npass = st[9] / ctas
        print(json.dumps(dict(B=B, ms=round(e0.elapsed_time(e1), 3), nfev=st[0] / B, nfac=st[2] / B, passes_per_cta=round(npass, 1),
                              kclk_per_pass=dict(total=round(st[5] / st[9] / 1e3, 2), solve=round(st[4] / st[9] / 1e3, 2),
                                                 fact_in_solve=round(st[11] / st[9] / 1e3, 2), decide=round(st[12] / st[9] / 1e3, 2), secular=round(st[13] / st[9] / 1e3, 2),
                                                 step=round(st[14] / st[9] / 1e3, 2), barrier=round(st[15] / st[9] / 1e3, 2), solve_warp1=round(st[10] / st[9] / 1e3, 2),
                                                 e1=round(st[6] / st[9] / 1e3, 2), e2=round(st[7] / st[9] / 1e3, 2), e3=round(st[8] / st[9] / 1e3, 2)),
                              evals_per_pass=round(st[0] / st[9], 2))), flush=True)

if __name__ == "__main__":
    main()
