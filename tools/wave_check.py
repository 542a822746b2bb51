#!/usr/bin/env python
"""GPU check of the wave kernel (lm_wave.cuh, b200lm_set_team(32)) against the one-warp kernel: same fits on the
same inputs (results agree to rounding / to the solver's stopping noise), plus timings at several batch sizes.

    python tools/wave_check.py [K ...] > gpurun_out/wave_check.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import lsqfit_b200 as lb  # noqa: E402
from lsqfit_b200 import configs  # noqa: E402
from team_check import problem, run, compare  # noqa: E402


def main():
    Ks = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [8, 3]
    sizes = [1000, 10000, 40000] + ([160000] if "--big" in sys.argv else [])
    for K in Ks:
        cfg, pdf = problem(K, ny=64, kind="dense")
        ny, npar = cfg["ny"], cfg["np"]
        plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
        p0 = torch.as_tensor(cfg["prior_mean"]).cuda()
        tol, maxit = (1e-8, 1e-10, 1e-10), 1000
        means = torch.as_tensor(configs.bootstrap_means(cfg, 2000, 11, cov=pdf.cov[:ny, :ny])).cuda()
        a, _, sa = run(plan, means, p0, 1, tol, maxit, want_fJ=True)
        b, _, sb = run(plan, means, p0, 32, tol, maxit, want_fJ=True)
        r = compare(a, b)
        r.update(K=K, last_team=plan.last_team(), nfev_warp=sa[0] / 2000, nfev_wave=sb[0] / 2000,
                 nfac_warp=sa[2] / 2000, nfac_wave=sb[2] / 2000,
                 status_hist_wave=np.bincount(b["status"] + 2).tolist())
        print(json.dumps(r), flush=True)
        # tight + polish: both at the stationary point
        a, _, _ = run(plan, means[:500], p0, 1, (1e-15, 0.0, 0.0), maxit, polish=50)
        b, _, _ = run(plan, means[:500], p0, 32, (1e-15, 0.0, 0.0), maxit, polish=50)
        r = compare(a, b)
        r.update(K=K, what="tight+polish")
        print(json.dumps(r), flush=True)
        for B in sizes:
            means = torch.as_tensor(configs.bootstrap_means(cfg, B, 12345, cov=pdf.cov[:ny, :ny])).cuda()
            for team in (1, 4, 32):
                _, ms, st = run(plan, means, p0, team, tol, maxit, reps=3)
                print(json.dumps(dict(K=K, B=B, team=plan.last_team(), ms=round(ms, 3), fits_per_s=round(B / ms * 1e3, 1),
                                      nfev=st[0] / B, nfac=st[2] / B)), flush=True)
        plan.close()


if __name__ == "__main__":
    main()
