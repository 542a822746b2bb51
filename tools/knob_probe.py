"""C3 batch (10^4 fits, ordered queue) against the launch-shape knobs: teams per CTA of the team kernel, warps per CTA of
the one-warp kernel.  Each configuration runs in its own process (the knobs are read from the environment at launch)."""
import json
import os
import subprocess
import sys

CHILD = r'''
import os, sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import lsqfit_b200 as lb
from lsqfit_b200 import configs
B = 10000
cfg = configs.c3(B=B)
ny, npar = cfg["ny"], cfg["np"]; N = ny + npar
full = np.zeros((N, N)); full[:ny, :ny] = cfg["ycov"]; full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
pdf = lb.PDF(np.concatenate([cfg["f"], cfg["prior_mean"]]), full, svdcut=cfg["svdcut"])
means = torch.as_tensor(configs.bootstrap_means(cfg, B, cfg["seed"], cov=pdf.cov[:ny, :ny])).cuda()
plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
plan.set_team(int(os.environ["PROBE_TEAM"]))
p0 = torch.as_tensor(cfg["p0"]).cuda()
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
out = plan.fit_batch(means, p0, tol=cfg["tol"], maxit=cfg["maxit"])
ts = []
for _ in range(7):
    flush.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.fit_batch(means, p0, tol=cfg["tol"], maxit=cfg["maxit"], out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(json.dumps(dict(ms=float(np.median(ts)), team=plan.last_team(), order=plan.last_order())))
'''
rows = []
for team, knob, vals in ((4, "B200LM_TEAMS", ("", "3", "2")), (2, "B200LM_TEAMS", ("", "6", "4")), (1, "B200LM_WARPS", ("", "10", "8", "6"))):
    for v in vals:
        env = dict(os.environ, PROBE_TEAM=str(team))
        env.pop("B200LM_TEAMS", None); env.pop("B200LM_WARPS", None)
        if v:
            env[knob] = v
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            d = dict(error=r.stderr[-300:])
        d.update(requested_team=team, knob=knob, value=v or "default")
        rows.append(d)
        print(json.dumps(d), flush=True)
json.dump(rows, open("gpurun_out/knob_probe.json", "w"), indent=1)
