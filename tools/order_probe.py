"""Does the ORDER in which the work queue hands out the fits matter?  C3 batch (10^4 fits) on the default kernel with the
copies (a) in their natural order, (b) sorted by the evaluation count they turn out to need (longest first: the best
any ordering can do), (c) sorted by chi2 at the start point (available before the fit: one evaluation per copy),
(d) sorted by the prior part of that chi2.  Run under gpurun."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lsqfit_b200 as lb
from lsqfit_b200 import configs

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
cfg = configs.c3(B=B)
ny, npar = cfg["ny"], cfg["np"]
N = ny + npar
full = np.zeros((N, N)); full[:ny, :ny] = cfg["ycov"]; full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
mean0 = np.concatenate([cfg["f"], cfg["prior_mean"]])
pdf = lb.PDF(mean0, full, svdcut=cfg["svdcut"])
means_h = configs.bootstrap_means(cfg, B, cfg["seed"], cov=pdf.cov[:ny, :ny])
plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts, device=0)
p0_d = torch.as_tensor(cfg["p0"]).to(dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def timed(means_d, reps=7):
    out = plan.fit_batch(means_d, p0_d, tol=cfg["tol"], maxit=cfg["maxit"])
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.fit_batch(means_d, p0_d, tol=cfg["tol"], maxit=cfg["maxit"], out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


res = {}
for team in (0, 1, 32):
    plan.set_team(team)
    md = torch.as_tensor(means_h).to(dev)
    t_nat, out = timed(md)
    nit = out.nit.cpu().numpy()
    _, _, chi2 = plan.residual_jacobian(p0_d, md)
    chi2 = chi2.cpu().numpy()
    prior = (((means_h[:, ny:] - cfg["p0"][None, :]) / cfg["prior_sdev"][None, :]) ** 2).sum(axis=1)
    rng = np.random.default_rng(0)
    row = dict(natural_ms=t_nat, kernel=plan.last_team(), max_nit=int(nit.max()), mean_nit=float(nit.mean()))
    for name, key in (("oracle_longest_first", -nit), ("by_initial_chi2", -chi2), ("by_prior_chi2", -prior),
                      ("random", rng.random(B)), ("shortest_first", nit)):
        order = np.argsort(key, kind="stable")
        t, o2 = timed(torch.as_tensor(means_h[order]).to(dev))
        row[name + "_ms"] = t
    # pre-pass cost: chi2 at p0 for every copy
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0.record(); plan.residual_jacobian(p0_d, md); t1.record(); torch.cuda.synchronize()
    row["prepass_resjac_ms_with_fJ_outputs"] = t0.elapsed_time(t1)
    res["team_%d" % team] = row
    print(json.dumps({("team_%d" % team): row}), flush=True)
json.dump(res, open("gpurun_out/order_probe.json", "w"), indent=1)

# ---- the library's own ordered queue (b200lm_set_order): input order vs start-point-chi2 order, same input batch ----
md = torch.as_tensor(means_h).to(dev)
lib = {}
for team in (0, 1, 32):
    plan.set_team(team)
    row = {}
    outs = {}
    for mode in (0, 1):
        plan.set_order(mode)
        t, o = timed(md)
        row["order_%d_ms" % mode] = t
        row["order_%d_used" % mode] = plan.last_order()
        outs[mode] = o
        outs[mode] = (o.x.clone(), o.nit.clone(), o.chi2.clone())
    row["bit_identical"] = bool(torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
                                and torch.equal(outs[0][2], outs[1][2]))
    lib["team_%d" % team] = row
    print(json.dumps({("lib_team_%d" % team): row}), flush=True)
plan.set_order(None)
res["library_ordered_queue"] = lib
json.dump(res, open("gpurun_out/order_probe.json", "w"), indent=1)
