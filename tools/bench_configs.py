#!/usr/bin/env python
"""Throughput + parity spot checks of the other BASELINE configs on one GPU:
  C2  NIST StRD x 10^4 perturbed starts per problem (p0 = start2 * (1 + 0.1 u), u ~ U(-1,1))
  C4  3-exp correlator, simulated_fit_iter semantics, 10^6 fits (whitening shared, p0 = pexact)
Writes gpurun_out/configs_r01.json.  Usage: python tools/bench_configs.py [nfits_c4]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lsqfit_b200 as lb
from lsqfit_b200 import configs

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

res = dict(nist=[], c4=None, c1=None)
probs = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "nist.json")))["problems"]
B = 10000
for k, pr in enumerate(probs):
    x = np.array(pr["x"]); ny, npar = len(pr["y"]), len(pr["p0"])
    mean = np.concatenate([pr["y"], pr["prior_mean"]]); sd = np.concatenate([pr["ysdev"], pr["prior_sdev"]])
    plan = lb.Plan(pr["form"], npar, ny, x, [(np.arange(ny + npar), 1.0 / sd)])
    rng = np.random.default_rng(20240 + k)
    p0 = np.array(pr["p0"])[None, :] * (1 + 0.1 * rng.uniform(-1, 1, size=(B, npar)))
    p0d = torch.as_tensor(p0).cuda(); md = torch.as_tensor(mean).cuda()
    t, out = timed(lambda: plan.fit_batch(md, p0d, tol=1e-10, maxit=1000))
    o = out.numpy()
    cert, csd = np.array(pr["certified"]), np.array(pr["certified_sdev"])
    good = (o["status"] > 0) & (np.max(np.abs(o["x"] - cert[None]) / csd[None], axis=1) < 1e-2)
    res["nist"].append(dict(name=pr["name"], ny=ny, np=npar, ms=t, fits_per_s=B / t * 1e3, converged=float((o["status"] > 0).mean()),
                            at_certified_minimum=float(good.mean()), mean_nit=float(o["nit"].mean()), max_nit=int(o["nit"].max())))
    print("%-9s ny=%3d np=%d  %8.3f ms  %10.0f fits/s  conv %.4f  certified %.4f  nit mean %.1f max %d" % (
        pr["name"], ny, npar, t, B / t * 1e3, (o["status"] > 0).mean(), good.mean(), o["nit"].mean(), o["nit"].max()))
    plan.close()

# ---- C4
n4 = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
cfg = configs.c4(B=n4)
ny, npar = cfg["ny"], cfg["np"]; N = ny + npar
full = np.zeros((N, N)); full[:ny, :ny] = cfg["ycov"]; full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
mean0 = np.concatenate([cfg["f"], cfg["prior_mean"]])
pdf = lb.PDF(mean0, full, svdcut=cfg["svdcut"])
means = configs.bootstrap_means(cfg, n4, cfg["seed"], cov=pdf.cov[:ny, :ny], vary_prior=False)
plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
md = torch.as_tensor(means).cuda(); p0d = torch.as_tensor(cfg["p0"]).cuda()
t, out = timed(lambda: plan.fit_batch(md, p0d, tol=cfg["tol"], maxit=cfg["maxit"]), reps=2)
nfev, njev, nfac = plan.last_stats()
o = out.numpy()
flops = nfev * configs.eval_flops(ny, npar, cfg["K"])
res["c4"] = dict(B=n4, ms=t, fits_per_s=n4 / t * 1e3, converged=float((o["status"] > 0).mean()), mean_nit=float(o["nit"].mean()),
                 max_nit=int(o["nit"].max()), tflops=flops / (t * 1e-3) / 1e12, chi2_dof_mean=float(o["chi2"].mean() / (N - npar)))
print("C4: B=%d %.2f ms %.0f fits/s conv %.5f nit mean %.2f max %d  %.2f TFLOP/s  <chi2/dof> %.3f" % (
    n4, t, n4 / t * 1e3, res["c4"]["converged"], res["c4"]["mean_nit"], res["c4"]["max_nit"], res["c4"]["tflops"], res["c4"]["chi2_dof_mean"]))
# ---- C4 end to end on the device: Philox copies -> fit -> parameter mean/covariance over the batch
from lsqfit_b200 import bootstrap as bs
val, vec = np.linalg.eigh(pdf.cov)
Lfac = vec * np.sqrt(np.clip(val, 0, None))
Lfac[ny:, :] = 0.0                                   # simulated fits: prior means stay fixed
Ld = torch.as_tensor(Lfac).cuda(); m0d = torch.as_tensor(mean0).cuda()
def pipeline():
    mm = bs.bootstrap_means(m0d, Ld, n4, cfg["seed"])
    oo = plan.fit_batch(mm, p0d, tol=cfg["tol"], maxit=cfg["maxit"])
    ok = oo.status > 0
    xm = oo.x[ok].mean(dim=0)
    return mm, oo, xm
tg, _ = timed(lambda: bs.bootstrap_means(m0d, Ld, n4, cfg["seed"]), reps=2)
tp, (mm, oo, xm) = timed(pipeline, reps=2)
res["c4"].update(generate_ms=tg, pipeline_ms=tp, pipeline_fits_per_s=n4 / tp * 1e3,
                 pmean_bias_sigma=float(torch.max(torch.abs(xm - p0d) / oo.x.std(dim=0) * (n4 ** 0.5))))
print("C4 device pipeline: generate %.2f ms, generate+fit+stats %.2f ms (%.0f fits/s), |<p>-pexact| = %.2f sigma_mean" % (
    tg, tp, n4 / tp * 1e3, res["c4"]["pmean_bias_sigma"]))
# ---- C1: examples/simple.py, ONE fit -- latency of the whole call (device launch and host-buffer call)
ex = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "examples.json")))["examples"]["simple"]
ny1 = len(ex["ymean"]); yc = np.zeros((ny1, ny1)); i = 0
for b in ex["ycov_blocks"]:
    b = np.array(b); yc[i:i + len(b), i:i + len(b)] = b; i += len(b)
t0 = time.perf_counter()
f1 = lb.nonlinear_fit(data=(np.array(ex["x"]), ex["ymean"], yc), fcn="simple", prior=(ex["prior_mean"], ex["prior_sdev"]))
t_first = time.perf_counter() - t0
plan1 = f1._spec.plan(0)
mean1 = np.concatenate([ex["ymean"], ex["prior_mean"]]); p01 = f1.p0
md1 = torch.as_tensor(mean1).cuda(); pd1 = torch.as_tensor(p01).cuda()
t_dev, _ = timed(lambda: plan1.fit_batch(md1, pd1), reps=50)
for _ in range(5): plan1.fit_batch_host(mean1[None, :], p01)
t0 = time.perf_counter()
for _ in range(50): plan1.fit_batch_host(mean1[None, :], p01)
t_host = (time.perf_counter() - t0) / 50
res["c1"] = dict(device_launch_us=1e3 * t_dev, host_call_us=1e6 * t_host, full_nonlinear_fit_ms=1e3 * t_first, nit=int(f1.nit),
                 chi2_dof=float(f1.chi2 / f1.dof))
print("C1 simple.py single fit: device launch %.0f us, host-buffer call %.0f us, whole nonlinear_fit (whitening, fit, results) %.1f ms, nit %d" % (
    1e3 * t_dev, 1e6 * t_host, 1e3 * t_first, f1.nit))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/configs_r01.json", "w"), indent=1)
