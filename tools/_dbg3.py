import numpy as np, sys, time, torch, os
sys.path.insert(0, '.')
import lsqfit_b200 as lb
from lsqfit_b200 import configs
cfg = configs.c3(B=10000)
ny, npar = cfg["ny"], cfg["np"]; N = ny + npar
full = np.zeros((N, N)); full[:ny, :ny] = cfg["ycov"]; full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
mean0 = np.concatenate([cfg["f"], cfg["prior_mean"]])
pdf = lb.PDF(mean0, full, svdcut=cfg["svdcut"])
means = configs.bootstrap_means(cfg, cfg["B"], cfg["seed"], cov=pdf.cov[:ny, :ny])
plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
md = torch.as_tensor(means).cuda(); p0 = torch.as_tensor(cfg["p0"]).cuda()
out = plan.fit_batch(md, p0, tol=cfg["tol"], maxit=cfg["maxit"])
nit = out.nit.cpu().numpy()
def timeit(m, reps=3):
    o = plan.fit_batch(m, p0, tol=cfg["tol"], maxit=cfg["maxit"])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): plan.fit_batch(m, p0, tol=cfg["tol"], maxit=cfg["maxit"], out=o)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
for target in [384, 116, 19]:
    i = int(np.argmin(np.abs(nit - target)))
    for ncopy in [1, 148, 592, 1776, 3552]:
        m = md[i:i+1].repeat(ncopy, 1).contiguous()
        t = timeit(m)
        print("fit with nit=%d x %4d copies: %.3f ms -> %.1f us/trial (%.0f cycles) per warp-trial-slot" % (nit[i], ncopy, t, t*1e3/nit[i], t*1e-3/nit[i]*1.965e9))
