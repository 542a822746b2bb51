import numpy as np, sys, time, torch
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
def timeit(m, reps=5):
    out = plan.fit_batch(m, p0, tol=cfg["tol"], maxit=cfg["maxit"])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): plan.fit_batch(m, p0, tol=cfg["tol"], maxit=cfg["maxit"], out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps, out
t, out = timeit(md)
nit = out.nit.cpu().numpy()
print("all: %.3f ms; nit mean %.1f median %d p90 %d p99 %d max %d" % (t, nit.mean(), np.median(nit), np.percentile(nit,90), np.percentile(nit,99), nit.max()))
print("nit histogram:", np.histogram(nit, bins=[0,10,20,30,50,100,200,400,1001])[0])
order = np.argsort(nit)
for frac in [0.5, 0.9, 0.99, 0.999]:
    sel = order[:int(frac*len(nit))]
    ms = md[torch.as_tensor(sel).cuda()].contiguous()
    t2, o2 = timeit(ms)
    print("fastest %.1f%%: B=%d max nit %d total nit %d: %.3f ms  -> %.0f fits/s" % (100*frac, len(sel), nit[sel].max(), nit[sel].sum(), t2, len(sel)/t2*1e3))
# replicate to larger batches
for rep in [4, 16]:
    ms = md.repeat(rep, 1).contiguous()
    t2, o2 = timeit(ms, reps=2)
    print("B=%d: %.3f ms -> %.0f fits/s" % (ms.shape[0], t2, ms.shape[0]/t2*1e3))
# --- does the initial chi2 predict the iteration count? (longest-processing-time-first scheduling)
f0, J0, chi0 = plan.residual_jacobian(p0, md)
chi0 = chi0.cpu().numpy()
print("corr(nit, chi2_0) = %.3f, corr(nit, log chi2_0) = %.3f" % (np.corrcoef(nit, chi0)[0,1], np.corrcoef(nit, np.log(chi0))[0,1]))
for name, key in [("chi2_0 descending", -chi0), ("nit descending (oracle schedule)", -nit.astype(float)), ("nit ascending (worst)", nit.astype(float))]:
    o = np.argsort(key, kind="stable")
    ms = md[torch.as_tensor(o).cuda()].contiguous()
    t2, _ = timeit(ms)
    print("%-36s %.3f ms" % (name, t2))
