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
md16 = md.repeat(16, 1).contiguous()
def timeit(m, reps=3):
    out = plan.fit_batch(m, p0, tol=cfg["tol"], maxit=cfg["maxit"])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): plan.fit_batch(m, p0, tol=cfg["tol"], maxit=cfg["maxit"], out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
for w in sys.argv[1:]:
    os.environ["B200LM_WARPS"] = w
    print("warps %s: B=10k %.3f ms   B=160k %.3f ms (%.0f fits/s)" % (w, timeit(md), *(lambda t: (t, 160000/t*1e3))(timeit(md16, 2))))
