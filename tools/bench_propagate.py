import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lsqfit_b200 as lb
from lsqfit_b200 import configs
cfg = configs.c3(B=10000)
ny, npar = cfg["ny"], cfg["np"]; N = ny + npar
full = np.zeros((N, N)); full[:ny, :ny] = cfg["ycov"]; full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
mean0 = np.concatenate([cfg["f"], cfg["prior_mean"]])
pdf = lb.PDF(mean0, full, svdcut=cfg["svdcut"])
means = configs.bootstrap_means(cfg, 10000, cfg["seed"], cov=pdf.cov[:ny, :ny])
plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts)
out = plan.fit_batch(means, cfg["p0"], tol=cfg["tol"])
C = torch.as_tensor(pdf.cov).cuda()
for _ in range(2): D, covp = plan.propagate(out.x, out.cov, C)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): D, covp = plan.propagate(out.x, out.cov, C)
e1.record(); torch.cuda.synchronize()
ok = out.status > 0
rel = torch.max(torch.abs(covp[ok] - out.cov[ok]) / torch.sqrt(torch.diagonal(out.cov[ok], dim1=1, dim2=2)[:, :, None] * torch.diagonal(out.cov[ok], dim1=1, dim2=2)[:, None, :]))
print("propagate B=10000 (C3): %.2f ms per call; max |cov(fit.p) - fit.cov| (correlation metric) %.2e" % (e0.elapsed_time(e1) / 5, float(rel)))
