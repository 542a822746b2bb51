import numpy as np, sys, torch
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
i = int(np.argmax(nit))
m = md[i:i+1].contiguous()
for _ in range(3):
    o = plan.fit_batch(m, p0, tol=cfg["tol"], maxit=cfg["maxit"])
torch.cuda.synchronize()
print("single fit nit", int(o.nit[0]))
