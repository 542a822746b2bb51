"""Small end-to-end exercise of every kernel family, meant to be run under compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lsqfit_b200 as lb
from lsqfit_b200 import configs
from lsqfit_b200.dense import DenseFit

cfg = configs.c5(ny=int(os.environ.get("SAN_NY", "200")), K=int(os.environ.get("SAN_K", "10")), seed=2)
for mode in ("1", "0"):
    os.environ["B200LM_WL_ONESIDED"] = mode
    fit = DenseFit((cfg["t"], cfg["ymean"], cfg["ycov"]), (cfg["prior_mean"], cfg["prior_sdev"]), svdcut=cfg["svdcut"], polish=3)
    print("dense", mode, fit.svdn, fit.chi2, fit.nit, float(np.max(np.abs(fit.p_cov - fit.cov))))
c = configs.correlator(3)
N = c["ny"] + c["np"]
full = np.zeros((N, N)); full[:c["ny"], :c["ny"]] = c["ycov"]; full[c["ny"]:, c["ny"]:] = np.diag(c["prior_sdev"] ** 2)
mean0 = np.concatenate([c["f"], c["prior_mean"]])
f = lb.nonlinear_fit(data=(c["x"], c["f"], c["ycov"]), prior=(c["prior_mean"], c["prior_sdev"]), fcn="multiexp")
bs = f.bootstrapped_fits(int(os.environ.get("SAN_B", "257")), seed=3)
m, cv = bs.pmean_stats()
print("bootstrap", m[:2], f.chi2)
w = lb.wavg([[2.1, 6.1], [1.9, 5.9]], np.diag([1.0, 1.0, 100.0, 100.0]))
print("wavg", w.mean)
D, covp = f._spec.plan(0).propagate(f.pmean.reshape(1, -1), f.cov.reshape(1, -1), f.yp_pdf.cov)
print("propagate", float(covp[0, 0, 0]))

# ---- round 2: team kernel, wave kernel, GSL policy, single-fit row kernels (split-K GEMM, in-kernel reduction), noise ----
# SAN_SKIP_TEAM=1 (synccheck): leave out the team kernel -- the tool stops at its producer/consumer named barriers
# (bar.sync id, 128 reached by the leader and the helper warps from different program points), so the kernels after it
# would never run under the tool
skip_team = os.environ.get("SAN_SKIP_TEAM") == "1"
if skip_team:
    os.environ["B200LM_TEAM"] = "1"
c8 = configs.correlator(8)
f8 = lb.nonlinear_fit(data=(c8["x"], c8["f"], c8["ycov"]), prior=(c8["prior_mean"], c8["prior_sdev"]), fcn="multiexp")
plan8 = f8._spec.plan(0)
means8 = f8.bootstrap_means(int(os.environ.get("SAN_B8", "97")), seed=5)
p08 = torch.as_tensor(c8["prior_mean"]).to(means8.device)
ref = None
for team in ((1, 32) if skip_team else (1, 4, 32)):
    if skip_team:
        os.environ["B200LM_TEAM"] = str(team)
    plan8.set_team(team)
    out = plan8.fit_batch(means8, p08, tol=(1e-8, 1e-10, 1e-10), maxit=1000)
    torch.cuda.synchronize()
    x = out.x.cpu().numpy()
    if ref is None:
        ref = x
    print("team", team, float(np.max(np.abs(x - ref))))
plan8.set_team(0)
if skip_team:
    os.environ["B200LM_TEAM"] = "1"
bg = f.bootstrapped_fits(65, seed=4, policy="gsl")
print("gsl policy", bg.pmean_stats()[0][:2])
Nu = 70000
xu = np.linspace(0.2, 2.0, Nu)
yu = 0.5 + 0.4 * np.exp(-0.7 * xu) + 1e-3 * np.random.default_rng(1).standard_normal(Nu)
fu = DenseFit((xu, yu, np.full(Nu, 1e-3)), (np.zeros(3), np.ones(3)), p0=[0.1, 0.1, 0.1], fcn="offset_exp", tol=1e-10)
print("uncorrelated", fu.nit, fu.chi2 / fu.dof, fu.fused)
fn = lb.nonlinear_fit(data=(c["x"], c["f"], c["ycov"]), prior=(c["prior_mean"], c["prior_sdev"]), fcn="multiexp", svdcut=1e-2,
                      noise=True, noise_seed=3)
print("noise", fn.svdn, fn.chi2)
