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
