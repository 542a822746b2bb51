#!/usr/bin/env python
"""Fixture for the full-size config-5 parity test (tests/test_gpu_parity.py::test_c5_full_size_vs_oracle).

Runs the CPU oracle (restatement of lsqfit.nonlinear_fit with the scipy_least_squares plugin, More' scaling)
on BASELINE config 5 -- 5000 correlated data points (sample covariance of 2500 draws), 2000 parameters,
svdcut = 1e-8 -- and stores the best-fit parameters, their standard deviations, chi2, logGBF, the iteration
count and the number of modified svd modes in tests/golden/c5_oracle.npz.  The inputs are regenerated from
lsqfit_b200/configs.py::c5 (seeded numpy), so the fixture stays small.  Takes ~10-30 minutes of CPU.

    python tests/golden/make_c5_fixture.py [ny K]
"""
import importlib.util
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("_cfg", os.path.join(ROOT, "lsqfit_b200", "configs.py"))
configs = importlib.util.module_from_spec(spec)
spec.loader.exec_module(configs)
from oracle.fit import nonlinear_fit as ofit  # noqa: E402

ny = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cfg = configs.c5(ny=ny, K=K)
warnings.simplefilter("ignore")
t0 = time.perf_counter()
fo = ofit("multiexp", cfg["x"], cfg["ymean"], cfg["ycov"], prior_mean=cfg["prior_mean"], prior_cov=cfg["prior_sdev"],
          p0=cfg["p0"], svdcut=cfg["svdcut"], tol=cfg["tol"], x_scale="jac")
dt = time.perf_counter() - t0
out = os.path.join(ROOT, "tests", "golden", "c5_oracle.npz" if (ny, K) == (5000, 1000) else "c5_oracle_%d_%d.npz" % (ny, K))
np.savez_compressed(out, pmean=fo.pmean, psdev=fo.psdev, chi2=fo.chi2, logGBF=fo.logGBF, nit=fo.nit, dof=fo.dof,
                    svdn=fo.yp_pdf.nmod, stopping_criterion=fo.stopping_criterion, cpu_seconds=dt, cores=os.cpu_count(),
                    ny=ny, K=K, data_checksum=float(np.sum(cfg["ymean"]) + np.trace(cfg["ycov"])))
print("wrote", out, "in %.0f s: chi2 %.10g nit %d svdn %d" % (dt, fo.chi2, fo.nit, fo.yp_pdf.nmod))
