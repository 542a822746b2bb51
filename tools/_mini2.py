import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsqfit_b200 as lb
from lsqfit_b200 import configs
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from team_check import problem, run
K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = 4
cfg, pdf = problem(K, 64, "dense", 1e-12)
means = configs.bootstrap_means(cfg, B, seed=17, cov=pdf.cov[:64, :64])
plan = lb.Plan("multiexp", cfg["np"], 64, cfg["x"], pdf.i_invwgts)
tol = (1e-10, 1e-12, 1e-12)
base, _, _ = run(plan, means, cfg["prior_mean"], 1, tol, 1, want_fJ=True)
for team in (4, 2):
    o, _, _ = run(plan, means, cfg["prior_mean"], team, tol, 1, want_fJ=True)
    print("team", team, "df", np.abs(o["f"] - base["f"]).max(), "dJ", np.abs(o["J"] - base["J"]).max(), "/", np.abs(base["J"]).max(),
          "dcov rel", np.nanmax(np.abs(o["cov"] - base["cov"]) / np.abs(base["cov"]).max()), "dchi2", np.abs(o["chi2"] / base["chi2"] - 1).max(),
          "logdet", o["logdet"][:2], base["logdet"][:2])
    JtJ = np.einsum("bij,bik->bjk", base["J"], base["J"])
    print("   cov.JtJ-1 (base)", np.abs(np.einsum("bij,bjk->bik", base["cov"], JtJ) - np.eye(cfg["np"])).max(), " (team)", np.abs(np.einsum("bij,bjk->bik", o["cov"], JtJ) - np.eye(cfg["np"])).max())
print("last_team", plan.last_team(), "stats", plan.last_stats_ex(6), "nit", o["nit"], base["nit"], "status", o["status"], "chi2", o["chi2"], base["chi2"])
