import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lsqfit_b200 as lb
from lsqfit_b200 import configs
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from team_check import problem, run, compare
K = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg, pdf = problem(K, 64, "dense", 1e-12)
means = configs.bootstrap_means(cfg, B, seed=17, cov=pdf.cov[:64, :64])
plan = lb.Plan("multiexp", cfg["np"], 64, cfg["x"], pdf.i_invwgts)
tol = (1e-10, 1e-12, 1e-12)
base, _, st1 = run(plan, means, cfg["prior_mean"], 1, tol, 600)
for team in (4, 2):
    o, _, st = run(plan, means, cfg["prior_mean"], team, tol, 600)
    print(team, compare(base, o), st[:3], st1[:3])
