#!/usr/bin/env python
"""Config-5 whitening: ny = 5000 correlated points, sample covariance of Ns = 2500 draws (rank deficient),
svdcut = 1e-8.  Times the device block-Jacobi whitening and checks it against numpy on the host."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lsqfit_b200 as lb
from lsqfit_b200.whiten import whiten_blocks
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
ns = n // 2
rng = np.random.default_rng(5000)
t = 8.0 * np.arange(1, n + 1) / n
f = np.exp(-0.5 * t) + 0.5 * np.exp(-1.3 * t)
sig = 1e-3 * np.abs(f)
idx = np.arange(n)
base = np.exp(-np.abs(idx[:, None] - idx[None, :]) / 50.0)
Lc = np.linalg.cholesky(base + 1e-12 * np.eye(n))
samples = (Lc @ rng.standard_normal((n, ns))).T * sig[None, :]
cov = np.cov(samples.T)
t0 = time.perf_counter()
W, Cc, nout, nmod, logdet = whiten_blocks(np.array([n], dtype=np.int32), cov.reshape(-1), 1e-8, None)
torch.cuda.synchronize()
t_dev = time.perf_counter() - t0
t0 = time.perf_counter()
W, Cc, nout, nmod, logdet = whiten_blocks(np.array([n], dtype=np.int32), cov.reshape(-1), 1e-8, None)
torch.cuda.synchronize()
t_dev2 = time.perf_counter() - t0
t0 = time.perf_counter()
D = np.diag(cov) ** -0.5
val, vec = np.linalg.eigh(cov * D[:, None] * D[None, :])
t_cpu = time.perf_counter() - t0
valmin = 1e-8 * val[-1]
nmod_ref = int((val < valmin).sum())
used = np.where(val < valmin, valmin, val)
logdet_ref = float(np.sum(np.log(used)) - 2 * np.sum(np.log(D)))
W = W.reshape(n, n)[: nout[0]]
v = rng.standard_normal(n) * sig
chi2_dev = float(np.sum((W @ v) ** 2))
Wref = (vec.T * D[None, :]) / np.sqrt(used)[:, None]
chi2_ref = float(np.sum((Wref @ v) ** 2))
res = dict(n=n, ns=ns, device_s_first=t_dev, device_s=t_dev2, numpy_eigh_s=t_cpu, nmod=int(nmod[0]), nmod_ref=nmod_ref,
           logdet=float(logdet[0]), logdet_ref=logdet_ref, chi2_dev=chi2_dev, chi2_ref=chi2_ref, host_cores=os.cpu_count())
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/whiten_large_r01.json", "w"), indent=1)
