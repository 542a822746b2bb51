import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lsqfit_b200 as lb
from lsqfit_b200.whiten import whiten_blocks
n = int(sys.argv[1]); cut = float(sys.argv[2])
rng = np.random.default_rng(n)
ns = n // 2
idx = np.arange(n)
base = np.exp(-np.abs(idx[:, None] - idx[None, :]) / 50.0)
L = np.linalg.cholesky(base + 1e-10 * np.eye(n))
sig = rng.uniform(0.5, 2.0, size=n) * 1e-3
samples = (L @ rng.standard_normal((n, ns))).T * sig[None, :]
cov = np.cov(samples.T)
W, Cc, nout, nmod, logdet = whiten_blocks(np.array([n], dtype=np.int32), cov.reshape(-1), cut, None)
W = W.reshape(n, n)
D = np.diag(cov) ** -0.5
corr = cov * D[:, None] * D[None, :]
val, vec = np.linalg.eigh(corr)
vmin = cut * val[-1]
print("nmod", nmod, int((val < vmin).sum()), "logdet", logdet, np.sum(np.log(np.maximum(val, vmin))) - 2 * np.sum(np.log(D)))
# rows of W: lam^-1/2 v^T D -> recover V and lam
U = W / D[None, :]
lam = 1.0 / np.sum(U * U, axis=1)
Vd = U * np.sqrt(lam)[:, None]
print("orthogonality of device eigenvectors", np.max(np.abs(Vd @ Vd.T - np.eye(n))))
lam_ref = np.sort(np.maximum(val, vmin))[::-1]
print("eigenvalue rel err", np.max(np.abs(lam - lam_ref) / lam_ref))
R = corr @ Vd.T - Vd.T * (Vd @ corr @ Vd.T).diagonal()[None, :]
print("residual |A v - v (v^T A v)|", np.max(np.abs(R)))
