import sys, os, time, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lsqfit_b200.whiten import whiten_blocks
for n in (128, 192, 256, 384, 512):
    rng = np.random.default_rng(n)
    ns = n // 2
    idx = np.arange(n)
    base = np.exp(-np.abs(idx[:, None] - idx[None, :]) / 50.0)
    L = np.linalg.cholesky(base + 1e-10 * np.eye(n))
    samples = (L @ rng.standard_normal((n, ns))).T * 1e-3
    cov = np.cov(samples.T)
    res = {}
    for wl in (100, 512):
        os.environ["B200LM_WL_MIN"] = str(wl)
        ts = []
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = whiten_blocks(np.array([n], dtype=np.int32), cov.reshape(-1), 1e-8, None, as_torch=True)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        res[wl] = (min(ts), int(out[3][0]), float(out[4][0]))
    t0 = time.perf_counter(); D = np.diag(cov) ** -0.5; np.linalg.eigh(cov * D[:, None] * D[None, :]); th = time.perf_counter() - t0
    print(n, "block-jacobi %.1f ms  single-CTA %.1f ms  numpy eigh %.1f ms" % (1e3 * res[100][0], 1e3 * res[512][0], 1e3 * th), res[100][1:], res[512][1:])
