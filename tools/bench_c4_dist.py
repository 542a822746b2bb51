#!/usr/bin/env python
"""Config 4 sharded: 10^6 simulated fits of the 3-exp correlator over N GPUs (torchrun, one rank per GPU).

Every rank generates ONLY its shard of the copy stream on the device (Philox counters make copies
[lo, hi) identical to rows [lo, hi) of a single-GPU run), fits it, and the packed results are
all-gathered over NCCL; moments are all-reduced.  Rank 0 then refits the first CHECK copies of every
shard on its own and compares bit for bit -- the sharded job must give exactly the single-GPU answer.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c4_dist.py [nfits]
"""
import json, os, sys, time
os.environ.pop("NCCL_DEBUG", None)
os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lsqfit_b200 as lb
from lsqfit_b200 import configs, bootstrap as bs, dist as lbdist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
cfg = configs.c4(B=B)
ny, npar = cfg["ny"], cfg["np"]; N = ny + npar
full = np.zeros((N, N)); full[:ny, :ny] = cfg["ycov"]; full[ny:, ny:] = np.diag(cfg["prior_sdev"] ** 2)
mean0 = np.concatenate([cfg["f"], cfg["prior_mean"]])
pdf = lb.PDF(mean0, full, svdcut=cfg["svdcut"], device=local)
val, vec = np.linalg.eigh(pdf.cov)
L = vec * np.sqrt(np.clip(val, 0, None)); L[ny:, :] = 0.0      # simulated fits: prior means fixed
Ld, m0d, p0d = torch.as_tensor(L).to(dev), torch.as_tensor(mean0).to(dev), torch.as_tensor(cfg["p0"]).to(dev)
plan = lb.Plan("multiexp", npar, ny, cfg["x"], pdf.i_invwgts, device=local)
lo, hi = lbdist.shard_range(B, rank, world)

def job():
    means = bs.bootstrap_means(m0d, Ld, hi - lo, cfg["seed"], first=lo, device=local)
    out = plan.fit_batch(means, p0d, tol=cfg["tol"], maxit=cfg["maxit"], want_cov=False)
    packed = lbdist.pack_results(out.x, out.chi2, out.nit, out.status)
    allp = lbdist.gather_results(packed, B_total=B)
    m, c, n = lbdist.moments(out.x, out.status > 0)
    return allp, m, c, n

def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

job(); barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps): allp, m, c, n = job()
e1.record(); barrier()
t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    x, chi2, nit, status = lbdist.unpack_results(allp)
    assert x.shape[0] == B
    CHECK = 2000
    worst = 0
    for r in range(world):                                   # single-GPU refit of the head of every shard
        l2, h2 = lbdist.shard_range(B, r, world)
        k = min(CHECK, h2 - l2)
        mm = bs.bootstrap_means(m0d, Ld, k, cfg["seed"], first=l2, device=local)
        o = plan.fit_batch(mm, p0d, tol=cfg["tol"], maxit=cfg["maxit"], want_cov=False)
        same = bool(torch.equal(o.x, x[l2:l2 + k])) and bool(torch.equal(o.chi2, chi2[l2:l2 + k]))
        worst += 0 if same else 1
    sd = torch.sqrt(torch.diagonal(c))
    bias = float(torch.max(torch.abs(m - p0d) / sd))
    print(json.dumps(dict(config=cfg["name"], n_gpus=world, B=B, ms=float(t[0]), fits_per_s=B / float(t[0]) * 1e3,
                          converged=float((status > 0).double().mean()), mean_nit=float(nit.double().mean()),
                          shards_bit_identical_to_single_gpu=(worst == 0), max_abs_pmean_minus_pexact_over_sdev=bias,
                          includes="device copy generation + fit + NCCL all-gather of [x|chi2|nit|status] + moment all-reduce")))
if world > 1: dist.destroy_process_group()
