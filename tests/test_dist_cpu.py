"""Multi-process sharding/gather logic on the gloo backend (world_size 2, CPU tensors).
The fit itself needs no collective (fits are independent); NCCL/gloo is used only to gather
the packed per-fit results and to all-reduce moments."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lsqfit_b200 import dist as lbdist


def test_shard_range_covers_batch():
    for B in (0, 1, 7, 10000, 1000003):
        for world in (1, 2, 3, 8):
            edges = [lbdist.shard_range(B, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            for (a, b), (c, d) in zip(edges[:-1], edges[1:]):
                assert b == c
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    x = torch.as_tensor(rng.normal(size=(5, 4)))
    chi2 = torch.as_tensor(rng.uniform(size=5))
    nit = torch.tensor([3, 4, 5, 1000, 7], dtype=torch.int32)
    st = torch.tensor([1, 2, 3, 0, -1], dtype=torch.int32)
    x2, c2, n2, s2 = lbdist.unpack_results(lbdist.pack_results(x, chi2, nit, st))
    assert torch.equal(x, x2) and torch.equal(chi2, c2) and torch.equal(nit, n2) and torch.equal(st, s2)


def _worker(rank, world, port, B, npar, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)                     # same "batch" on every rank
        allx = torch.as_tensor(rng.normal(size=(B, npar)))
        chi2 = torch.as_tensor(rng.uniform(size=B))
        nit = torch.as_tensor(rng.integers(3, 50, size=B).astype(np.int32))
        st = torch.as_tensor(rng.integers(0, 4, size=B).astype(np.int32))
        lo, hi = lbdist.shard_range(B, rank, world)
        packed = lbdist.pack_results(allx[lo:hi], chi2[lo:hi], nit[lo:hi], st[lo:hi])
        full = lbdist.gather_results(packed, B_total=B)
        x2, c2, n2, s2 = lbdist.unpack_results(full)
        ok = (torch.equal(x2, allx) and torch.equal(c2, chi2) and torch.equal(n2, nit) and torch.equal(s2, st))
        xbad = allx.clone()
        xbad[st <= 0] = float('nan')                         # failed fits may hold NaN: they are selected out, not multiplied by 0
        m, cov, n = lbdist.moments(xbad[lo:hi], st[lo:hi] > 0)
        sel = allx[st > 0]
        ok = ok and n == int((st > 0).sum())
        ok = ok and torch.allclose(m, sel.mean(dim=0), atol=1e-12)
        ok = ok and torch.allclose(cov, torch.as_tensor(np.cov(sel.numpy().T)), atol=1e-12)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [11, 1000])
def test_gather_and_moments_world2(B):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
