"""The C-ABI collectives (b200lm_comm_init / b200lm_gather / b200lm_allreduce_sum, csrc/comm.cu) on two GPUs.
Run by `gpurun --gpus 2 -- python -m pytest tests/test_comm_gpu.py -m gpu`; skipped on a one-GPU box.  (The host-side
sharding logic has its world-size-2 gloo tests in tests/test_dist_cpu.py.)"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import numpy as np
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")                      # host channel for the NCCL id only
from lsqfit_b200 import dist as lbdist
comm = lbdist.Comm(rank)
dev = torch.device("cuda", rank)
x = torch.arange(12, dtype=torch.float64, device=dev).reshape(4, 3) + 100 * rank
g = comm.gather(x)
torch.cuda.synchronize()
want = torch.cat([torch.arange(12, dtype=torch.float64).reshape(4, 3) + 100 * r for r in range(world)])
assert torch.equal(g.cpu(), want), g
b = torch.full((5,), float(rank + 1), dtype=torch.float64, device=dev)
comm.allreduce_sum(b)
torch.cuda.synchronize()
assert torch.equal(b.cpu(), torch.full((5,), float(sum(range(1, world + 1))), dtype=torch.float64))
# the sharded-batch helpers on top of it
packed = torch.randn(7, 6, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
allp = lbdist.gather_results(packed, comm=comm)
assert allp.shape == (7 * world, 6) and torch.equal(allp[7 * rank: 7 * rank + 7], packed)
m, c, n = lbdist.moments(packed, torch.ones(7, dtype=torch.bool, device=dev), comm=comm)
assert n == 7 * world and torch.allclose(m, allp.mean(dim=0))
comm.close()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_c_abi_collectives_two_gpus(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
