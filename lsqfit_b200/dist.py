"""Multi-GPU sharding of a batch of independent fits (one process per GPU).

Fits are independent, so the batch index range is cut into contiguous shards (gathered
output order == input order) and the fitting itself needs no collective.  NCCL (over
NVLink/NVSwitch) is used only after the fit: an all-gather of the packed per-fit results
and an all-reduce of first/second moments for bootstrap averages.
The same code runs on the gloo backend with CPU tensors for the world_size-2 tests.
"""
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """Contiguous [lo, hi) of rank's share of B fits; sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(x, chi2, nit, status):
    """[B, np+3] float64: x | chi2 | nit | status (small, one row per fit)."""
    return torch.cat([x, chi2[:, None], nit.to(torch.float64)[:, None],
                      status.to(torch.float64)[:, None]], dim=1).contiguous()


def unpack_results(packed):
    npar = packed.shape[1] - 3
    return (packed[:, :npar], packed[:, npar], packed[:, npar + 1].to(torch.int32),
            packed[:, npar + 2].to(torch.int32))


def gather_results(packed, B_total=None, group=None):
    """All-gather the packed results of every rank in rank order.  Shards may differ in size
    by one row, so they are padded to the largest shard and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        return packed
    world = dist.get_world_size(group)
    n = torch.tensor([packed.shape[0]], device=packed.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    pad = packed
    if packed.shape[0] < nmax:
        pad = torch.cat([packed, packed.new_zeros((nmax - packed.shape[0], packed.shape[1]))])
    out = packed.new_empty((world * nmax, packed.shape[1]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = [out[r * nmax: r * nmax + sizes[r]] for r in range(world)]
    return torch.cat(parts)


def moments(x, ok, group=None):
    """Mean and covariance of the converged best-fit parameters over ALL ranks
    (one all-reduce of count, sum x, sum x x^T)."""
    okf = ok.to(x.dtype)[:, None]
    xs = x * okf
    npar = x.shape[1]
    buf = torch.cat([okf.sum().reshape(1), xs.sum(dim=0), (xs.T @ x).reshape(-1)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    n = buf[0]
    m = buf[1:1 + npar] / n
    second = buf[1 + npar:].reshape(npar, npar) / n
    cov = (second - m[:, None] * m[None, :]) * (n / torch.clamp(n - 1, min=1))
    return m, cov, int(n.item())
