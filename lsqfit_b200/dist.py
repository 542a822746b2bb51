"""Multi-GPU sharding of a batch of independent fits (one process per GPU).

Fits are independent, so the batch index range is cut into contiguous shards (gathered
output order == input order) and the fitting itself needs no collective.  NCCL (over
NVLink/NVSwitch) is used only after the fit: an all-gather of the packed per-fit results
and an all-reduce of first/second moments for bootstrap averages.
The same code runs on the gloo backend with CPU tensors for the world_size-2 tests.
"""
import ctypes as C

import torch
import torch.distributed as dist


class Comm(object):
    """NCCL communicator behind the C ABI (b200lm_comm_init / b200lm_gather / b200lm_allreduce_sum, csrc/comm.cu).
    ``torch.distributed`` is used only to ship the 128-byte NCCL id from rank 0 to the others (any host channel
    would do); the collectives on the data path are the library's own calls."""

    def __init__(self, device, rank=None, world=None, group=None):
        from . import _cabi
        self._cabi = _cabi
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        ident = C.create_string_buffer(128)
        if self.rank == 0:
            _cabi.check(_cabi.lib.b200lm_comm_unique_id(ident))
        obj = [ident.raw]
        dist.broadcast_object_list(obj, src=0, group=group)
        self._h = C.c_void_p()
        self.device = int(device)
        _cabi.check(_cabi.lib.b200lm_comm_init(self.device, self.rank, self.world, obj[0], C.byref(self._h)))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream)

    def gather(self, t):
        """rows of every rank, in rank order (equal shards)"""
        t = t.contiguous()
        out = t.new_empty((self.world * t.shape[0],) + tuple(t.shape[1:]))
        self._cabi.check(self._cabi.lib.b200lm_gather(self._h, t.data_ptr(), out.data_ptr(), t.numel(), self._stream()))
        return out

    def allreduce_sum(self, t):
        assert t.is_contiguous() and t.dtype == torch.float64
        self._cabi.check(self._cabi.lib.b200lm_allreduce_sum(self._h, t.data_ptr(), t.numel(), self._stream()))
        return t

    def close(self):
        if self._h:
            self._cabi.lib.b200lm_comm_destroy(self._h)
            self._h = C.c_void_p()


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


def gather_results(packed, B_total=None, group=None, comm=None):
    """All-gather the packed results of every rank in rank order: ONE collective, no host
    synchronisation.  Shards follow ``shard_range`` (sizes differ by at most one row), so every
    rank can compute all shard sizes from ``B_total`` alone; without ``B_total`` the shards are
    taken to be equal (the weak-scaling case).  Smaller shards are padded to the largest."""
    if not (dist.is_available() and dist.is_initialized()):
        return packed
    world = dist.get_world_size(group)
    n = packed.shape[0]
    if B_total is None:
        sizes = [n] * world
    else:
        sizes = [shard_range(B_total, r, world)[1] - shard_range(B_total, r, world)[0] for r in range(world)]
        if sizes[dist.get_rank(group)] != n:
            raise ValueError("shard of %d rows does not match shard_range(B_total=%d)" % (n, B_total))
    nmax = max(sizes)
    pad = packed
    if n < nmax:
        pad = torch.cat([packed, packed.new_zeros((nmax - n, packed.shape[1]))])
    if comm is not None:
        out = comm.gather(pad)                                   # b200lm_gather (C ABI)
    else:
        out = packed.new_empty((world * nmax, packed.shape[1]))
        dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if min(sizes) == nmax:
        return out
    return torch.cat([out[r * nmax: r * nmax + sizes[r]] for r in range(world)])


def moments(x, ok, group=None, comm=None):
    """Mean and covariance of the converged best-fit parameters over ALL ranks
    (one all-reduce of count, sum x, sum x x^T)."""
    okb = ok.to(torch.bool)[:, None]
    # select, do not multiply: a failed fit may hold NaN/Inf, and 0 * NaN would poison every rank
    xs = torch.where(okb, x, torch.zeros_like(x))
    npar = x.shape[1]
    buf = torch.cat([okb.sum().to(x.dtype).reshape(1), xs.sum(dim=0), (xs.T @ xs).reshape(-1)])
    if comm is not None:
        comm.allreduce_sum(buf)                                  # b200lm_allreduce_sum (C ABI)
    elif dist.is_available() and dist.is_initialized():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    n = buf[0]
    m = buf[1:1 + npar] / n
    second = buf[1 + npar:].reshape(npar, npar) / n
    cov = (second - m[:, None] * m[None, :]) * (n / torch.clamp(n - 1, min=1))
    return m, cov, int(n.item())
