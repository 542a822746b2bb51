"""Philox4x32-10 and the normal stream of csrc/bootstrap.cu, restated in numpy -- test infrastructure.

The reference draws its bootstrap / simulated copies with gvar's generator (numpy's global RNG
behind ``gvar.bootstrap_iter`` / ``gvar.raniter``; call sites src/lsqfit/__init__.py:1532-1535,
1615-1624), which cannot be reproduced bit for bit on a device and is not part of the fit's parity
contract (only the DISTRIBUTION mean + L z matters).  The device uses the counter-based Philox4x32-10
of Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3" (SC'11); this module
pins that algorithm against the paper's published known-answer vectors (tests/test_cabi_cpu.py) and
gives the GPU tests a host-side stream to compare with.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32 array, key: (2,) ints.  Returns the (..., 4) uint32 output block."""
    c = np.asarray(ctr, dtype=np.uint64).copy()
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[..., 0]
        p1 = M1 * c[..., 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c[..., 1] ^ np.uint64(k0)
        n2 = hi0 ^ c[..., 3] ^ np.uint64(k1)
        c = np.stack([n0, lo1, n2, lo0], axis=-1)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def normals(first, count, seed):
    """Elements first .. first+count-1 of the device's normal stream (csrc/bootstrap.cu:
    normals_kernel): element g is normal (g & 1) of pair (g >> 1)."""
    p0, p1 = first >> 1, (first + count - 1) >> 1
    pair = np.arange(p0, p1 + 1, dtype=np.uint64)
    ctr = np.zeros((pair.size, 4), dtype=np.uint64)
    ctr[:, 0] = pair & MASK
    ctr[:, 1] = pair >> np.uint64(32)
    w = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).astype(np.uint64)
    u1 = ((((w[:, 0] >> np.uint64(5)) << np.uint64(26)) | (w[:, 1] >> np.uint64(6))).astype(np.float64) + 0.5) * 2.0 ** -53
    u2 = ((((w[:, 2] >> np.uint64(5)) << np.uint64(26)) | (w[:, 3] >> np.uint64(6))).astype(np.float64) + 0.5) * 2.0 ** -53
    r = np.sqrt(-2.0 * np.log(u1))
    z = np.empty(2 * pair.size)
    z[0::2] = r * np.cos(2.0 * np.pi * u2)
    z[1::2] = r * np.sin(2.0 * np.pi * u2)
    off = first - 2 * p0
    return z[off:off + count], w.astype(np.uint32)
