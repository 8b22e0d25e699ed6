"""Host twin of the in-kernel Philox4x32-10 generator (l2hmc_b200/csrc/common.cuh).

The reference draws momentum, direction bits and accept uniforms from unseeded TF RNGs
(utils/dynamics.py:248,276; utils/sampler.py:34,54).  Here every draw is a pure function of
(seed, call counter, global chain id), so a test (or another process holding a different shard of the
chains) can regenerate exactly the numbers the kernel used:

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (chain_lo, chain_hi, block, (call_counter << 2) | stream)
    stream 0: momentum -- block b gives dims 4b..4b+3, Box-Muller on word pairs (0,1) and (2,3)
    stream 1: word0 & 1 = direction bit (1 = forward); (word1 >> 8) * 2^-24 = accept uniform
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10. c* are uint32 arrays (same shape); k0/k1 python ints."""
    c0 = c0.astype(np.uint64)
    c1 = c1.astype(np.uint64)
    c2 = c2.astype(np.uint64)
    c3 = c3.astype(np.uint64)
    for r in range(10):
        kk0 = np.uint64((k0 + r * _W0) & 0xFFFFFFFF)
        kk1 = np.uint64((k1 + r * _W1) & 0xFFFFFFFF)
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ kk0) & _MASK, lo1, (hi0 ^ c3 ^ kk1) & _MASK, lo0
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def _words(seed: int, counter: int, chains: np.ndarray, block, stream: int):
    chains = np.asarray(chains, dtype=np.uint64)
    c0 = (chains & _MASK).astype(np.uint32)
    c1 = (chains >> np.uint64(32)).astype(np.uint32)
    c2 = np.broadcast_to(np.asarray(block, dtype=np.uint32), c0.shape)
    c3 = np.full(c0.shape, ((int(counter) << 2) | stream) & 0xFFFFFFFF, dtype=np.uint32)
    return philox4x32_10(c0, c1, c2, c3, int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)


def _box_muller(a, b):
    scale = np.float32(5.9604644775390625e-08)
    u1 = ((a >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * scale
    u2 = ((b >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * scale
    r = np.sqrt(np.float32(-2.0) * np.log(u1), dtype=np.float32)
    ang = np.float32(6.2831855) * u2
    return (r * np.cos(ang, dtype=np.float32)).astype(np.float32), (r * np.sin(ang, dtype=np.float32)).astype(np.float32)


def normals(seed: int, counter: int, n: int, d: int, chain_offset: int = 0) -> np.ndarray:
    """Momentum v [n, d] as the kernel draws it (fp32; equal to the device up to libm ulps)."""
    chains = np.arange(chain_offset, chain_offset + n, dtype=np.uint64)
    nb = (d + 3) // 4
    out = np.empty((n, nb * 4), dtype=np.float32)
    for b in range(nb):
        w = _words(seed, counter, chains, b, 0)
        out[:, 4 * b + 0], out[:, 4 * b + 1] = _box_muller(w[0], w[1])
        out[:, 4 * b + 2], out[:, 4 * b + 3] = _box_muller(w[2], w[3])
    return np.ascontiguousarray(out[:, :d])


def direction_and_uniform(seed: int, counter: int, n: int, chain_offset: int = 0):
    """(dir uint8 [n], u float32 [n]) exactly as the kernel draws them (bit-exact)."""
    chains = np.arange(chain_offset, chain_offset + n, dtype=np.uint64)
    w = _words(seed, counter, chains, 0, 1)
    d = (w[0] & np.uint32(1)).astype(np.uint8)
    u = (w[1] >> np.uint32(8)).astype(np.float32) * np.float32(5.9604644775390625e-08)
    return d, u
