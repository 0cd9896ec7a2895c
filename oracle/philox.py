"""TEST INFRASTRUCTURE ONLY.  Bit-exact numpy mirror of csrc/philox.cuh.

Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11 --
the published Random123 algorithm; known-answer vectors from Random123's kat_vectors are checked in
tests/test_oracle_cpu.py).  Counter = (env_id, step_lo, step_hi, stream), key = (seed_lo, seed_hi).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

# stream ids (must match csrc/philox.cuh)
RS_STEP_A, RS_STEP_B = 0, 1
RS_RESET = [2, 3, 4, 5, 6, 7, 8, 9, 10]
RS_ROWS = 64
RS_RESET_COM = 11
RS_OBST = 0x1000      # + round*8 + j//2 ; obstacle j takes words (0,1) if j even else (2,3)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable uint32-valued arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint64(int(k0) & 0xFFFFFFFF)
    k1 = np.uint64(int(k1) & 0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ k0
        n2 = hi0 ^ c3 ^ k1
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = np.uint64((int(k0) + W0) & 0xFFFFFFFF)
        k1 = np.uint64((int(k1) + W1) & 0xFFFFFFFF)
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def u01(x):
    """24-bit uniform in [0,1): (x >> 8) * 2**-24, exactly representable in fp32."""
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def uniform4(seed: int, env_ids, step: int, stream: int) -> np.ndarray:
    """(len(env_ids), 4) float32 uniforms for Philox counter (env, step, stream)."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    lo = env_ids & MASK
    hi = (env_ids >> np.uint64(32)) & MASK
    c3 = (np.uint64(stream) ^ (hi << np.uint64(8))) & MASK
    r = philox4x32_10(lo, np.uint64(step & 0xFFFFFFFF), np.uint64((step >> 32) & 0xFFFFFFFF), c3,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack([u01(x) for x in r], axis=-1)


def uniform8x16(seed: int, env_ids, step: int, stream: int) -> np.ndarray:
    """(len(env_ids), 8) float32: the 16-bit halves (lo, hi) of the four Philox words, each * 2**-16."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    lo = env_ids & MASK
    hi = (env_ids >> np.uint64(32)) & MASK
    c3 = (np.uint64(stream) ^ (hi << np.uint64(8))) & MASK
    r = philox4x32_10(lo, np.uint64(step & 0xFFFFFFFF), np.uint64((step >> 32) & 0xFFFFFFFF), c3,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    cols = []
    for w in r:
        cols.append((w & np.uint32(0xFFFF)).astype(np.float32) * np.float32(1.0 / 65536.0))
        cols.append((w >> np.uint32(16)).astype(np.float32) * np.float32(1.0 / 65536.0))
    return np.stack(cols, axis=-1).astype(np.float32)
