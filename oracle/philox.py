"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

numpy restatement of the counter-based generator the kernels use in throughput mode: Philox4x32-10 (Salmon, Moraes,
Dror, Shaw, SC'11; Random123) keyed by the 64-bit seed, counter = (global env id, step, call, 0).  The reference has
no counterpart (it draws with torch.rand over compacted index sets); this pins gr_fill_rand and, through
it, the in-kernel draws (tests/test_philox_chain.py)."""
from __future__ import annotations

import numpy as np

from generalizableracing_b200 import layout as L_

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [n,4] uint32, key [n,2] uint32 -> [n,4] uint32."""
    c = ctr.astype(np.uint64).copy()
    k0, k1 = key[:, 0].astype(np.uint64), key[:, 1].astype(np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[:, 0]
        p1 = np.uint64(M1) * c[:, 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = np.stack([(hi1 ^ c[:, 1] ^ k0) & mask, lo1, (hi0 ^ c[:, 3] ^ k1) & mask, lo0], axis=1)
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return c.astype(np.uint32)


def _call(env_ids, seed, step, call):
    n = len(env_ids)
    ctr = np.stack([np.asarray(env_ids, np.uint32), np.full(n, step, np.uint32), np.full(n, call, np.uint32), np.zeros(n, np.uint32)], axis=1)
    key = np.stack([np.full(n, seed & 0xFFFFFFFF, np.uint32), np.full(n, (seed >> 32) & 0xFFFFFFFF, np.uint32)], axis=1)
    return philox4x32_10(ctr, key)


def rnd_rows(env_ids, seed: int, step: int) -> np.ndarray:
    """The dense per-step random tensor rnd[n, RND_STRIDE] of the boundary (layout.py), float32."""
    n = len(env_ids)
    out = np.zeros((n, L_.RND_STRIDE), np.float32)
    x = _call(env_ids, seed, step, 0)                     # slots 0..7: one Box-Muller pair per 32-bit word
    u1 = ((x & 0xFFFF).astype(np.float32) + np.float32(1)) * np.float32(2.0 ** -16)
    u2 = (x >> 16).astype(np.float32) * np.float32(2.0 ** -16)
    r = np.sqrt(np.float32(-1.3862943611198906) * np.log2(u1))
    ang = (u2 - np.float32(0.5)) * np.float32(6.283185307179586)
    out[:, 0:8:2] = r * np.cos(ang)
    out[:, 1:8:2] = r * np.sin(ang)
    for call in range(2, L_.RND_STRIDE // 4):
        w = _call(env_ids, seed, step, call)
        out[:, 4 * call:4 * call + 4] = (w >> 8).astype(np.float32) * np.float32(2.0 ** -24)
    return out
