"""Oracle (test infrastructure): Philox4x32-10 and the kernels' noise layout in numpy.

Philox4x32-10 is the published counter-based generator of Salmon, Moraes, Dror & Shaw (SC'11) -- the same
round function and constants as Random123 / cuRAND / torch.  Known-answer vectors from the Random123
distribution (kat_vectors) are checked in tests/test_oracle_misc.py.

Layout restated from nfmc_b200/csrc/common.cuh ("Counter convention"): for chain c, step t, stream s and the
lane group size gs chosen for event size d,
    counter = (quad*32 + j, s | (t >> 32) << 8, t & 0xffffffff, c),   key = (seed lo, seed hi)
    pair p = 0           -> word 0 of quad 0 on lane j = 0 is the accept uniform  u = (w >> 8) * 2^-24
    pair p = e + 1       -> Box-Muller on words (2(p%2), 2(p%2)+1) of quad p//2  -> (lo[e], hi[e])
with lo[e] = element j + gs*e of the low half [0, d//2) and hi[e] = element d//2 + j + gs*e.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = np.broadcast_to(key[..., 0], c[0].shape).astype(np.uint32).copy()
    k1 = np.broadcast_to(key[..., 1], c[0].shape).astype(np.uint32).copy()
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0.astype(np.uint64), lo1, hi0 ^ c[3] ^ k1.astype(np.uint64), lo0]
        with np.errstate(over="ignore"):
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack([v.astype(np.uint32) for v in c], axis=-1)


def layout_for_dim(d: int):
    """(gs, E) exactly as nfmc_b200/csrc/host_common.cuh: layout_for_dim."""
    db = d - d // 2
    gs = 1
    while (db + gs - 1) // gs > 16:
        gs *= 2
    e = (db + gs - 1) // gs
    E = 4 if e <= 4 else 7 if e <= 7 else 13 if e <= 13 else 16
    return gs, E


def box_muller(a: np.ndarray, b: np.ndarray):
    """Exact-arithmetic version of common.cuh: box_muller (the kernel uses MUFU approximations, ~1e-6)."""
    u1 = a.astype(np.float64) * 2.0 ** -32 + 2.0 ** -33
    frac = (b >> np.uint32(9)).astype(np.float64) * 2.0 ** -23        # mantissa of [1,2) float minus 1
    th = (frac - 0.5) * 2.0 * np.pi
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(th), r * np.sin(th)


def step_noise(seed: int, stream: int, step: int, chain0: int, n: int, d: int):
    """Normals [n, d] (float64) and accept uniforms [n] (float32) for one step, as the kernels draw them."""
    gs, _ = layout_for_dim(d)
    da, db = d // 2, d - d // 2
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    chains = (np.arange(n, dtype=np.uint64) + np.uint64(chain0)).astype(np.uint32)
    cy = np.uint32((stream | ((step >> 32) << 8)) & 0xFFFFFFFF)
    cz = np.uint32(step & 0xFFFFFFFF)
    normals = np.zeros((n, d), dtype=np.float64)

    def quad(q, j):
        ctr = np.zeros((n, 4), dtype=np.uint32)
        ctr[:, 0] = np.uint32(q * 32 + j)
        ctr[:, 1] = cy
        ctr[:, 2] = cz
        ctr[:, 3] = chains
        return philox4x32_10(ctr, key)

    uniforms = ((quad(0, 0)[:, 0] >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    n_slots = (db + gs - 1) // gs
    for j in range(gs):
        cache = {}
        for e in range(n_slots):
            k = j + gs * e
            p = e + 1
            q = p // 2
            if q not in cache:
                cache[q] = quad(q, j)
            w = cache[q]
            z0, z1 = box_muller(w[:, 2 * (p % 2)], w[:, 2 * (p % 2) + 1])
            if k < da:
                normals[:, k] = z0
            if k < db:
                normals[:, da + k] = z1
    return normals, uniforms


def scalar_uniforms(seed: int, stream: int, step: int, chain0: int, n: int, count: int):
    """``count`` per-chain uniforms [n, count] (float32) for one step: word i % 4 of counter quad i // 4 on lane j = 0 --
    the layout the elliptical-slice kernel uses for {u, theta0, bracket draws} with stream = 2 (csrc/ess_kernel.cu)."""
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    chains = (np.arange(n, dtype=np.uint64) + np.uint64(chain0)).astype(np.uint32)
    out = np.zeros((n, count), dtype=np.float32)
    for q in range((count + 3) // 4):
        ctr = np.zeros((n, 4), dtype=np.uint32)
        ctr[:, 0] = np.uint32(q * 32)
        ctr[:, 1] = np.uint32((stream | ((step >> 32) << 8)) & 0xFFFFFFFF)
        ctr[:, 2] = np.uint32(step & 0xFFFFFFFF)
        ctr[:, 3] = chains
        w = philox4x32_10(ctr, key)
        for i in range(4):
            if 4 * q + i < count:
                out[:, 4 * q + i] = (w[:, i] >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return out
