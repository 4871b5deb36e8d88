"""CPU emulation of the tensor-core flow kernel's arithmetic, driven by the PACKED blob (test infrastructure).

Decodes the blob exactly as csrc/tc_common.cuh addresses it -- UMMA K-major images ``[K/8][rows][8]``, the constant-one
bias columns of GEMM 1, the K-step layout of the hidden activations in tensor memory (step s = hidden units
[8s, 8s+8) and [Hp/2 + 8s, Hp/2 + 8s + 8), second k-group fetched ``Hp/16`` k-groups further on), the folded
log2(e) / 1/2 / log(1-m) constants -- with bf16 rounding where the kernel rounds.  A packing or addressing mistake shows
up here, without a GPU, as a mismatch against the oracle flow.
"""
import math

import torch

MIN_SCALE = 1e-3


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def decode_blob(blob_u8: torch.Tensor, d: int, Lc: int, Hp: int):
    da = d // 2
    n2p = ((2 * (d - da) + 15) // 16) * 16
    k1 = ((da + 2 + 15) // 16) * 16
    n_aff = (Lc + 1) * 4 * d + 4
    aff = blob_u8[: n_aff * 4].view(torch.float32)
    off = n_aff * 4
    cps = []
    for _ in range(Lc):
        w1 = blob_u8[off: off + k1 * Hp * 2].view(torch.bfloat16).to(torch.float32).reshape(k1 // 8, Hp, 8)
        off += k1 * Hp * 2
        wl = blob_u8[off: off + Hp * n2p * 2].view(torch.bfloat16).to(torch.float32).reshape(Hp // 8, n2p, 8)
        off += Hp * n2p * 2
        bl = blob_u8[off: off + n2p * 4].view(torch.float32)
        off += n2p * 4
        cps.append((w1, wl, bl))
    assert off == blob_u8.numel()
    return aff, cps, (da, n2p, k1)


def run_pass(blob_u8, d, Lc, Hp, x, inverse: bool):
    """Returns (y, log_det) as the kernel computes them (physical coordinates handled as the kernel does)."""
    aff, cps, (da, n2p, k1) = decode_blob(blob_u8, d, Lc, Hp)
    n = x.shape[0]
    flip = Lc % 2 == 1
    src = x.flip(1) if (inverse and flip) else x
    lo, hi = src[:, :da].clone(), src[:, da:].clone()
    ld2 = torch.zeros(n)
    log_const = float(aff[(Lc + 1) * 4 * d])

    def affine(idx):
        nonlocal lo, hi
        base = idx * 4 * d + (2 * d if inverse else 0)
        tab = aff[base: base + 2 * d].reshape(d, 2)
        lo = tab[:da, 0] * lo + tab[:da, 1]
        hi = tab[da:, 0] * hi + tab[da:, 1]

    affine(Lc if inverse else 0)
    for i in range(Lc):
        l = Lc - 1 - i if inverse else i
        w1, wl, bl = cps[l]
        src_hi = l % 2 == 0
        s = hi if src_hi else lo
        # GEMM 1: A = [S | 1 | 1 | 0...] in bf16, B = W1 image
        a1 = torch.zeros(n, k1)
        a1[:, :da] = _bf16(s)
        a1[:, da] = 1.0
        a1[:, da + 1] = 1.0
        w1m = w1.permute(1, 0, 2).reshape(Hp, k1)              # [h][k]
        hpre = a1 @ w1m.t()
        hid = _bf16(torch.tanh(hpre))                           # tanh.approx.f32 of the fp32 accumulator, then bf16
        # GEMM 2 in the kernel's K-step order: step s multiplies hidden units [8s,8s+8) with k-group s of the image and
        # [Hp/2+8s, ...) with k-group s + Hp/16
        u = torch.zeros(n, n2p)
        ks = Hp // 16
        for st in range(ks):
            for half, kg in ((0, st), (1, st + ks)):
                hsel = hid[:, (Hp // 2) * half + 8 * st: (Hp // 2) * half + 8 * st + 8]
                u += hsel @ wl[kg].t()                         # wl[kg]: [n2p][8]
        u = u + bl
        ua, ub = u[:, 0::2][:, :da], u[:, 1::2][:, :da]
        alpha = torch.exp2(ua) + MIN_SCALE
        if inverse:
            if src_hi:
                lo = (lo - ub) / alpha
            else:
                hi = (hi - ub) / alpha
        else:
            if src_hi:
                lo = alpha * lo + ub
            else:
                hi = alpha * hi + ub
        ld2 = ld2 + torch.log2(alpha).sum(1)
        affine(l if inverse else l + 1)
    y = torch.cat([lo, hi], dim=1)
    if (not inverse) and flip:
        y = y.flip(1)
    ld = ld2 * math.log(2.0) + log_const
    return y, (-ld if inverse else ld)
