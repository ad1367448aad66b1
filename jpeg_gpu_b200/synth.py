"""Synthetic inputs for tests and benchmarks.

Quantisation tables: JPEG Annex-K luma/chroma scaled to quality 85 by the
libjpeg rule, natural order.  Coefficients per block: DC ~ N(0, 300/q0);
AC at zig-zag position k is non-zero with probability 0.9*exp(-k/8), magnitude
1+floor(Exp(24*exp(-k/12)/q_k)), random sign; everything clipped so that
|coef*q| <= 2047 (the IEEE-1180 input range, test/dct.c:136).  Blocks are
stored in the reference's image.coef layout (src/xjpeg.c:550-563).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .batch import ImageDesc

SEED_BASE = 20261017

ANNEX_K_LUMA = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64)
ANNEX_K_CHROMA = np.array([
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99], dtype=np.int64)

# zig-zag position -> natural index (ITU-T T.81 figure A.6)
NATURAL = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5,
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63], dtype=np.int64)
ZIGZAG_OF_NATURAL = np.argsort(NATURAL)


def quality_tables(quality: int = 85) -> np.ndarray:
    """(4, 64) uint16: slot 0 luma, slot 1 chroma, slots 2-3 copies (libjpeg's
    jpeg_quality_scaling + jpeg_add_quant_table, baseline-clamped)."""
    scale = 5000 // quality if quality < 50 else 200 - 2 * quality
    out = np.zeros((4, 64), dtype=np.uint16)
    for slot, base in enumerate((ANNEX_K_LUMA, ANNEX_K_CHROMA, ANNEX_K_LUMA, ANNEX_K_CHROMA)):
        out[slot] = np.clip((base * scale + 50) // 100, 1, 255)
    return out


def _plane_blocks(rng: np.random.Generator, n: int, q: np.ndarray, kind: str) -> np.ndarray:
    """n blocks (n, 64) int16 in natural order for table q (64,)."""
    q = q.astype(np.int64)
    lim = 2047 // q
    blk = np.zeros((n, 64), dtype=np.int64)
    if kind == "natural":
        blk[:, 0] = np.rint(rng.normal(0.0, 300.0 / q[0], size=n))
        for k in range(1, 64):
            nat = NATURAL[k]
            on = rng.random(n) < 0.9 * np.exp(-k / 8.0)
            mag = 1 + np.floor(rng.exponential(24.0 * np.exp(-k / 12.0) / q[nat], size=n))
            sign = np.where(rng.random(n) < 0.5, -1, 1)
            blk[:, nat] = np.where(on, mag * sign, 0)
    elif kind == "dense":      # full +-2047/q range everywhere: forces clamping at 0 and 255
        blk = rng.integers(-lim, lim + 1, size=(n, 64))
    elif kind == "dc":
        blk[:, 0] = rng.integers(-lim[0], lim[0] + 1, size=n)
    elif kind == "zero":
        pass
    elif kind == "impulse":    # one coefficient per block, cycling through the 64 positions
        pos = np.arange(n) % 64
        val = rng.integers(-lim[pos], lim[pos] + 1)
        blk[np.arange(n), pos] = val
    elif kind == "int16":      # anything a short can hold (dequantisation wraps)
        blk = rng.integers(-32768, 32768, size=(n, 64))
        return blk.astype(np.int16)
    else:
        raise ValueError(kind)
    return np.clip(blk, -lim, lim).astype(np.int16)


def image_coefficients(desc: ImageDesc, qtabs: np.ndarray, seed: int, kind: str = "natural") -> np.ndarray:
    """One image's coefficient buffer (layout.coef_len int16) in the reference layout."""
    lay = desc.query_layout()
    rng = np.random.default_rng(seed)
    out = np.zeros(lay.coef_len, dtype=np.int16)
    for p, t in zip(lay.planes, desc.tq):
        n = p.hblocks * p.vblocks
        out[p.coef_off:p.coef_off + n * 64] = _plane_blocks(rng, n, qtabs[t], kind).reshape(-1)
    return out


def batch_coefficients(descs: List[ImageDesc], coef_len: int, qtabs: np.ndarray, kinds: Sequence[str] = ("natural",),
                       first_index: int = 0) -> np.ndarray:
    """Whole-batch buffer; image i uses seed SEED_BASE + first_index + i."""
    buf = np.zeros(coef_len, dtype=np.int16)
    for i, d in enumerate(descs):
        c = image_coefficients(d, qtabs[d.qtab_set] if qtabs.ndim == 3 else qtabs, SEED_BASE + first_index + i,
                               kinds[i % len(kinds)])
        buf[d.coef_off:d.coef_off + c.size] = c
    return buf


def torch_batch_coefficients(descs: List[ImageDesc], coef_len: int, qtabs: np.ndarray, device, first_index: int = 0):
    """Same distribution generated on the GPU (bench inputs are GB-sized).
    Statistically identical to batch_coefficients('natural'); a different
    random stream, seeded per image with SEED_BASE + first_index + i."""
    import torch
    buf = torch.zeros(coef_len, dtype=torch.int16, device=device)
    ks = torch.arange(64, device=device, dtype=torch.float32)
    zz = torch.as_tensor(ZIGZAG_OF_NATURAL, device=device, dtype=torch.float32)  # zig-zag index of each natural slot
    p_on = 0.9 * torch.exp(-zz / 8.0)
    cache = {}
    for i, d in enumerate(descs):
        lay = d.query_layout()
        g = torch.Generator(device=device)
        g.manual_seed(SEED_BASE + first_index + i)
        qset = qtabs[d.qtab_set] if qtabs.ndim == 3 else qtabs
        for p, t in zip(lay.planes, d.tq):
            n = p.hblocks * p.vblocks
            key = int(t)
            if key not in cache:
                q = torch.as_tensor(qset[t].astype(np.float32), device=device)
                cache[key] = (q, torch.floor(2047.0 / q), 24.0 * torch.exp(-zz / 12.0) / q)
            q, lim, scale = cache[key]
            u = torch.rand((n, 64), generator=g, device=device)
            e = -torch.log1p(-torch.rand((n, 64), generator=g, device=device)) * scale
            sgn = torch.where(torch.rand((n, 64), generator=g, device=device) < 0.5, -1.0, 1.0)
            blk = torch.where(u < p_on, (1 + torch.floor(e)) * sgn, torch.zeros((), device=device))
            blk[:, 0] = torch.round(torch.randn(n, generator=g, device=device) * (300.0 / q[0]))
            blk = torch.minimum(torch.maximum(blk, -lim), lim)
            buf[d.coef_off + p.coef_off: d.coef_off + p.coef_off + n * 64] = blk.to(torch.int16).reshape(-1)
            del u, e, sgn, blk
    del ks
    return buf
