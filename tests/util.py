"""Shared helpers for the parity tests."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

import jpeg_gpu_b200 as J
import oracle


def make_batch(shapes: Sequence[tuple], want_yuv: bool = True, n_sets: int = 1):
    """shapes: (width, height, subsampling-name).  Returns descs + buffer sizes."""
    descs = []
    for i, (w, h, ss) in enumerate(shapes):
        hs, vs = J.SUBSAMPLINGS[ss]
        descs.append(J.ImageDesc(w, h, hs, vs, tq=(0, 1, 1)[:len(hs)], qtab_set=i % n_sets))
    coef_len, rgb_len, yuv_len = J.pack_batch(descs, want_yuv=want_yuv)
    return descs, coef_len, rgb_len, yuv_len


def oracle_batch(lib: oracle.OracleLib, descs: List[J.ImageDesc], coef: np.ndarray, qtabs: np.ndarray,
                 rgb_len: int, yuv_len: int, nthreads: int = 4):
    """Runs the CPU oracle over the same batch buffers."""
    rows = []
    for d in descs:
        g = oracle.geometry(d.width, d.height, d.hsamp, d.vsamp)
        rows.append(oracle.make_desc(g, d.tq, d.coef_off, d.rgb_off, d.qtab_set, d.yuv_off))
    q = qtabs.reshape(-1, 4, 64)
    return lib.decode_batch(np.stack(rows), coef, q, rgb_len, yuv_len, nthreads)


def gpu_batch(ctx: J.Context, descs: List[J.ImageDesc], coef: np.ndarray, qtabs: np.ndarray, rgb_len: int,
              yuv_len: int, force_generic: bool = False, want_rgb: bool = True):
    """Device-resident path: torch tensors in, numpy out."""
    import torch
    dev = torch.device("cuda", ctx.device)
    d_coef = torch.from_numpy(coef).to(dev)
    d_q = torch.from_numpy(qtabs.astype(np.int16).reshape(-1)).to(dev)  # bit pattern of the uint16 tables
    d_rgb = torch.full((max(rgb_len, 16),), 0xAB, dtype=torch.uint8, device=dev) if want_rgb else None
    d_yuv = torch.full((max(yuv_len, 16),), 0xCD, dtype=torch.uint8, device=dev) if yuv_len else None
    plan = ctx.plan(descs, rgb=want_rgb, yuv=bool(yuv_len), force_generic=force_generic)
    plan.run(d_coef, d_q, d_rgb, d_yuv)
    torch.cuda.synchronize(dev)
    plan.close()
    return (d_rgb.cpu().numpy() if want_rgb else None), (d_yuv.cpu().numpy() if yuv_len else None)


def compare_batch(descs, got_rgb, got_yuv, exp_rgb, exp_yuv):
    """Bit-exact comparison restricted to the bytes each image owns."""
    for i, d in enumerate(descs):
        lay = d.query_layout()
        if got_yuv is not None and d.yuv_off >= 0:
            a = got_yuv[d.yuv_off:d.yuv_off + lay.data_len]
            b = exp_yuv[d.yuv_off:d.yuv_off + lay.data_len]
            if not np.array_equal(a, b):
                bad = np.nonzero(a != b)[0]
                raise AssertionError(f"image {i} ({d.width}x{d.height} {d.hsamp}/{d.vsamp}): {bad.size} plane bytes differ, "
                                     f"first at {bad[0]}: got {a[bad[0]]} want {b[bad[0]]}")
        if got_rgb is not None:
            a = got_rgb[d.rgb_off:d.rgb_off + lay.rgb_len]
            b = exp_rgb[d.rgb_off:d.rgb_off + lay.rgb_len]
            if not np.array_equal(a, b):
                bad = np.nonzero(a != b)[0]
                raise AssertionError(f"image {i} ({d.width}x{d.height} {d.hsamp}/{d.vsamp}): {bad.size} rgb bytes differ, "
                                     f"first at {bad[0]}: got {a[bad[0]]} want {b[bad[0]]}")
