#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ from the REFERENCE ITSELF.

Run in the build container (where /root/reference is mounted and oracle/_ref has
been compiled by `make -C oracle ref`):

    python tests/golden/make_golden.py

For every fixture it writes <name>.jpg (a small baseline JPEG made with Pillow)
and <name>.npz holding what the compiled reference produces for that file:
  hdr_*        header fields as xjpeg_decode_header_ fills them (src/jpeg_wrap.c:263-319)
  quant        image.coef after xjpeg JPEG_DECODE_QUANT          (src/xjpeg.c:497-499,520-523,550-563)
  dct          image.coef after xjpeg JPEG_DECODE_DCT            (src/xjpeg.c:501-503,524-527)
  yuv          the padded planes after xjpeg JPEG_DECODE_YUV     (src/xjpeg.c:565-584, src/dct.c)
  rgb          the colour oracle (oracle_pipeline.c, res/yuv.fs.glsl:11-23) applied to `yuv`
  pack, index, packed   image.coef (the run/level words), image.index and image_plane.packed after
               xjpeg JPEG_DECODE_PACK                            (src/xjpeg.c:484-496,513-519,531-535)
plus blocks.npz: random int16 blocks and the reference's glj_real_idct8x8 output.
The GPU box has no /root/reference; tests read only these files.
"""
import io
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def picture(w, h, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 5 + yy * 3) % 256, (yy * 7 + xx) % 256, (xx * 2 + yy * 9) % 256], -1)
    return np.clip(base + rng.integers(-24, 25, size=base.shape), 0, 255).astype(np.uint8)


FIXTURES = [
    # name, (w, h), mode, pillow kwargs
    ("gray_48x40", (48, 40), "L", dict(quality=85)),
    ("gray_odd_37x21", (37, 21), "L", dict(quality=60)),
    ("c444_40x24", (40, 24), "RGB", dict(quality=85, subsampling=0)),
    ("c422_56x24", (56, 24), "RGB", dict(quality=85, subsampling=1)),
    ("c420_64x48", (64, 48), "RGB", dict(quality=85, subsampling=2)),
    ("c420_odd_70x50", (70, 50), "RGB", dict(quality=92, subsampling=2)),
    ("c420_opt_64x32", (64, 32), "RGB", dict(quality=75, subsampling=2, optimize=True)),
    ("c420_rst_80x48", (80, 48), "RGB", dict(quality=85, subsampling=2, restart_marker_blocks=2)),
    ("c422_rst_33x17", (33, 17), "RGB", dict(quality=50, subsampling=1, restart_marker_blocks=1)),
    ("c444_q100_24x24", (24, 24), "RGB", dict(quality=100, subsampling=0)),
    ("c420_q10_96x64", (96, 64), "RGB", dict(quality=10, subsampling=2)),
]


def main():
    ref = oracle.reference()
    for i, (name, (w, h), mode, kw) in enumerate(FIXTURES):
        img = picture(w, h, 1000 + i)
        pil = Image.fromarray(img[..., 0] if mode == "L" else img)
        bio = io.BytesIO()
        pil.save(bio, "JPEG", **kw)
        jpg = bio.getvalue()
        with open(os.path.join(HERE, name + ".jpg"), "wb") as f:
            f.write(jpg)
        hdr, g, quant = ref.ref_decode(jpg, "quant")
        _, _, dct = ref.ref_decode(jpg, "dct")
        _, _, yuv = ref.ref_decode(jpg, "yuv")
        planes = [yuv[p.data_off:p.data_off + p.width * p.height].reshape(p.height, p.width) for p in g.planes]
        # colour oracle on the reference's own planes (C implementation linked against the reference IDCT)
        rgb, planes2 = ref.decode_image(g, quant, hdr["qtabs"], hdr["tq"])
        assert all(np.array_equal(a, b) for a, b in zip(planes, planes2)), name
        assert np.array_equal(rgb, oracle.np_rgb_from_planes(g, planes)), name
        _, _, pack, index, packed = ref.ref_decode_pack(jpg)
        # the consumer restatement (res/horz_pack_yuv.fs.glsl:105-127) must give back the QUANT planes
        assert np.array_equal(ref.unpack_image(g, pack, index), quant), name
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            hdr_width=hdr["width"], hdr_height=hdr["height"], hdr_bits=hdr["bits"], hdr_ncomps=hdr["ncomps"],
            hdr_restart_interval=hdr["restart_interval"], hdr_hsamp=np.array(hdr["hsamp"]),
            hdr_vsamp=np.array(hdr["vsamp"]), hdr_tq=np.array(hdr["tq"]), hdr_qtabs=hdr["qtabs"],
            hdr_qvalid=np.array(hdr["qvalid"]), quant=quant, dct=dct, yuv=yuv, rgb=rgb.reshape(-1),
            pack=pack, index=index, packed=packed)
        print(f"{name}: {len(jpg)} B jpeg, {g.coef_len} coefs, subsamp h{hdr['hsamp']} v{hdr['vsamp']}")
    rng = np.random.default_rng(20261017)
    blocks = np.concatenate([
        rng.integers(-2048, 2048, size=(2000, 8, 8)),
        rng.integers(-300, 301, size=(1000, 8, 8)),
        rng.integers(-32768, 32768, size=(500, 8, 8)),
        np.zeros((1, 8, 8), dtype=np.int64),
    ]).astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "blocks.npz"), coef=blocks, idct=ref.ref_idct_blocks(blocks))
    print("blocks.npz:", blocks.shape)


if __name__ == "__main__":
    main()
