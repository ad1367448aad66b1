#!/usr/bin/env python
"""Generates tests/golden/big/: the two BASELINE-sized files, from the REFERENCE ITSELF.

    python tests/golden/make_golden_big.py       (build container: /root/reference mounted,
                                                  oracle/_ref built by `make -C oracle ref`)

  gray_512x512   BASELINE config 1: one 512x512 baseline grayscale JPEG (Pillow, mode L, q=85).
                 .npz holds everything make_golden.py stores for the small fixtures.
  c420_1920x1080 BASELINE config 2's shape as a real file (Pillow, subsampling=2, q=85).
                 To keep the repository small the .npz holds the header fields plus SHA-256
                 digests of the reference's QUANT planes, YUV planes and of the colour oracle's
                 pixels; tests hash what they produced (bit-exactness needs no more) and, where
                 oracle/_ref is present, also compare against the live reference.
These live in a sub-directory so that the per-fixture CPU tests (host emulation of the device
entropy decoder etc.) keep iterating over the small files only.
"""
import hashlib
import io
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

BIG = os.path.join(HERE, "big")


def picture(w, h, seed, noise):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 96 * np.sin(xx / 37.0) * np.cos(yy / 23.0), (yy * 3 + xx) % 256 * 0.6 + 40,
                     128 + 100 * np.sin((xx + 2 * yy) / 61.0)], -1)
    # a few hard edges so that clamping at 0/255 and large AC terms occur
    base[(xx // 64 + yy // 48) % 5 == 0] = 250
    base[(xx // 96 + yy // 80) % 7 == 0] = 3
    return np.clip(base + rng.integers(-noise, noise + 1, size=base.shape), 0, 255).astype(np.uint8)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = oracle.reference()
    os.makedirs(BIG, exist_ok=True)
    for name, (w, h), mode, kw, full in [
        ("gray_512x512", (512, 512), "L", dict(quality=85), True),
        ("c420_1920x1080", (1920, 1080), "RGB", dict(quality=85, subsampling=2), False),
    ]:
        img = picture(w, h, 77 + w, 5)
        pil = Image.fromarray(img[..., 0] if mode == "L" else img)
        bio = io.BytesIO()
        pil.save(bio, "JPEG", **kw)
        jpg = bio.getvalue()
        with open(os.path.join(BIG, name + ".jpg"), "wb") as f:
            f.write(jpg)
        hdr, g, quant = ref.ref_decode(jpg, "quant")
        _, _, yuv = ref.ref_decode(jpg, "yuv")
        rgb, planes = ref.decode_image(g, quant, hdr["qtabs"], hdr["tq"], nthreads=8)
        assert np.array_equal(np.concatenate([p.ravel() for p in planes]), yuv), name
        _, _, pack, index, packed = ref.ref_decode_pack(jpg)
        assert np.array_equal(ref.unpack_image(g, pack, index), quant), name
        common = dict(hdr_width=hdr["width"], hdr_height=hdr["height"], hdr_bits=hdr["bits"], hdr_ncomps=hdr["ncomps"],
                      hdr_restart_interval=hdr["restart_interval"], hdr_hsamp=np.array(hdr["hsamp"]),
                      hdr_vsamp=np.array(hdr["vsamp"]), hdr_tq=np.array(hdr["tq"]), hdr_qtabs=hdr["qtabs"],
                      hdr_qvalid=np.array(hdr["qvalid"]), packed=packed,
                      sha_quant=sha(quant), sha_yuv=sha(yuv), sha_rgb=sha(rgb.reshape(-1)), sha_pack=sha(pack))
        if full:
            common.update(quant=quant, yuv=yuv, rgb=rgb.reshape(-1), pack=pack, index=index)
        np.savez_compressed(os.path.join(BIG, name + ".npz"), **common)
        print(f"{name}: {len(jpg)} B jpeg, {g.coded_blocks} blocks, {os.path.getsize(os.path.join(BIG, name + '.npz'))} B npz")


if __name__ == "__main__":
    main()
