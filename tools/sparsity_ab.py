"""A/B of the zero-row shortcut (JGPU_SKIP_ZERO_ROWS) on two coefficient sets: the synthetic distribution of SURVEY 8d and
the coefficients of a photograph-like picture (1/f noise, flat regions, edges) encoded by Pillow at q85 4:2:0.
Run once per library variant: JGPU_LIB_PATH=... python tools/sparsity_ab.py"""
import io
import os
import sys

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, ".")
import jpeg_gpu_b200 as J
from jpeg_gpu_b200 import synth
from jpeg_gpu_b200.decoder import Decoder


def photo_like(w, h, seed=7, slope=1.6):
    rng = np.random.default_rng(seed)
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.fftfreq(w)[None, :]
    f = np.sqrt(fx * fx + fy * fy)
    f[0, 0] = 1.0
    chans = []
    for c in range(3):
        spec = (rng.normal(size=(h, w)) + 1j * rng.normal(size=(h, w))) / f ** slope
        img = np.real(np.fft.ifft2(spec))
        img = (img - img.mean()) / img.std()
        chans.append(img)
    pic = np.stack(chans, -1) * 48 + 128
    pic[: h // 3] = pic[: h // 3] * 0.15 + 170            # a flat "sky"
    pic[:, w // 2: w // 2 + 6] = 20                       # edges
    pic[h // 2: h // 2 + 4] = 235
    return np.clip(pic, 0, 255).astype(np.uint8)


def main():
    dev = torch.device("cuda", 0)
    ctx = J.Context(0)
    w, h, n = 3840, 2160, 64
    hs, vs = J.SUBSAMPLINGS["420"]
    descs = [J.ImageDesc(w, h, hs, vs, tq=(0, 1, 1)) for _ in range(n)]
    coef_len, rgb_len, _ = J.pack_batch(descs)
    L = descs[0].query_layout().coef_len
    q85 = synth.quality_tables(85)
    sets = {}
    sets["synthetic (SURVEY 8d)"] = (synth.image_coefficients(descs[0], q85, 1), q85)
    for label, slope in (("smooth photograph-like picture (1/f^1.6), Pillow q85", 1.6), ("detailed photograph-like picture (1/f^1.0), Pillow q85", 1.0)):
        buf = io.BytesIO()
        Image.fromarray(photo_like(w, h, slope=slope)).save(buf, "JPEG", quality=85, subsampling=2)
        with Decoder(buf.getvalue(), impl="jfront") as d:
            hdr = d.decode_header()
            sets[f"{label}, {len(buf.getvalue()) / 1e6:.2f} MB"] = (d.decode_image("quant")["coef"][:L].astype(np.int16), hdr.qtabs.astype(np.uint16))
    for name, (c, q) in sets.items():
        lay = descs[0].query_layout()
        # fraction of (64-block group, coefficient row) that is all zero, per coefficient row
        p0 = lay.planes[0]
        blocks = c[p0.coef_off:p0.coef_off + p0.hblocks * p0.vblocks * 64].reshape(-1, 8, 8)
        g = blocks[: len(blocks) // 64 * 64].reshape(-1, 64, 8, 8)
        zero = (g == 0).all(axis=(1, 3)).mean(axis=0)
        d_coef = torch.zeros(coef_len, dtype=torch.int16, device=dev)
        one = torch.from_numpy(c).to(dev)
        for dd in descs:
            d_coef[dd.coef_off:dd.coef_off + L] = one
        d_q = torch.from_numpy(q.astype(np.int16).reshape(-1)).to(dev)
        d_rgb = torch.zeros(rgb_len, dtype=torch.uint8, device=dev)
        plan = ctx.plan(descs, rgb=True)
        for _ in range(3):
            plan.run(d_coef, d_q, d_rgb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            plan.run(d_coef, d_q, d_rgb)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        csum = int(d_rgb[:w * h * 3].to(torch.int64).sum().item())
        print(f"{os.path.basename(os.environ.get('JGPU_LIB_PATH', 'default'))}: {name}: {ms:.4f} ms for {n} images, "
              f"pixel checksum {csum}; luma rows 1-7 all zero in a 64-block group: " + " ".join(f"{z:.2f}" for z in zero[1:]))
        plan.close()


if __name__ == "__main__":
    main()
