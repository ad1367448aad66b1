"""Device-resident throughput of the fused kernel per (size, sampling) kind: which members of a mixed batch are slow."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import jpeg_gpu_b200 as J
from jpeg_gpu_b200 import synth

dev = torch.device("cuda", 0)
ctx = J.Context(0)
q = synth.quality_tables(85)
d_q = torch.from_numpy(q.astype(np.int16).reshape(-1)).to(dev)
peak = 6447.8
import os
SIZES = [tuple(int(v) for v in s.split("x")) for s in os.environ.get("KIND_SIZES", "512x512,1920x1080,3840x2160,70x50,1000x563,1537x771,1008x563,1536x771").split(",")]
MODES = os.environ.get("KIND_MODES", "gray,444,420,422").split(",")
for (w, h) in SIZES:
    for ss in MODES:
        n = max(4, int(600e6 / (w * h)))       # ~600 Mpx per batch
        n = min(n, 4096)
        hs, vs = J.SUBSAMPLINGS[ss]
        descs = [J.ImageDesc(w, h, hs, vs, tq=(0, 1, 1)[:len(hs)]) for _ in range(n)]
        coef_len, rgb_len, _ = J.pack_batch(descs)
        d_coef = torch.zeros(coef_len, dtype=torch.int16, device=dev)
        one = synth.torch_batch_coefficients(descs[:1], descs[0].query_layout().coef_len + 64, q, dev)
        L = descs[0].query_layout().coef_len
        for d in descs:
            d_coef[d.coef_off:d.coef_off + L] = one[:L]
        d_rgb = torch.zeros(rgb_len, dtype=torch.uint8, device=dev)
        plan = ctx.plan(descs, rgb=True)
        for _ in range(3):
            plan.run(d_coef, d_q, d_rgb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            plan.run(d_coef, d_q, d_rgb)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{w}x{h} {ss:4s} x{n:5d}: {ms:7.3f} ms  {n*w*h/1e6/ms*1e3/1e6:6.3f} Tpx/s  frac {plan.bytes/ms/1e6/peak:5.3f}")
        plan.close()
        del d_coef, d_rgb
