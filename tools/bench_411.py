import sys
import numpy as np, torch
sys.path.insert(0, ".")
import jpeg_gpu_b200 as J
from jpeg_gpu_b200 import synth
dev = torch.device("cuda", 0); ctx = J.Context(0)
q = synth.quality_tables(85); d_q = torch.from_numpy(q.astype(np.int16).reshape(-1)).to(dev)
for ss, force in [("411", False), ("411", True), ("422", False)]:
    w, h, n = 3840, 2160, 128
    hs, vs = J.SUBSAMPLINGS[ss]
    descs = [J.ImageDesc(w, h, hs, vs, tq=(0, 1, 1)) for _ in range(n)]
    coef_len, rgb_len, _ = J.pack_batch(descs)
    L = descs[0].query_layout().coef_len
    one = synth.torch_batch_coefficients(descs[:1], L + 64, q, dev)
    d_coef = torch.zeros(coef_len, dtype=torch.int16, device=dev)
    for d in descs: d_coef[d.coef_off:d.coef_off + L] = one[:L]
    d_rgb = torch.zeros(rgb_len, dtype=torch.uint8, device=dev)
    plan = ctx.plan(descs, rgb=True, force_generic=force)
    for _ in range(3): plan.run(d_coef, d_q, d_rgb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): plan.run(d_coef, d_q, d_rgb)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"4K {ss} x{n} {'generic' if force else 'fused'}: {ms:.3f} ms, {n*w*h/ms/1e9:.3f} Tpx/s, frac {plan.bytes/ms/1e6/6447.8:.3f}")
    plan.close(); del d_coef, d_rgb
