"""Debugging aid: per-warp start/end times of the fused kernel (library built with -DJGPU_MCU_TRACE,
selected with JGPU_LIB_PATH).  Prints how evenly the warps of the last launch finished."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import jpeg_gpu_b200 as J
from jpeg_gpu_b200 import synth, _capi
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "4k420_b16"
descs, coef_len, rgb_len, mine = bench.build_batch(wl)
dev = torch.device("cuda", 0)
q = synth.quality_tables(85)
d_q = torch.from_numpy(q.astype(np.int16).reshape(-1)).to(dev)
d_coef = synth.torch_batch_coefficients(descs, coef_len, q, dev)
d_rgb = torch.zeros(rgb_len, dtype=torch.uint8, device=dev)
ctx = J.Context(0)
plan = ctx.plan(descs, rgb=True)
for _ in range(4):
    plan.run(d_coef, d_q, d_rgb)
torch.cuda.synchronize()
n = 148 * 12
buf = np.zeros(4 * n, dtype=np.uint64)
L = _capi.lib()
L.jgpu_mcu_trace_read.argtypes = [C.c_void_p, C.c_int]
assert L.jgpu_mcu_trace_read(buf.ctypes.data_as(C.c_void_p), 4 * n) == 0
t = buf.reshape(n, 4).astype(np.int64)
smid, t0, t1, steps = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
ok = steps > 0
base = t0[ok].min()
print(f"{wl}: {ok.sum()} warps, kernel span {(t1[ok].max() - base) / 1e3:.1f} us; start spread {(t0[ok].max() - base) / 1e3:.1f} us")
dur = (t1 - t0)[ok] / 1e3
print(f"per-warp duration us: min {dur.min():.1f} p10 {np.percentile(dur, 10):.1f} median {np.median(dur):.1f} p90 {np.percentile(dur, 90):.1f} max {dur.max():.1f}; steps {np.unique(steps[ok])}")
end = (t1 - base)[ok] / 1e3
print(f"per-warp end time us: min {end.min():.1f} p10 {np.percentile(end, 10):.1f} median {np.median(end):.1f} p90 {np.percentile(end, 90):.1f} max {end.max():.1f}")
per_sm = {}
for s_, e_, st_ in zip(smid[ok], end, steps[ok]):
    per_sm.setdefault(int(s_), []).append(e_)
ends = np.array([max(v) for v in per_sm.values()])
cnt = np.array([len(v) for v in per_sm.values()])
print(f"{len(per_sm)} SMs; warps per SM {np.unique(cnt)}; SM end time us: min {ends.min():.1f} median {np.median(ends):.1f} max {ends.max():.1f}")
# warp index within CTA vs duration
w = np.arange(n)[ok] % 12
for k in range(12):
    print(f"  warp {k:2d}: median dur {np.median(dur[w == k]):.1f} us", end="")
print()
order = np.argsort(ends)
print("slowest SMs:", [(list(per_sm.keys())[i], round(float(ends[i]), 1)) for i in order[-8:]])
print("fastest SMs:", [(list(per_sm.keys())[i], round(float(ends[i]), 1)) for i in order[:8]])
med = np.median(dur)
idx = np.arange(n)[ok]
slow = dur > 1.25 * med
print(f"slow warps (> 1.25 x median): {slow.sum()} of {ok.sum()}")
from collections import Counter
print(" by gw % 8:", sorted(Counter((idx[slow] % 8).tolist()).items()))
print(" by warp in CTA:", sorted(Counter((idx[slow] % 12).tolist()).items()))
print(" by CTA:", sorted(Counter((idx[slow] // 12).tolist()).items())[:40])
print(" all warps by gw % 8: median dur", [round(float(np.median(dur[idx % 8 == k])), 1) for k in range(8)])
print(" slow warps' durations:", np.round(np.sort(dur[slow])[::max(1, slow.sum() // 20)], 0).tolist())
