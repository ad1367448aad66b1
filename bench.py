#!/usr/bin/env python
"""bench.py — coefficient -> RGB block decode throughput on B200.

Metric (BASELINE.json): Mpixels/s decoded (coeff -> RGB); HBM GB/s vs roofline.
Workload at every N: BASELINE config 3 — 3840x2160 4:2:0, batch 256 synthetic
coefficient sets PER GPU (weak scaling), inputs resident in HBM for `value`,
host buffers with H2D/D2H inside the timed region for `e2e`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, subsampling, images per GPU)
    "4k420_b256": (3840, 2160, "420", 256),
    "4k422_b128": (3840, 2160, "422", 128),
    "1080p420_b1": (1920, 1080, "420", 1),
    "4k420_b16": (3840, 2160, "420", 16),
    # the other sampling modes of the fused kernel (tuning runs; not BASELINE configurations)
    "4k444_b64": (3840, 2160, "444", 64),
    "4kgray_b256": (3840, 2160, "gray", 256),
    "4k440_b128": (3840, 2160, "440", 128),
    "1080p420_b512": (1920, 1080, "420", 512),
}
METRIC = "Mpixels/s decoded (coeff->RGB)"
UNIT = "Mpixels/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi style clock / throttle-reason samples during the timed region (NVML)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


MIXED_PER_GPU = 96   # BASELINE config 5: images per GPU of the mixed-resolution stress batch


def mixed_shapes(world: int):
    """BASELINE config 5: a seed-fixed shuffle of {512^2, 1080p, 4K} x {gray, 4:4:4,
    4:2:0, 4:2:2} plus odd sizes that exercise cropping and tails; MIXED_PER_GPU x world images."""
    rng = np.random.default_rng(20261017)
    sizes = [(512, 512), (1920, 1080), (3840, 2160), (70, 50), (1000, 563), (1537, 771)]
    kinds = ["gray", "444", "420", "422"]
    out = []
    for _ in range(MIXED_PER_GPU * world):
        w, h = sizes[int(rng.choice(len(sizes), p=[0.25, 0.3, 0.25, 0.05, 0.1, 0.05]))]
        out.append((w, h, kinds[int(rng.integers(len(kinds)))]))
    return out


def build_batch(workload: str, rank: int = 0, world: int = 1):
    """Returns (descs, coef_len, rgb_len, global image indices of this rank)."""
    import jpeg_gpu_b200 as J
    from jpeg_gpu_b200 import shard
    if workload == "mixed_stress":
        shapes = mixed_shapes(world)
        mine = shard.shard_lpt([w * h for (w, h, _) in shapes], world)[rank]
        descs = []
        for i in mine:
            w, h, ss = shapes[i]
            hs, vs = J.SUBSAMPLINGS[ss]
            descs.append(J.ImageDesc(w, h, hs, vs, tq=(0, 1, 1)[:len(hs)]))
    else:
        w, h, ss, n = WORKLOADS[workload]
        hs, vs = J.SUBSAMPLINGS[ss]
        descs = [J.ImageDesc(w, h, hs, vs, tq=(0, 1, 1)[:len(hs)]) for _ in range(n)]
        mine = list(range(rank * n, (rank + 1) * n))
    coef_len, rgb_len, _ = J.pack_batch(descs)
    return descs, coef_len, rgb_len, mine


# -- synthetic inputs for the CPU arm: restated here so that the reference arm maps nothing of the product
#    (jpeg_gpu_b200/synth.py generates the same distribution for the GPU arm)
_ANNEX_K = (
    [16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87,
     80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92,
     95, 98, 112, 100, 103, 99],
    [17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99,
     99, 99] + [99] * 32)
_NATURAL = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
            28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
            54, 47, 55, 62, 63]
_SUBSAMP = {"gray": ((1,), (1,)), "444": ((1, 1, 1), (1, 1, 1)), "422": ((2, 1, 1), (1, 1, 1)),
            "420": ((2, 1, 1), (2, 1, 1)), "440": ((1, 1, 1), (2, 1, 1))}


def _cpu_tables(quality=85):
    scale = 5000 // quality if quality < 50 else 200 - 2 * quality
    out = np.zeros((4, 64), dtype=np.uint16)
    for slot in range(4):
        out[slot] = np.clip((np.array(_ANNEX_K[slot & 1], dtype=np.int64) * scale + 50) // 100, 1, 255)
    return out


def _cpu_plane(rng, n, q):
    q = q.astype(np.int64)
    lim = 2047 // q
    blk = np.zeros((n, 64), dtype=np.int64)
    blk[:, 0] = np.rint(rng.normal(0.0, 300.0 / q[0], size=n))
    for k in range(1, 64):
        nat = _NATURAL[k]
        on = rng.random(n) < 0.9 * np.exp(-k / 8.0)
        mag = 1 + np.floor(rng.exponential(24.0 * np.exp(-k / 12.0) / q[nat], size=n))
        blk[:, nat] = np.where(on, mag * np.where(rng.random(n) < 0.5, -1, 1), 0)
    return np.clip(blk, -lim, lim).astype(np.int16)


def workload_string(workload: str) -> str:
    """`config.workload` of both arms (the driver compares them)."""
    if workload == "mixed_stress":
        return (f"mixed-resolution stress (BASELINE config 5): {MIXED_PER_GPU} images per GPU of "
                "{512^2,1080p,4K,70x50,1000x563,1537x771} x {gray,444,420,422}, LPT-sharded by pixels")
    w, h, ss, n = WORKLOADS[workload]
    return f"{w}x{h} {ss}, batch {n} per GPU, synthetic coefficient planes (SURVEY 8d)"


def cpu_reference_run(workload: str, n_images: int, repeats: int, threads: int):
    """Times the reference's CPU implementation of the path (oracle/_ref when it was compiled, else
    our C port) on `n_images` images of the workload.  Imports only oracle/ (never the product)."""
    import oracle
    lib = oracle.best()
    w, h, ss, n = WORKLOADS[workload]
    hs, vs = _SUBSAMP[ss]
    k = max(1, min(n_images, n))
    g = oracle.geometry(w, h, hs, vs)
    tq = (0, 1, 1)[:len(hs)]
    q = _cpu_tables(85)
    coef_stride = (g.coef_len + 63) // 64 * 64
    rgb_stride = (g.rgb_len + 255) // 256 * 256
    # distinct coefficients for 2 images, tiled: the CPU cost does not depend on the values' identity
    coef = np.zeros(coef_stride * k, dtype=np.int16)
    for i in range(min(k, 2)):
        rng = np.random.default_rng(20261017 + i)
        for p, t in zip(g.planes, tq):
            nb = p.hblocks * p.vblocks
            coef[i * coef_stride + p.coef_off:i * coef_stride + p.coef_off + nb * 64] = _cpu_plane(rng, nb, q[t]).reshape(-1)
    for i in range(2, k):
        coef[i * coef_stride:i * coef_stride + g.coef_len] = coef[(i & 1) * coef_stride:(i & 1) * coef_stride + g.coef_len]
    rows = np.stack([oracle.make_desc(g, tq, i * coef_stride, i * rgb_stride, 0) for i in range(k)])
    qq = q.reshape(1, 4, 64)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        lib.decode_batch(rows, coef, qq, rgb_stride * k, 0, threads)
        times.append(time.perf_counter() - t0)
    return lib.kind, k * w * h / 1e6, times


class JsonStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its
    version banner on stdout at init), so everything else is sent to stderr: fd 1 is pointed at
    fd 2 for the life of the process and the line goes out through the saved descriptor."""

    def __init__(self):
        sys.stdout.flush()
        self._fd = os.dup(1)
        os.dup2(2, 1)

    def emit(self, obj) -> None:
        sys.stdout.flush()
        os.write(self._fd, (json.dumps(obj) + "\n").encode())


def host_memory_available() -> int:
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def measured_traffic(workload: str):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this
    workload (profiles/r2_traffic.json, else the round-1 file), or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                t = json.load(f).get(workload)
            if t:
                return t["dram_bytes_read"] + t["dram_bytes_write"]
    return None


def device_leg(ctx, workload, rank, world, steps, warmup, dev, q, d_q, force_generic=False, keep=False,
               parity_images=4, yuv=False):
    """One workload, inputs resident in HBM: W warm-ups, K timed launches bracketed by CUDA events, a
    barrier + synchronize on both sides, max over ranks; then images of the TIMED output buffer are
    checked bit for bit against the CPU oracle.  Returns the JSON fragment (+ the buffers with keep)."""
    import torch
    import jpeg_gpu_b200 as J
    import oracle
    from jpeg_gpu_b200 import shard, synth
    threads = os.cpu_count() or 1
    descs, coef_len, rgb_len, mine = build_batch(workload, rank, world)
    yuv_len = 0
    if yuv:
        coef_len, rgb_len, yuv_len = J.pack_batch(descs, want_yuv=True)
    d_coef = synth.torch_batch_coefficients(descs, coef_len, q, dev, first_index=mine[0] if mine else 0)
    d_out = torch.zeros(yuv_len if yuv else rgb_len, dtype=torch.uint8, device=dev)
    plan = ctx.plan(descs, rgb=not yuv, yuv=yuv, force_generic=force_generic)
    run = (lambda: plan.run(d_coef, d_q, None, d_out)) if yuv else (lambda: plan.run(d_coef, d_q, d_out))
    px_rank = sum(d.width * d.height for d in descs)
    px_total = shard.sum_over_ranks(float(px_rank), dev)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warmup):
        run()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clk:
        ev0.record()
        for _ in range(steps):
            run()
        ev1.record()
        barrier()
    ms = shard.max_over_ranks(ev0.elapsed_time(ev1), dev) / steps
    value = px_total / 1e6 / (ms / 1e3)

    # parity of the timed buffer: first, last and seeded-random images of this rank vs the oracle
    lib = oracle.best()
    rng = np.random.default_rng(1234 + rank)
    n = len(descs)
    if workload == "mixed_stress":
        picks = sorted(set(range(0, n, 8)))
    else:
        picks = {0, n - 1}
        while len(picks) < min(parity_images, n):
            picks.add(int(rng.integers(0, n)))
        picks = sorted(picks)
    bad = 0
    for k in picks:
        d = descs[k]
        lay = d.query_layout()
        g = oracle.geometry(d.width, d.height, d.hsamp, d.vsamp)
        c = d_coef[d.coef_off:d.coef_off + lay.coef_len].cpu().numpy()
        want_rgb, want_planes = lib.decode_image(g, c, q, d.tq, nthreads=min(threads, 16))
        if yuv:
            got = d_out[d.yuv_off:d.yuv_off + lay.data_len].cpu().numpy()
            bad += int(not np.array_equal(got, np.concatenate([p.ravel() for p in want_planes])))
        else:
            got = d_out[d.rgb_off:d.rgb_off + lay.rgb_len].cpu().numpy()
            bad += int(not np.array_equal(got, want_rgb.reshape(-1)))
    parity = {"checked": int(shard.sum_over_ranks(float(len(picks)), dev)),
              "mismatching_images": int(shard.sum_over_ranks(float(bad), dev)), "oracle": lib.kind,
              "what": "images of the timed output buffer, bit for bit"}

    peak, peak_src = peaks()
    achieved = plan.bytes / (ms * 1e-3) / 1e9
    coef_bytes = sum(128 * d.query_layout().coded_blocks for d in descs)
    out = {
        "workload": workload_string(workload), "value": value, "unit": UNIT, "ms_per_step": ms, "steps": steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(workload + ("_yuv" if yuv else "")), "peak_source": peak_src,
                     "read_only_frac": coef_bytes / (ms * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes_per_launch": plan.bytes},
        "parity": parity, "kernel_launches_per_step": plan.launches, "clocks": clk.summary(),
        "l2": "inputs larger than L2 (%.2f GB coef + %.2f GB out per GPU)" % (coef_len * 2 / 1e9, d_out.numel() / 1e9),
    }
    if keep:
        return out, (descs, d_coef, d_out, plan, barrier)
    plan.close()
    del d_coef, d_out
    torch.cuda.empty_cache()
    return out, None


def main():
    out = JsonStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k420_b256", choices=sorted(WORKLOADS) + ["mixed_stress"])
    ap.add_argument("--e2e-steps", type=int, default=None, help="steps of the host-buffer legs (default: min(steps, 5))")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config-4 / config-5 / planes sub-legs")
    ap.add_argument("--force-generic", action="store_true")
    ap.add_argument("--yuv", action="store_true", help="planes (xjpeg's YUV output) instead of pixels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    mixed = args.workload == "mixed_stress"
    headline = args.workload == "4k420_b256" and not args.yuv and not args.force_generic
    if mixed:
        w = h = 0
        ss, n_img = "mixed", MIXED_PER_GPU
        args.no_e2e = True    # the host-buffer legs are measured on the uniform workloads
        if args.impl == "reference":
            raise SystemExit("--impl reference runs the uniform workloads")
    else:
        w, h, ss, n_img = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        # the whole batch of the workload per step (2 distinct coefficient sets, tiled): about a
        # second of CPU work per step on 16 threads
        kind, mpx, times = cpu_reference_run(args.workload, n_img, args.warmup + args.steps, threads)
        timed = times[args.warmup:]
        total = sum(timed)
        value = mpx * len(timed) / total
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_string(args.workload),
                       "path": "CPU path of the reference (xjpeg dequant + glj_real_idct8x8 + clamp, yuv.fs.glsl colour), "
                               f"{threads} host threads, every step = the batch of ONE GPU ({n_img} images)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"{n_img} images per step x {len(timed)} steps, {threads} threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        out.emit(line)
        return 0

    # ---------------------------------------------------------------- our arm
    import torch
    import jpeg_gpu_b200 as J
    from jpeg_gpu_b200 import shard, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device; there is no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa = None
    if world > 1:
        # one process per GPU: keep its host buffers and copy threads next to its GPU's PCIe root
        numa = shard.bind_to_gpu_numa_node(local_rank)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # quantisation tables: rank 0's copy is THE copy (one broadcast)
    q = shard.broadcast_tables(synth.quality_tables(85), device=dev)
    d_q = torch.from_numpy(q.astype(np.int16).reshape(-1)).to(dev)
    ctx = J.Context(local_rank)
    main_leg, kept = device_leg(ctx, args.workload, rank, world, args.steps, args.warmup, dev, q, d_q,
                                force_generic=args.force_generic, keep=True, yuv=args.yuv)
    descs, d_coef, d_rgb, plan, barrier = kept

    # end-to-end: host buffers through the C ABI, copies inside the timed region.
    # `e2e`      the dense QUANT planes cross the link (north_star's input format);
    # `e2e_pack` the reference's PACK stream crosses instead (JPEG_DECODE_PACK, what its GL path
    #            uploads with `-o pack`) and is expanded on the device.
    # Every rank stages its whole batch in pinned host memory (as the one-GPU run does) when the host
    # has the memory for it, else a quarter of it four times per step; `staging` says which.
    e2e = e2e_pack = None
    if not args.no_e2e and not args.yuv:
        from concurrent.futures import ThreadPoolExecutor
        from jpeg_gpu_b200.batch import pack_from_quant
        e2e_steps = args.e2e_steps or min(args.steps, 5)
        full_bytes = d_coef.numel() * 2 + d_rgb.numel() + d_coef.numel() // 3
        whole = world == 1 or host_memory_available() > 2 * full_bytes * world
        sub_n = n_img if whole else max(1, n_img // 4)
        calls = -(-n_img // sub_n)
        sub = descs[:sub_n]
        sub_coef_len = sub[-1].coef_off + sub[-1].query_layout().coef_len
        sub_rgb_len = sub[-1].rgb_off + sub[-1].query_layout().rgb_len
        sub_px = sub_n * w * h
        sub_coef_bytes = sum(128 * d.query_layout().coded_blocks for d in sub)
        staging = "whole batch pinned" if whole else f"{sub_n} images pinned, {calls} calls per step"
        h_coef = torch.empty(sub_coef_len, dtype=torch.int16).pin_memory()
        h_coef.copy_(d_coef[:sub_coef_len].cpu())
        h_rgb = torch.zeros(sub_rgb_len, dtype=torch.uint8).pin_memory()
        ctx.decode_batch_host(sub, h_coef, q, h_rgb, None, force_generic=args.force_generic)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps * calls):
            ctx.decode_batch_host(sub, h_coef, q, h_rgb, None, force_generic=args.force_generic)
        barrier()
        dt = shard.max_over_ranks(time.perf_counter() - t0, dev)
        assert torch.equal(h_rgb[:1 << 20], d_rgb[:1 << 20].cpu()), "host path and device path disagree"
        e2e = {"value": world * sub_px * calls * e2e_steps / 1e6 / dt, "unit": UNIT,
               "h2d_bytes_per_step": int((sub_coef_bytes + q.nbytes) * calls), "d2h_bytes_per_step": int(sub_rgb_len * calls),
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps, "calls_per_step": calls, "staging": staging,
               "input": "dense QUANT planes (int16), pinned host memory"}

        # PACK leg: pack the same planes on the host (outside the timed region: in the reference
        # the Huffman reader writes this stream directly), then time words+index in, RGB out
        coef_np = h_coef.numpy()
        with ThreadPoolExecutor(max_workers=min(threads, 32)) as ex:
            parts = list(ex.map(lambda d: pack_from_quant(d, coef_np[d.coef_off:d.coef_off + d.query_layout().coef_len]), sub))
        pack_off = np.zeros(sub_n + 1, dtype=np.int64)
        pack_off[1:] = np.cumsum([p.size for p, _ in parts])
        h_pack = torch.empty(int(pack_off[-1]), dtype=torch.int16).pin_memory()
        h_index = torch.zeros(sub_coef_len // 64, dtype=torch.int32).pin_memory()
        for d, (p, ix), o in zip(sub, parts, pack_off[:-1]):
            h_pack[int(o):int(o) + p.size] = torch.from_numpy(p.view(np.int16))
            h_index[d.coef_off // 64:d.coef_off // 64 + ix.size] = torch.from_numpy(ix)
        del parts
        h_rgb.zero_()
        ctx.decode_batch_host_packed(sub, h_pack, pack_off, h_index, q, h_rgb, None, force_generic=args.force_generic)
        assert torch.equal(h_rgb[:1 << 20], d_rgb[:1 << 20].cpu()), "packed host path and device path disagree"
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps * calls):
            ctx.decode_batch_host_packed(sub, h_pack, pack_off, h_index, q, h_rgb, None, force_generic=args.force_generic)
        barrier()
        dt = shard.max_over_ranks(time.perf_counter() - t0, dev)
        e2e_pack = {"value": world * sub_px * calls * e2e_steps / 1e6 / dt, "unit": UNIT,
                    "h2d_bytes_per_step": int((h_pack.numel() * 2 + h_index.numel() * 4 + pack_off.nbytes + q.nbytes) * calls),
                    "d2h_bytes_per_step": int(sub_rgb_len * calls), "steps": e2e_steps,
                    "ms_per_step": 1e3 * dt / e2e_steps, "calls_per_step": calls, "staging": staging,
                    "input": "PACK run/level words (uint16) + per-block index (int32), pinned host memory; "
                             "expanded on the device by k_unpack",
                    "bytes_per_block": (h_pack.numel() * 2 + h_index.numel() * 4) / (sub_coef_bytes / 128)}
        del h_coef, h_rgb, h_pack, h_index
    plan.close()
    del d_coef, d_rgb
    torch.cuda.empty_cache()

    # JPEG files in, RGB out (jgpu_decode_jpegs), on every rank: Huffman decoding on the GPU, so what
    # crosses the link is the file (0.49 B/px) and the pixels (3 B/px) or the planes (1.5 B/px).  Not
    # BASELINE's metric (that starts at the coefficient planes) -- reported beside it because this is
    # what a user of the reference's viewer loop actually waits for.
    e2e_jpeg = None
    if not args.no_e2e and not args.yuv and args.workload.startswith("4k"):
        try:
            import io
            from PIL import Image
            rng = np.random.default_rng(synth.SEED_BASE)
            yy, xx = np.mgrid[0:h, 0:w]
            base = np.stack([(xx * 5 + yy * 3) % 256, (yy * 7 + xx) % 256, (xx * 2 + yy * 9) % 256], -1)
            pic = np.clip(base + rng.integers(-24, 25, size=base.shape), 0, 255).astype(np.uint8)
            bio = io.BytesIO()
            Image.fromarray(pic[..., 0] if ss == "gray" else pic).save(
                bio, "JPEG", quality=85, subsampling={"420": 2, "422": 1, "444": 0, "gray": 0, "440": 0}[ss],
                restart_marker_blocks=w // 16)
            files = [bio.getvalue()] * 32
            host_threads = max(1, threads // world)
            total, _ = J.probe_jpegs(files)
            jpeg_rgb = torch.zeros(total, dtype=torch.uint8).pin_memory()
            e2e_jpeg = {"unit": UNIT, "files_per_step": len(files) * world, "steps": 3, "host_threads_per_rank": host_threads,
                        "jpeg_bytes": len(files[0]),
                        "input": f"{w}x{h} {ss} baseline JPEG (Pillow, q85, one restart interval per MCU row), "
                                 f"32 files per rank, host buffers in and out",
                        "h2d_bytes_per_step": len(files) * len(files[0]) * world, "d2h_bytes_per_step": int(total) * world}

            def timed(fn, reps=3):
                fn()   # warm-up (plan, staging)
                barrier()
                t0 = time.perf_counter()
                for _ in range(reps):
                    fn()
                barrier()
                dt = shard.max_over_ranks(time.perf_counter() - t0, dev)
                return reps * len(files) * world * w * h / 1e6 / dt

            e2e_jpeg["value"] = timed(lambda: ctx.decode_jpegs(files, jpeg_rgb, nthreads=host_threads, entropy="gpu"))
            if world == 1:
                e2e_jpeg["cpu_entropy_value"] = timed(lambda: ctx.decode_jpegs(files, jpeg_rgb, nthreads=host_threads, entropy="cpu"))
            # the same with the pixels left in device memory (callers whose next stage runs on the GPU)
            # (half the host threads: this call is as long as the chain of kernels, and the thread that feeds
            # it must not be starved by the unstuffing workers -- the library's own default for this output)
            jpeg_dev = torch.empty(total, dtype=torch.uint8, device=dev)
            dev_threads = max(1, host_threads // 2)
            e2e_jpeg["device_out_value"] = timed(lambda: ctx.decode_jpegs(files, jpeg_dev, nthreads=dev_threads, entropy="gpu"))
            e2e_jpeg["device_out_host_threads_per_rank"] = dev_threads
            del jpeg_dev
            if world == 1:
                # the same on 128 files (3.2 GB of pixels): the pipeline's steady state, not its ramp
                files128 = files * 4
                total128, _ = J.probe_jpegs(files128)
                jpeg_dev = torch.empty(total128, dtype=torch.uint8, device=dev)
                v = timed(lambda: ctx.decode_jpegs(files128, jpeg_dev, nthreads=dev_threads, entropy="gpu"))
                e2e_jpeg["device_out_128_files_value"] = v * 4   # timed() counts len(files) files
                del jpeg_dev, files128
            # planes instead of pixels (what the reference's xjpeg backend itself produces)
            yuv_total, _ = J.probe_jpegs(files, out="yuv")
            jpeg_yuv = torch.zeros(yuv_total, dtype=torch.uint8).pin_memory()
            e2e_jpeg["yuv_out_value"] = timed(lambda: ctx.decode_jpegs(files, jpeg_yuv, nthreads=host_threads, entropy="gpu", out="yuv"))
            e2e_jpeg["yuv_d2h_bytes_per_step"] = int(yuv_total) * world
            del jpeg_yuv, jpeg_rgb
            # the reference's own CPU path for the same files: xjpeg_decode_image(YUV), i.e. Huffman +
            # dequant + IDCT into planes (it has no colour conversion), one file per thread on all host
            # cores (oracle/_ref; checker code, timed here as the baseline)
            if rank == 0 and world == 1:
                try:
                    import oracle
                    if oracle.have_reference():
                        ref = oracle.reference()
                        k = min(len(files), threads)
                        ref.ref_decode(files[0], "yuv")
                        t0 = time.perf_counter()
                        with ThreadPoolExecutor(threads) as ex:
                            list(ex.map(lambda f: ref.ref_decode(f, "yuv"), files[:k]))
                        dt = time.perf_counter() - t0
                        e2e_jpeg["cpu_reference_value"] = k * w * h / 1e6 / dt
                        e2e_jpeg["cpu_reference"] = (f"xjpeg_decode_image(YUV) of the compiled reference, {k} files on "
                                                     f"{threads} threads, planes only (it has no RGB output)")
                except Exception as exc:   # the checker is optional for this leg
                    e2e_jpeg["cpu_reference"] = f"unavailable: {exc}"
            e2e_jpeg["entropy"] = "Huffman decoding on the GPU (jgpu_huff.cu), host threads only unstuff"
        except ImportError:
            e2e_jpeg = None

    # BASELINE configs 4 and 5 and the planes output of the headline workload, outside the headline's
    # timed region, at every N
    extra = None
    if headline and not args.no_extra:
        extra = {}
        sub_steps = max(5, args.steps // 2)
        extra["config4"], _ = device_leg(ctx, "4k422_b128", rank, world, sub_steps, args.warmup, dev, q, d_q)
        extra["config5"], _ = device_leg(ctx, "mixed_stress", rank, world, sub_steps, args.warmup, dev, q, d_q)
        extra["planes_out"], _ = device_leg(ctx, "4k420_b256", rank, world, sub_steps, args.warmup, dev, q, d_q, yuv=True)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not mixed:
        kind, mpx, times = cpu_reference_run(args.workload, 16, 4, threads)
        best = min(times[1:])
        cpu = {"value": mpx / best, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"16 of the {n_img} images, best of 3 after 1 warm-up, {threads} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": main_leg["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_leg["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload),
                       "l2": main_leg["l2"],
                       "path": ("generic (2 kernels)" if args.force_generic else "fused kernel") + (", planes out" if args.yuv else ""),
                       "kernel_launches_per_step": main_leg["kernel_launches_per_step"], "parallelism": f"images sharded x{world}",
                       "host_binding": numa},
            "roofline": main_leg["roofline"], "cpu_baseline": cpu, "e2e": e2e, "e2e_pack": e2e_pack, "e2e_jpeg": e2e_jpeg,
            "gpu_launches": args.steps * main_leg["kernel_launches_per_step"], "clocks": main_leg["clocks"],
            "parity": main_leg["parity"], "extra": extra,
        }
        out.emit(line)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
