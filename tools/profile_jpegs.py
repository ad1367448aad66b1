"""Times jgpu_decode_jpegs_ex on a batch of generated 4K files (JGPU_TRACE=1 prints the host-side
phases); run under ncu to list the kernels.  python tools/profile_jpegs.py [n_files] [entropy] [rst]"""
import io
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpeg_gpu_b200 as J  # noqa: E402


def main():
    from PIL import Image
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    entropy = sys.argv[2] if len(sys.argv) > 2 else "gpu"
    rst = int(sys.argv[3]) if len(sys.argv) > 3 else 240
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    w, h = 3840, 2160
    rng = np.random.default_rng(20261017)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 5 + yy * 3) % 256, (yy * 7 + xx) % 256, (xx * 2 + yy * 9) % 256], -1)
    # PROFILE_NOISE / PROFILE_Q / PROFILE_SS (0: 4:4:4, 1: 4:2:2, 2: 4:2:0, L: grey) / PROFILE_SMOOTH=1 vary the picture
    noise = int(os.environ.get("PROFILE_NOISE", "24"))
    q = int(os.environ.get("PROFILE_Q", "85"))
    ss = os.environ.get("PROFILE_SS", "2")
    if os.environ.get("PROFILE_SMOOTH") == "1":
        base = np.stack([(xx + yy) // 24 % 256, (yy * 2 + xx) // 32 % 256, (xx * 3) // 40 % 256], -1)
    pic = np.clip(base + rng.integers(-noise, noise + 1, size=base.shape), 0, 255).astype(np.uint8)
    bio = io.BytesIO()
    if ss == "L":
        Image.fromarray(pic[..., 0]).save(bio, "JPEG", quality=q, restart_marker_blocks=rst)
    else:
        Image.fromarray(pic).save(bio, "JPEG", quality=q, subsampling=int(ss), restart_marker_blocks=rst)
    files = [bio.getvalue()] * n
    total, _ = J.probe_jpegs(files)
    device_out = os.environ.get("PROFILE_DEVICE_OUT") == "1"
    out = torch.zeros(total, dtype=torch.uint8, device="cuda:0") if device_out else torch.zeros(total, dtype=torch.uint8).pin_memory()
    ctx = J.Context(0)
    nthreads = int(os.environ.get("PROFILE_THREADS", "0"))
    ctx.decode_jpegs(files, out, entropy=entropy, nthreads=nthreads)
    for _ in range(reps):
        t0 = time.perf_counter()
        ctx.decode_jpegs(files, out, entropy=entropy, nthreads=nthreads)
        dt = time.perf_counter() - t0
        print(f"{entropy}: {n} files of {len(files[0])} bytes in {dt * 1e3:.2f} ms = {n * w * h / 1e6 / dt:.0f} Mpx/s", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
