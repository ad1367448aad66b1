"""Damaged files through both entropy decoders: the device path (with its verification and its
fallback) must give the host path's status and pixels for every file.
python tools/fuzz_jpegs_gpu.py [seed] [files]"""
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpeg_gpu_b200 as J  # noqa: E402


def main():
    from PIL import Image
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    rng = np.random.default_rng(seed)
    gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    base = [open(os.path.join(gold, n), "rb").read() for n in sorted(os.listdir(gold)) if n.endswith(".jpg")]
    for (w, h, ss, rst) in [(640, 360, 2, 0), (333, 222, 1, 5), (320, 240, 2, 20), (200, 120, 0, 3)]:
        pic = rng.integers(0, 255, size=(h, w, 3)).astype(np.uint8) // 2 + 60
        bio = io.BytesIO()
        Image.fromarray(pic).save(bio, "JPEG", quality=80, subsampling=ss, restart_marker_blocks=rst)
        base.append(bio.getvalue())
    files = []
    for i in range(count):
        b = bytearray(base[i % len(base)])
        sos = bytes(b).index(b"\xff\xda")
        mode = i % 4
        if mode == 0:
            for pos in rng.integers(sos + 14, len(b) - 2, size=int(rng.integers(1, 4))):
                b[pos] = int(rng.integers(0, 256))
        elif mode == 1:
            b = b[:int(rng.integers(sos + 20, len(b)))]
        elif mode == 2:
            pos = int(rng.integers(sos + 14, len(b) - 4))
            del b[pos:pos + int(rng.integers(1, 4))]
        files.append(bytes(b))     # mode 3: undamaged
    os.dup2(os.open(os.devnull, os.O_WRONLY), 2)   # the readers' one-line error messages
    ctx = J.Context(0)
    got, gi = ctx.decode_jpegs(files, strict=False, entropy="gpu")
    ref, ri = ctx.decode_jpegs(files, strict=False, entropy="cpu")
    bad = on_device = rejected = 0
    for k, (a, b) in enumerate(zip(gi, ri)):
        rejected += a.status != 0
        on_device += a.status == 0 and a.tasks > 1
        if a.status != b.status:
            bad += 1
            print("status differs", k, k % 4, a.status, b.status)
        elif a.status == 0 and not np.array_equal(got[a.rgb_off:a.rgb_off + a.rgb_len], ref[b.rgb_off:b.rgb_off + b.rgb_len]):
            bad += 1
            print("pixels differ", k, k % 4)
    print(f"{count} files: {on_device} decoded on the device, {rejected} rejected by both, {bad} disagreements")
    ctx.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
