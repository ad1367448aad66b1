"""Small JPEG batch through the device entropy decoder, for compute-sanitizer runs:
compute-sanitizer --tool memcheck|racecheck python tools/sanitize_jpegs.py"""
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpeg_gpu_b200 as J  # noqa: E402


def main():
    from PIL import Image
    rng = np.random.default_rng(5)
    files = []
    for (w, h, ss, rst) in [(640, 360, 2, 0), (333, 222, 1, 5), (200, 120, 0, 0), (97, 61, "L", 3), (1280, 720, 2, 80)]:
        pic = rng.integers(0, 255, size=(h, w, 3)).astype(np.uint8) // 2 + 60
        bio = io.BytesIO()
        if ss == "L":
            Image.fromarray(pic[..., 0]).save(bio, "JPEG", quality=80, restart_marker_blocks=rst)
        else:
            Image.fromarray(pic).save(bio, "JPEG", quality=80, subsampling=ss, restart_marker_blocks=rst)
        files.append(bio.getvalue())
    gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    files += [open(os.path.join(gold, n), "rb").read() for n in sorted(os.listdir(gold)) if n.endswith(".jpg")]
    ctx = J.Context(0)
    a, ia = ctx.decode_jpegs(files, entropy="gpu")
    b, ib = ctx.decode_jpegs(files, entropy="cpu")
    y, iy = ctx.decode_jpegs(files, entropy="gpu", out="yuv")
    assert all(i.status == 0 for i in ia) and np.array_equal(a, b)
    print("ok", len(files), "files", sum(i.tasks for i in ia), "subsequences")
    ctx.close()


if __name__ == "__main__":
    main()
