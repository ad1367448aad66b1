"""Writes a synthetic baseline JPEG (the bench's picture): python tools/make_jpeg.py out.jpg [w h subsampling rst quality]"""
import sys

import numpy as np
from PIL import Image


def main():
    out = sys.argv[1]
    w, h, ss, rst, q = (int(v) for v in (sys.argv[2:7] + ["3840", "2160", "2", "240", "85"][len(sys.argv) - 2:]))
    rng = np.random.default_rng(20261017)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 5 + yy * 3) % 256, (yy * 7 + xx) % 256, (xx * 2 + yy * 9) % 256], -1)
    pic = np.clip(base + rng.integers(-24, 25, size=base.shape), 0, 255).astype(np.uint8)
    Image.fromarray(pic).save(out, "JPEG", quality=q, subsampling=ss, restart_marker_blocks=rst)


if __name__ == "__main__":
    main()
