"""Access to the committed golden vectors (tests/golden/, made by make_golden.py from the
compiled reference)."""
import glob
import os

import numpy as np

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.jpg")))


def load(name):
    with open(os.path.join(GOLDEN, name + ".jpg"), "rb") as f:
        jpg = f.read()
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = oracle.geometry(int(z["hdr_width"]), int(z["hdr_height"]), [int(v) for v in z["hdr_hsamp"]],
                        [int(v) for v in z["hdr_vsamp"]])
    return jpg, z, g


def blocks():
    z = np.load(os.path.join(GOLDEN, "blocks.npz"))
    return z["coef"], z["idct"]
