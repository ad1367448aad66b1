"""tools/jpeg_gpu_cli.c: the reference's command line (src/jpeg_gpu.c:473-506,614-700) headless.
Text formats follow the reference's printf formats so dumps are diffable."""
import os
import re
import subprocess

import numpy as np
import pytest

from golden_util import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "jpeg_gpu_b200", "jpeg_gpu_cli")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=120)


def numbers(text):
    return np.array([int(v) for v in re.findall(r"-?\d+", text)])


def test_header_dump():
    r = run("-H", os.path.join(GOLDEN, "c420_rst_80x48.jpg"))
    assert r.returncode == 0
    assert "Image Size         : 80x48" in r.stdout
    assert "Chroma Subsampling : 4:2:0" in r.stdout
    assert "Minimum Coded Unit : 2x2 1x1 1x1" in r.stdout
    assert "Restart Interval   : 2" in r.stdout
    _, z, _ = load("c420_rst_80x48")
    tables = r.stdout.split("Quant Table 0 Bits : 8\n")[1]
    assert np.array_equal(numbers(tables.split("Quant Table 1")[0]), z["hdr_qtabs"][0])


@pytest.mark.parametrize("name,out,key", [("gray_48x40", "quant", "quant"), ("c422_56x24", "dct", "dct")])
def test_coefficient_dump_matches_reference_planes(name, out, key):
    _, z, g = load(name)
    r = run("-i", "jfront", "-o", out, "-d", os.path.join(GOLDEN, name + ".jpg"))
    assert r.returncode == 0
    planes = r.stdout.split("Plane ")[1:]
    assert len(planes) == len(g.planes)
    for text, p in zip(planes, g.planes):
        got = numbers(text.split("\n", 1)[1])
        assert np.array_equal(got, z[key][p.coef_off:p.coef_off + p.width * p.height])


def test_bad_arguments_and_missing_gpu():
    assert run("-i", "nvjpeg", os.path.join(GOLDEN, "gray_48x40.jpg")).returncode != 0
    assert run("-o", "png", os.path.join(GOLDEN, "gray_48x40.jpg")).returncode != 0
    assert run().returncode != 0
    r = run("-i", "jfront", "-o", "rgb", os.path.join(GOLDEN, "gray_48x40.jpg"))
    assert r.returncode != 0 and "Unsupported output" in r.stderr


@pytest.mark.gpu
def test_cuda_rgb_and_yuv_dump(gpu_ctx):
    _, z, g = load("c420_odd_70x50")
    path = os.path.join(GOLDEN, "c420_odd_70x50.jpg")
    r = run("-i", "cuda", "-o", "rgb", "-d", path)
    assert r.returncode == 0, r.stderr
    chans = [numbers(t.split("\n", 1)[1]) for t in r.stdout.split("Plane ")[1:]]
    rgb = z["rgb"].reshape(50, 70, 3)
    for c in range(3):
        assert np.array_equal(chans[c], rgb[..., c].reshape(-1))
    r = run("-i", "cuda", "-o", "yuv", "-d", path)
    got = np.concatenate([numbers(t.split("\n", 1)[1]) for t in r.stdout.split("Plane ")[1:]])
    assert np.array_equal(got, z["yuv"])
    r = run("-i", "cuda", "-o", "rgb", "-n", "20", path)
    assert r.returncode == 0 and "20 frames" in r.stdout
