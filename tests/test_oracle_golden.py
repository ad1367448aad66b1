"""The CPU oracle against the golden vectors the compiled reference produced
(tests/golden/, see make_golden.py), and against the reference's own tests."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from golden_util import NAMES, blocks, load


def test_fixtures_present():
    assert len(NAMES) >= 10


def test_port_idct_matches_reference_blocks(port):
    """oracle_idct.c == glj_real_idct8x8 (src/dct.c:100-121) on 3501 stored blocks, bit for bit."""
    coef, want = blocks()
    assert np.array_equal(port.idct_blocks(coef), want)


def test_numpy_idct_matches_reference_blocks():
    coef, want = blocks()
    assert np.array_equal(oracle.np_idct8x8(coef), want)


@pytest.mark.parametrize("name", NAMES)
def test_port_pipeline_matches_reference_decode(port, name):
    """quant planes -> (port) dequant+IDCT+clamp == xjpeg's YUV output; colour == stored RGB."""
    _, z, g = load(name)
    rgb, planes = port.decode_image(g, z["quant"], z["hdr_qtabs"], [int(v) for v in z["hdr_tq"]])
    got = np.concatenate([p.ravel() for p in planes])
    assert np.array_equal(got, z["yuv"])
    assert np.array_equal(rgb.reshape(-1), z["rgb"])


@pytest.mark.parametrize("name", NAMES)
def test_dct_planes_are_quant_times_table(name):
    """src/xjpeg.c:501-503,524-527: DCT output = (short)(quant * tbl)."""
    _, z, g = load(name)
    for p, t in zip(g.planes, z["hdr_tq"]):
        n = p.hblocks * p.vblocks
        q = z["quant"][p.coef_off:p.coef_off + 64 * n].reshape(n, 8, 8)
        d = z["dct"][p.coef_off:p.coef_off + 64 * n].reshape(n, 8, 8)
        assert np.array_equal(oracle.np_dequant(q, z["hdr_qtabs"][int(t)]), d)


def test_threaded_batch_equals_simple_loops(port):
    """jgo_decode_batch (strip-wise, threaded: the CPU baseline) == the plain whole-image loops."""
    rows, coefs, off_c, off_r = [], [], 0, 0
    items = []
    for name in NAMES:
        _, z, g = load(name)
        tq = [int(v) for v in z["hdr_tq"]]
        rows.append(oracle.make_desc(g, tq, off_c, off_r, len(items)))
        items.append((z, g, off_r))
        coefs.append(z["quant"])
        off_c += g.coef_len
        off_r += g.rgb_len
    qt = np.stack([z["hdr_qtabs"] for z, _, _ in items])
    rgb, _ = port.decode_batch(np.stack(rows), np.concatenate(coefs), qt, off_r, 0, nthreads=4)
    for z, g, o in items:
        assert np.array_equal(rgb[o:o + g.rgb_len], z["rgb"])


def test_reference_own_tests_pass():
    """test/dct.c (IEEE-1180, 31 checks) and test/image.c (12 checks) built from the mount."""
    if not oracle.have_reference():
        pytest.skip("oracle/_ref not built")
    for exe in ("test_dct", "test_image"):
        path = os.path.join(oracle.HERE, "_ref", exe)
        if not os.path.exists(path):
            pytest.skip(f"{exe} not built")
        r = subprocess.run([path], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "Errors:  0" in r.stdout + r.stderr


def test_port_matches_compiled_reference_live(port, reference):
    """Only in the build container: 200k random blocks, port vs the reference object."""
    rng = np.random.default_rng(5)
    blk = rng.integers(-2048, 2048, size=(200000, 8, 8), dtype=np.int16)
    assert np.array_equal(port.idct_blocks(blk), reference.ref_idct_blocks(blk))
    blk = rng.integers(-32768, 32768, size=(50000, 8, 8), dtype=np.int16)
    assert np.array_equal(port.idct_blocks(blk), reference.ref_idct_blocks(blk))
