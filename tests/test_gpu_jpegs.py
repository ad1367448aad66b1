"""GPU: JPEG files in, RGB out (jgpu_decode_jpegs): multi-threaded entropy front end (images and
restart intervals in parallel) + fused kernel, against the oracle fed with the reference
reader's own QUANT planes (goldens) or the sequential front end's.  Bit-exact."""
import io

import numpy as np
import pytest

import jpeg_gpu_b200 as J
import oracle
from golden_util import NAMES, load

pytestmark = pytest.mark.gpu


def _expected_rgb(checker, jpg):
    """Sequential front end -> QUANT planes -> CPU oracle."""
    with J.Decoder(jpg, impl="jfront") as dec:
        h = dec.decode_header()
        quant = dec.decode_image("quant")["coef"]
    g = oracle.geometry(h.width, h.height, h.hsamp, h.vsamp)
    rgb, _ = checker.decode_image(g, quant, h.qtabs, h.tq, nthreads=8)
    return rgb


@pytest.mark.parametrize("entropy", ["cpu", "gpu"])
def test_golden_files_decode_to_the_oracles_pixels(gpu_ctx, entropy):
    files = [load(n)[0] for n in NAMES]
    rgb, infos = gpu_ctx.decode_jpegs(files, nthreads=4, entropy=entropy)
    for name, inf in zip(NAMES, infos):
        _, z, _ = load(name)
        assert inf.status == 0
        got = rgb[inf.rgb_off:inf.rgb_off + inf.rgb_len]
        assert np.array_equal(got, z["rgb"]), name


def _big_jpegs():
    from PIL import Image
    rng = np.random.default_rng(7)
    out = []
    for (w, h, ss, rst) in [(1920, 1080, 2, 0), (1920, 1080, 2, 120), (1000, 563, 1, 7), (2048, 1536, 0, 256),
                            (1537, 771, 2, 97), (640, 480, "L", 80), (3840, 2160, 2, 240)]:
        yy, xx = np.mgrid[0:h, 0:w]
        base = np.stack([(xx * 5 + yy * 3) % 256, (yy * 7 + xx) % 256, (xx * 2 + yy * 9) % 256], -1)
        img = np.clip(base + rng.integers(-40, 41, size=base.shape), 0, 255).astype(np.uint8)
        bio = io.BytesIO()
        if ss == "L":
            Image.fromarray(img[..., 0]).save(bio, "JPEG", quality=85, restart_marker_blocks=rst)
        else:
            Image.fromarray(img).save(bio, "JPEG", quality=85, subsampling=ss, restart_marker_blocks=rst)
        out.append(bio.getvalue())
    return out


def test_large_files_split_by_restart_interval(gpu_ctx, checker):
    pytest.importorskip("PIL")
    files = _big_jpegs()
    for nthreads in (1, 16):
        rgb, infos = gpu_ctx.decode_jpegs(files, nthreads=nthreads, entropy="cpu")
        assert infos[0].tasks == 1 and infos[1].tasks > 1 and infos[6].tasks > 8
        for jpg, inf in zip(files, infos):
            assert inf.status == 0
            want = _expected_rgb(checker, jpg)
            got = rgb[inf.rgb_off:inf.rgb_off + inf.rgb_len].reshape(inf.shape)
            assert np.array_equal(got, want)


@pytest.mark.parametrize("entropy", ["cpu", "gpu"])
def test_bad_files_do_not_stop_the_batch(gpu_ctx, capfd, entropy):
    good = load("c420_64x48")[0]
    cut = load("c420_rst_80x48")[0][:600]          # scan runs out: decodes what is there, like the reference
    files = [good, b"junk", good, cut]
    rgb, infos = gpu_ctx.decode_jpegs(files, strict=False, entropy=entropy)
    assert [i.status for i in infos][:3] == [0, 1, 0]
    z = load("c420_64x48")[1]
    for i in (0, 2):
        assert np.array_equal(rgb[infos[i].rgb_off:infos[i].rgb_off + infos[i].rgb_len], z["rgb"])
    with pytest.raises(RuntimeError):
        gpu_ctx.decode_jpegs([b"junk", b""], entropy=entropy)
    capfd.readouterr()
