"""BASELINE config 1: "single 512x512 baseline grayscale JPEG via xjpeg CPU path (jpeg_wrap plumbing,
no GPU)".  The reference's own XJPEG_DECODE_CTX_VTBL (src/jpeg_wrap.c:254-358, compiled unmodified into
oracle/_ref) is walked through alloc -> header -> image_init -> image(YUV) with the reference's own
image_init, and pinned against the committed goldens; the product's CPU front end (JFRONT_DECODE_CTX_VTBL)
must fill the same surface with the same QUANT / PACK buffers."""
import ctypes as C

import numpy as np
import pytest

from golden_util import NAMES, load
from jpeg_gpu_b200 import _capi
from ref_reader import BIG_NAMES, RefLib, Session, load_big, sha


@pytest.fixture(scope="module")
def ref(reference):
    return RefLib()


def test_config1_gray_512_through_the_reference_table(ref):
    jpg, z = load_big("gray_512x512")
    with Session(ref.xjpeg, jpg, ref.lib.image_init, ref.lib.image_zero, ref.lib.image_clear) as s:
        assert s.header() == 0
        h = s.hdr
        assert (h.width, h.height, h.ncomps, h.bits) == (512, 512, 1, 8)
        assert _capi.SUBSAMP_NAMES[h.subsamp] == "Mono"
        assert (h.comp[0].hblocks, h.comp[0].vblocks) == (64, 64)
        assert np.array_equal(np.ctypeslib.as_array(h.quant[0].tbl), z["hdr_qtabs"][0])
        assert s.image("yuv") == 0
        assert np.array_equal(s.planes(), z["yuv"])
        # steady state (src/jpeg_gpu.c:1231-1237)
        for out, key in (("quant", "quant"), ("yuv", "yuv")):
            s.reset()
            assert s.header() == 0
            assert s.image(out) == 0
            got = s.coef() if out == "quant" else s.planes()
            assert np.array_equal(got, z[key])
        s.reset()
        assert s.header() == 0
        assert s.image("rgb") == 1          # the xjpeg backend has no RGB output (src/jpeg_wrap.c:335-339)


@pytest.mark.parametrize("name", BIG_NAMES)
def test_product_front_end_fills_the_reference_surface_identically(ref, name):
    """JFRONT (the product's CPU reader) and XJPEG (the reference's) on the reference's own surface."""
    jpg, z = load_big(name)
    jfront = _capi.vtbl("JFRONT_DECODE_CTX_VTBL")
    res = {}
    for tag, vt in (("ref", ref.xjpeg), ("ours", jfront)):
        with Session(vt, jpg, ref.lib.image_init, ref.lib.image_zero, ref.lib.image_clear) as s:
            assert s.header() == 0
            assert s.image("quant") == 0
            res[tag] = s.coef()
    assert sha(res["ref"]) == str(z["sha_quant"])
    assert np.array_equal(res["ref"], res["ours"])


@pytest.mark.parametrize("name", NAMES)
def test_small_goldens_through_the_reference_table(ref, name):
    jpg, z, g = load(name)
    with Session(ref.xjpeg, jpg, ref.lib.image_init, ref.lib.image_zero, ref.lib.image_clear) as s:
        assert s.header() == 0
        assert s.image("yuv") == 0
        assert np.array_equal(s.planes(), z["yuv"])
