"""The drop-in, executed: the reference's OWN reader (XJPEG_DECODE_CTX_VTBL of src/jpeg_wrap.c:254-358,
compiled unmodified into oracle/_ref) is plugged into the CUDA backend with cuda_decode_set_frontend()
-- the one line INTEGRATION.md section 2 asks a maintainer to add -- and the reference's own protocol is
walked on the reference's own surface (image_init of src/image.c):

    alloc -> header -> image_init -> image(YUV|RGB)              src/jpeg_gpu.c:612-704
    reset -> header -> image                (steady state)       src/jpeg_gpu.c:1231-1237

with both upload formats (QUANT planes, PACK stream).  Done = planes bit-exact with what
xjpeg_decode_image(YUV) wrote for the same file (goldens AND the live reference table), pixels equal to
the colour oracle's, on the 11 small goldens + BASELINE config 1's 512x512 grey file + a 1080p 4:2:0 file."""
import ctypes as C

import numpy as np
import pytest

from golden_util import NAMES, load
from jpeg_gpu_b200 import _capi
from ref_reader import BIG_NAMES, RefLib, Session, load_big, sha

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref(gpu_ctx, reference):
    return RefLib()


@pytest.fixture()
def plugged(ref):
    """CUDA backend with the reference's reader as its front end; restored afterwards."""
    L = _capi.lib()
    L.cuda_decode_set_frontend(ref.xjpeg_address)
    try:
        yield _capi.vtbl("CUDA_DECODE_CTX_VTBL")
    finally:
        L.cuda_decode_set_frontend(None)
        assert L.cuda_decode_set_upload(_capi.JPEG_DECODE_QUANT) == 0


def expected(name):
    if name in BIG_NAMES:
        jpg, z = load_big(name)
    else:
        jpg, z, _ = load(name)
    want = {k: (str(z["sha_" + k]) if "sha_" + k in z.files else sha(z[k])) for k in ("yuv", "rgb", "quant")}
    return jpg, want


@pytest.mark.parametrize("upload", ["quant", "pack"])
@pytest.mark.parametrize("name", NAMES + BIG_NAMES)
def test_reference_reader_drives_the_cuda_backend(ref, plugged, name, upload):
    jpg, want = expected(name)
    assert _capi.lib().cuda_decode_set_upload(_capi.OUT_NAMES[upload]) == 0
    # what the reference's table itself writes for this file, live
    with Session(ref.xjpeg, jpg, ref.lib.image_init, ref.lib.image_zero, ref.lib.image_clear) as r:
        assert r.header() == 0 and r.image("yuv") == 0
        live_planes = r.planes()
    assert sha(live_planes) == want["yuv"]
    with Session(plugged, jpg, ref.lib.image_init, ref.lib.image_zero, ref.lib.image_clear) as s:
        # first frame
        assert s.header() == 0
        assert s.image("yuv") == 0
        assert np.array_equal(s.planes(), live_planes)
        # steady state: the benchmark loop of the reference
        for it in range(3):
            s.reset()
            assert s.header() == 0
            assert s.image("rgb") == 0
            assert sha(s.pixels()) == want["rgb"], (name, upload, it)
        s.reset()
        assert s.header() == 0
        assert s.image("yuv") == 0
        assert np.array_equal(s.planes(), live_planes)
        # the CPU-side formats are the front end's: the reference's reader writes them itself
        s.reset()
        assert s.header() == 0
        assert s.image("quant") == 0
        assert sha(s.coef()) == want["quant"]


def test_reference_reader_errors_keep_the_convention(ref, plugged):
    """A file without a frame header: the reference's parser reports "Error reading jpeg headers"
    (src/jpeg_wrap.c:275-278) -> EXIT_FAILURE through our table too; an unsupported `out` is refused
    with EXIT_FAILURE; nothing aborts.  (A cut-off file is NOT fed to the reference's reader: built
    with its default flags -- GLJ_ENABLE_VALIDATION off, Makefile:25 -- it divides by zero on one.)"""
    jpg, _ = expected("c420_64x48")
    with Session(plugged, bytes([0xff, 0xd8, 0xff, 0xd9]), ref.lib.image_init, ref.lib.image_zero, ref.lib.image_clear) as s:
        assert s.header() == 1
    with Session(plugged, jpg, ref.lib.image_init, ref.lib.image_zero, ref.lib.image_clear) as s:
        assert s.header() == 0
        assert s.vt.decode_image(s.dec, C.byref(s.img), 9) == 1


def test_explicit_options_instead_of_the_thread_knobs(ref):
    """cuda_decode_alloc_ex: the reference's reader and the PACK upload chosen per context, nothing
    set on the thread; the other four slots are the table's."""
    jpg, want = expected("c420_rst_80x48")
    L = _capi.lib()
    vt = _capi.vtbl("CUDA_DECODE_CTX_VTBL")
    opt = _capi.cuda_decode_options(ref.xjpeg_address, 0, _capi.JPEG_DECODE_PACK, 0)

    class Explicit:   # a table whose alloc slot passes the options
        decode_header, decode_image, decode_reset, decode_free = vt.decode_header, vt.decode_image, vt.decode_reset, vt.decode_free

        @staticmethod
        def decode_alloc(info):
            return L.cuda_decode_alloc_ex(info, C.byref(opt))

    with Session(Explicit, jpg, ref.lib.image_init, ref.lib.image_zero, ref.lib.image_clear) as s:
        assert s.header() == 0 and s.image("rgb") == 0
        assert sha(s.pixels()) == want["rgb"]
        s.reset()
        assert s.header() == 0 and s.image("yuv") == 0
        assert sha(s.planes()) == want["yuv"]
