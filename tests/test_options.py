"""Per-thread decoder options (no process-wide knobs), the explicit allocator's argument checks, and
layout checks that protect the colour stage (host logic only: no GPU needed)."""
import ctypes as C
import threading

import jpeg_gpu_b200 as J
from jpeg_gpu_b200 import _capi


def options():
    o = _capi.cuda_decode_options()
    _capi.lib().cuda_decode_get_options(C.byref(o))
    return o.frontend, o.device, o.upload, o.entropy_on_device


def test_setters_are_per_thread():
    L = _capi.lib()
    base = options()
    assert base == (None, -1, _capi.JPEG_DECODE_QUANT, -1)
    seen = {}

    def other():
        L.cuda_decode_set_device(3)
        assert L.cuda_decode_set_upload(_capi.JPEG_DECODE_PACK) == 0
        assert L.cuda_decode_set_entropy(1) == 0
        seen["other"] = options()

    t = threading.Thread(target=other)
    t.start()
    t.join()
    assert seen["other"] == (None, 3, _capi.JPEG_DECODE_PACK, 1)
    assert options() == base, "another thread's settings leaked into this one"
    assert L.cuda_decode_set_upload(_capi.JPEG_DECODE_RGB) == 1 and options() == base
    assert L.cuda_decode_set_entropy(5) == 1 and options() == base


def test_alloc_ex_checks_its_options():
    L = _capi.lib()
    data = (C.c_ubyte * 4)(0xFF, 0xD8, 0xFF, 0xD9)
    info = _capi.jpeg_info(4, C.cast(data, C.POINTER(C.c_ubyte)))
    bad = _capi.cuda_decode_options(None, 0, _capi.JPEG_DECODE_YUV, 0)
    assert not L.cuda_decode_alloc_ex(C.byref(info), C.byref(bad))
    assert b"upload" in L.jgpu_last_error()
    assert not L.cuda_decode_alloc_ex(C.byref(info), None)
    good = _capi.cuda_decode_options(None, 0, _capi.JPEG_DECODE_PACK, 0)
    dec = L.cuda_decode_alloc_ex(C.byref(info), C.byref(good))
    assert dec
    _capi.vtbl("CUDA_DECODE_CTX_VTBL").decode_free(dec)


def test_component_0_must_be_the_most_finely_sampled_both_ways():
    d = J.ImageDesc(64, 64, (1, 1, 1), (1, 2, 2))    # luma 1x1 under chroma 1x2
    try:
        d.query_layout()
    except ValueError as e:
        assert "vertical" in str(e)
    else:
        raise AssertionError("a vertically decimated component 0 was accepted")
    d = J.ImageDesc(64, 64, (1, 2, 2), (1, 1, 1))
    try:
        d.query_layout()
    except ValueError as e:
        assert "horizontal" in str(e)
    else:
        raise AssertionError("a horizontally decimated component 0 was accepted")


def test_surface_flags_without_a_gpu():
    """jgpu_image_init_ex(PINNED) falls back to ordinary memory when page-locking is not possible,
    and jgpu_image_clear frees whichever kind it got."""
    L = _capi.lib()
    hdr = _capi.jpeg_header()
    hdr.bits, hdr.width, hdr.height, hdr.ncomps = 8, 32, 24, 1
    hdr.comp[0].hsamp = hdr.comp[0].vsamp = 1
    hdr.comp[0].hblocks, hdr.comp[0].vblocks = 4, 3
    for flags in (0, _capi.JGPU_IMAGE_PINNED):
        img = _capi.image()
        assert L.jgpu_image_init_ex(C.byref(img), C.byref(hdr), flags) == 0
        assert img.pixels and img.width == 32
        L.jgpu_image_clear(C.byref(img))
