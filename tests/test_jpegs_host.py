"""Host side of the JPEG-files entry points: header probing (no GPU) and the restart-interval
split of the entropy front end against the sequential reader."""
import io

import numpy as np
import pytest

import jpeg_gpu_b200 as J
from golden_util import NAMES, load


def test_probe_reports_headers_and_packs_offsets():
    files = [load(n)[0] for n in NAMES]
    total, infos = J.probe_jpegs(files)
    off = 0
    for name, inf in zip(NAMES, infos):
        _, z, _ = load(name)
        assert inf.status == 0 and inf.message is None
        assert (inf.width, inf.height, inf.ncomps) == (int(z["hdr_width"]), int(z["hdr_height"]), int(z["hdr_ncomps"]))
        assert (inf.hsamp0, inf.vsamp0) == (int(z["hdr_hsamp"][0]), int(z["hdr_vsamp"][0]))
        assert inf.restart_interval == int(z["hdr_restart_interval"])
        assert inf.rgb_off == off and inf.rgb_off % 256 == 0
        assert inf.rgb_len == inf.width * inf.height * (1 if inf.ncomps == 1 else 3)
        off += -(-inf.rgb_len // 256) * 256
    assert total == off


def test_probe_rejects_garbage_without_failing_the_batch(capfd):
    good = load("c420_64x48")[0]
    total, infos = J.probe_jpegs([b"not a jpeg", good, good[:40], b""])
    assert [i.status for i in infos] == [1, 0, 1, 1]
    assert infos[1].rgb_off == 0 and total == -(-infos[1].rgb_len // 256) * 256
    assert all(i.message for i in infos if i.status)
    capfd.readouterr()   # the front end reports each bad header on stderr, like the reference


def _segments_decode(jpg, order):
    """Decodes a file restart interval by restart interval, in the given order, through the
    front end's internal split entry points (jgpu_front.h)."""
    import ctypes as C
    from jpeg_gpu_b200 import _capi
    L = _capi.lib()

    class Segs(C.Structure):
        _fields_ = [("nseg", C.c_int), ("mcus_per_seg", C.c_int), ("total_mcus", C.c_int), ("seg_pos", C.POINTER(C.c_int))]

    L.jfront_find_segments.argtypes = [C.c_void_p, C.POINTER(Segs)]
    L.jfront_decode_segments.argtypes = [C.c_void_p, C.POINTER(_capi.image), C.c_int, C.POINTER(Segs), C.c_int, C.c_int,
                                         C.POINTER(C.c_char_p)]
    L.jfront_segments_free.argtypes = [C.POINTER(Segs)]
    v = _capi.vtbl("JFRONT_DECODE_CTX_VTBL")
    buf = (C.c_ubyte * len(jpg)).from_buffer_copy(jpg)
    info = _capi.jpeg_info(len(jpg), C.cast(buf, C.POINTER(C.c_ubyte)))
    ctx = v.decode_alloc(C.byref(info))
    hdr = _capi.jpeg_header()
    assert v.decode_header(ctx, C.byref(hdr)) == 0
    img = _capi.image()
    assert L.jgpu_image_init(C.byref(img), C.byref(hdr)) == 0
    L.jgpu_image_zero(C.byref(img))
    segs = Segs()
    assert L.jfront_find_segments(ctx, C.byref(segs)) == 0
    n = segs.nseg
    err = C.c_char_p()
    for s in order(n):
        assert L.jfront_decode_segments(ctx, C.byref(img), _capi.JPEG_DECODE_QUANT, C.byref(segs), s, s + 1, C.byref(err)) == 0, err.value
    blocks = sum(((img.plane[i].width >> 3) << img.plane[i].xdec) * img.plane[i].cstride for i in range(img.nplanes))
    coef = np.ctypeslib.as_array(img.coef, shape=(blocks * 64,)).copy()
    L.jfront_segments_free(C.byref(segs))
    L.jgpu_image_clear(C.byref(img))
    v.decode_free(ctx)
    return n, coef


@pytest.mark.parametrize("name", NAMES)
def test_restart_intervals_decode_independently(name):
    """Interval by interval, last first: same QUANT planes as the reference's sequential reader."""
    jpg, z, _ = load(name)
    n, coef = _segments_decode(jpg, lambda n: range(n - 1, -1, -1))
    ri = int(z["hdr_restart_interval"])
    if ri:
        hmax, vmax = int(max(z["hdr_hsamp"])), int(max(z["hdr_vsamp"]))
        nmcu = -(-int(z["hdr_width"]) // (8 * hmax)) * -(-int(z["hdr_height"]) // (8 * vmax))
        assert n == -(-nmcu // ri) and n > 1
    else:
        assert n == 1
    assert np.array_equal(coef, z["quant"])


def test_malformed_restart_markers_are_left_to_the_sequential_reader():
    import ctypes as C
    from jpeg_gpu_b200 import _capi
    jpg = bytearray(load("c420_rst_80x48")[0])
    # renumber the second RST marker: RST1 -> RST5
    pos = [i for i in range(len(jpg) - 1) if jpg[i] == 0xFF and 0xD0 <= jpg[i + 1] <= 0xD7]
    jpg[pos[1] + 1] = 0xD5
    with pytest.raises(AssertionError):
        _segments_decode(bytes(jpg), lambda n: range(n))


def test_batches_of_any_buffer_type_marshal_alike():
    """The wrapper fills the jgpu_jpeg array column-wise (addresses taken in one C call, sizes through numpy):
    bytes, bytearray and memoryview inputs, empty entries and a batch of hundreds give what one-by-one probing gives."""
    base = [load(n)[0] for n in NAMES]
    files = []
    for i in range(300):
        f = base[i % len(base)]
        files.append(f if i % 3 == 0 else bytearray(f) if i % 3 == 1 else memoryview(f))
    files[17] = b""
    total, infos = J.probe_jpegs(files)
    off = 0
    for i, inf in enumerate(infos):
        if i == 17:
            assert inf.status == 1 and inf.message
            continue
        _, one = J.probe_jpegs([bytes(files[i])])
        assert (inf.status, inf.width, inf.height, inf.ncomps, inf.rgb_len) == \
               (0, one[0].width, one[0].height, one[0].ncomps, one[0].rgb_len)
        assert inf.rgb_off == off
        off += -(-inf.rgb_len // 256) * 256
    assert total == off
