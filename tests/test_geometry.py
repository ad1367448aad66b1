"""Host geometry (jgpu_layout_query, C) vs the oracle's restatement of image_init
(src/image.c:24-97) and, in the build container, vs the reference's image_init itself."""
import numpy as np
import pytest

import jpeg_gpu_b200 as J
import oracle

SIZES = [(1, 1), (8, 8), (17, 33), (32, 24), (70, 50), (512, 512), (1000, 563), (1920, 1080), (3840, 2160),
         (65504, 16), (16, 65504)]


@pytest.mark.parametrize("ss", sorted(J.SUBSAMPLINGS))
def test_layout_matches_oracle(port, ss):
    hs, vs = J.SUBSAMPLINGS[ss]
    for w, h in SIZES:
        lay = J.ImageDesc(w, h, hs, vs, tq=(0, 1, 1)[:len(hs)]).query_layout()
        for g in (oracle.geometry(w, h, hs, vs), port.geometry(w, h, hs, vs)):
            assert (lay.nhmb, lay.nvmb, lay.coef_len, lay.data_len, lay.rgb_len) == \
                   (g.nhmb, g.nvmb, g.coef_len, g.data_len, g.rgb_len)
            for a, b in zip(lay.planes, g.planes):
                assert (a.hblocks, a.vblocks, a.width, a.height, a.xdec, a.ydec, a.cstride, a.coef_off, a.data_off) == \
                       (b.hblocks, b.vblocks, b.width, b.height, b.xdec, b.ydec, b.cstride, b.coef_off, b.data_off)


def test_reference_image_test_case():
    """test/image.c:21-55: 32x24 4:2:0 -> xdec/ydec/ystride per plane."""
    lay = J.ImageDesc(32, 24, (2, 1, 1), (2, 1, 1)).query_layout()
    assert [(p.xdec, p.ydec, p.width) for p in lay.planes] == [(0, 0, 32), (1, 1, 16), (1, 1, 16)]
    assert [(p.hblocks, p.vblocks) for p in lay.planes] == [(4, 4), (2, 2), (2, 2)]


def test_baseline_config_sizes():
    """bytes per image at the BASELINE configs."""
    def coded(w, h, ss):
        hs, vs = J.SUBSAMPLINGS[ss]
        return J.ImageDesc(w, h, hs, vs, tq=(0, 1, 1)[:len(hs)]).query_layout()
    assert coded(512, 512, "gray").coded_blocks * 128 == 524288
    assert coded(1920, 1080, "420").coded_blocks * 128 == 6266880
    lay = coded(3840, 2160, "420")
    assert lay.coded_blocks * 128 == 24883200 and lay.coef_len * 2 == 24944640 and lay.rgb_len == 24883200
    assert coded(3840, 2160, "422").coded_blocks * 128 == 33177600
    assert coded(3840, 2160, "444").coded_blocks * 128 == 49766400
    assert coded(1920, 1080, "420").planes[0].height == 1088


def test_layout_matches_reference_image_init(reference):
    for ss, (hs, vs) in J.SUBSAMPLINGS.items():
        for w, h in SIZES[:-2]:
            lay = J.ImageDesc(w, h, hs, vs, tq=(0, 1, 1)[:len(hs)]).query_layout()
            ref = reference.ref_layout(w, h, hs, vs)
            assert lay.coef_len == ref[0] * 64
            for i, p in enumerate(lay.planes):
                o = ref[1 + 8 * i:]
                assert (p.width, p.height, p.xdec, p.ydec, p.width, p.cstride, p.coef_off) == tuple(int(v) for v in o[:7])


def test_rejections():
    for bad in [dict(width=0, height=8), dict(width=8, height=70000), dict(hsamp=(3, 1, 1)), dict(hsamp=(1, 2, 1)),
                dict(tq=(0, 1, 7)),
                dict(width=65535, height=16)]:  # pads to 65536: image_plane.width is an unsigned short
        kw = dict(width=16, height=16, hsamp=(2, 1, 1), vsamp=(2, 1, 1), tq=(0, 1, 1))
        kw.update(bad)
        with pytest.raises(ValueError):
            J.ImageDesc(**kw).query_layout()


def test_algorithmic_bytes_per_pixel():
    """3 B/px grey, 6 (4:2:0), 7 (4:2:2), 9 (4:4:4) at MCU-aligned sizes."""
    for ss, bpp in [("gray", 3), ("420", 6), ("422", 7), ("444", 9)]:
        hs, vs = J.SUBSAMPLINGS[ss]
        lay = J.ImageDesc(3840, 2160 if ss != "420" else 2160 - 2160 % 16, hs, vs, tq=(0, 1, 1)[:len(hs)]).query_layout()
        px = 3840 * (2160 if ss != "420" else 2160 - 2160 % 16)
        assert (128 * lay.coded_blocks + lay.rgb_len) == bpp * px
