"""The colour oracle (jgo_colour_offsets, from res/yuv.fs.glsl:11-23).  It is OUR definition
(the xjpeg backend has no RGB output, src/jpeg_wrap.c:321-342); these tests pin it and state
its distance from a straightforward float evaluation of the shader: <= 1 LSB per channel."""
import numpy as np

import oracle


def test_c_and_numpy_offsets_agree(port):
    cb, cr = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    ro, go, bo = oracle.np_colour_offsets(cb, cr)
    for c, r in [(0, 0), (255, 255), (128, 128), (0, 255), (255, 0), (17, 201), (200, 33), (127, 129)]:
        assert port.colour_offsets(c, r) == (int(ro[c, r]), int(go[c, r]), int(bo[c, r]))
    assert ro[128, 128] == go[128, 128] == bo[128, 128] == 0


def test_rounding_ties_are_resolved_to_even(port):
    """The only exact .5 offsets are B at |Cb-128| = 125 (1.772*125 = 221.5, a tie in real
    arithmetic too).  The oracle (lrintf) and the kernel (RN add of 1.5*2^23) both take the
    even neighbour, 222; R and G never tie."""
    cb, cr = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    cbf = cb.astype(np.float32) - np.float32(128)
    crf = cr.astype(np.float32) - np.float32(128)
    ties = {}
    for name, v in (("r", np.float32(1.402) * crf), ("b", np.float32(1.772) * cbf),
                    ("g", np.float32(-0.34414) * cbf + np.float32(-0.71414) * crf)):
        f = v.astype(np.float64) - np.floor(v.astype(np.float64))
        ties[name] = sorted(set(int(c) for c in cb[f == 0.5])) if name == "b" else int((f == 0.5).sum())
    assert ties == {"r": 0, "g": 0, "b": [3, 253]}
    assert port.colour_offsets(3, 128)[2] == -222 and port.colour_offsets(253, 128)[2] == 222
    magic = np.float32(12582912.0)
    for c in (3, 253):
        bits = (np.float32(1.772) * np.float32(c - 128) + magic).view(np.uint32) & 0xffff
        assert np.int16(bits) == port.colour_offsets(c, 128)[2]


def test_within_one_lsb_of_float_shader_evaluation():
    """yuv.fs.glsl evaluates mat3*vec3 in float and GL rounds float->unorm8; whatever order
    or FMA fusion a GL implementation picks stays within 1 LSB of the oracle.
    Tolerance: 1 (stated by BASELINE.json's north_star for the GLSL comparator)."""
    rng = np.random.default_rng(3)
    y = rng.integers(0, 256, size=200000).astype(np.float64)
    cb = rng.integers(0, 256, size=200000)
    cr = rng.integers(0, 256, size=200000)
    ro, go, bo = oracle.np_colour_offsets(cb, cr)
    ours = np.clip(np.stack([y + ro, y + go, y + bo], -1), 0, 255)
    u, v = cb - 128.0, cr - 128.0
    exact = np.stack([y + 1.402 * v, y - 0.34414 * u - 0.71414 * v, y + 1.772 * u], -1)
    shader = np.rint(np.clip(exact, 0, 255))
    assert np.abs(ours - shader).max() <= 1
    assert (ours != shader).mean() < 1e-3
