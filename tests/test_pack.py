"""The PACK (zero-run packed) coefficient format, host side, against what the reference's own
reader wrote for the golden files (tests/golden/*.npz `pack`, `index`, `packed`: xjpeg
JPEG_DECODE_PACK, src/xjpeg.c:484-496,513-519,531-535) and against the consumer
res/horz_pack_yuv.fs.glsl:105-127 as restated in oracle/oracle_pack.c."""
import numpy as np
import pytest

import jpeg_gpu_b200 as J
from jpeg_gpu_b200 import synth
from jpeg_gpu_b200.batch import pack_batch_streams, pack_from_quant
from golden_util import NAMES, load


def _desc(z):
    n = int(z["hdr_ncomps"])
    return J.ImageDesc(int(z["hdr_width"]), int(z["hdr_height"]), [int(v) for v in z["hdr_hsamp"][:n]],
                       [int(v) for v in z["hdr_vsamp"][:n]], tq=[int(v) for v in z["hdr_tq"][:n]])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_consumer_expands_reference_pack_to_reference_quant(port, name):
    _, z, g = load(name)
    assert np.array_equal(port.unpack_image(g, z["pack"], z["index"]), z["quant"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_producer_matches_reference_pack(port, name):
    _, z, g = load(name)
    pack, index, packed = port.pack_image(g, z["quant"])
    assert np.array_equal(pack, z["pack"])
    assert np.array_equal(index, z["index"])
    assert np.array_equal(packed, z["packed"])


@pytest.mark.parametrize("name", NAMES)
def test_product_packer_matches_reference_pack(name):
    """jgpu_pack_from_quant (host utility of the product) reproduces the reference's stream."""
    _, z, _ = load(name)
    pack, index = pack_from_quant(_desc(z), z["quant"])
    assert np.array_equal(pack, z["pack"])
    assert np.array_equal(index, z["index"])


@pytest.mark.parametrize("name", NAMES)
def test_front_end_pack_matches_reference_pack(name):
    """JFRONT_DECODE_CTX_VTBL in PACK mode writes the words, index and counts xjpeg writes."""
    jpg, z, _ = load(name)
    with J.Decoder(jpg, impl="jfront") as dec:
        dec.decode_header()
        r = dec.decode_image("pack")
    assert np.array_equal(r["pack"].view(np.uint16), z["pack"])
    assert np.array_equal(r["index"], z["index"])
    assert list(r["packed"]) == list(z["packed"])


@pytest.mark.parametrize("ss", ["gray", "444", "422", "420", "440", "411"])
def test_pack_round_trip_synthetic(port, ss):
    """Dense -> PACK -> dense on synthetic planes, product packer vs oracle producer and consumer.
    Includes ZRL runs (impulse blocks), blocks that run to coefficient 63 (dense) and empty blocks."""
    hs, vs = J.SUBSAMPLINGS[ss]
    q = synth.quality_tables(85)
    for kind in ["natural", "dense", "dc", "zero", "impulse"]:
        d = J.ImageDesc(70, 50, hs, vs, tq=(0, 1, 1)[:len(hs)])
        coef_len, _, _ = J.pack_batch([d])
        coef = synth.batch_coefficients([d], coef_len, q, kinds=[kind])
        g = port.geometry(70, 50, hs, vs)
        pack, index = pack_from_quant(d, coef)
        opack, oindex, _ = port.pack_image(g, coef)
        assert np.array_equal(pack, opack) and np.array_equal(index, oindex), (ss, kind)
        back = port.unpack_image(g, pack, index)
        # padding blocks of the reference layout are not coded: compare plane by plane
        for p in g.planes:
            n = 64 * p.hblocks * p.vblocks
            assert np.array_equal(back[p.coef_off:p.coef_off + n], coef[p.coef_off:p.coef_off + n]), (ss, kind)


def test_pack_truncates_to_twelve_bits_like_the_reference(port):
    """pack word = value & 0xfff (src/xjpeg.c:493,516): values outside [-2048, 2047] alias."""
    d = J.ImageDesc(8, 8, (1,), (1,), tq=(0,))
    coef = np.zeros(64, dtype=np.int16)
    coef[0], coef[1], coef[63] = 3000, -2049, 5
    pack, index = pack_from_quant(d, coef)
    assert list(pack[:2]) == [3000 & 0xfff, (0 << 12) | (-2049 & 0xfff)]
    assert pack[-1] == ((62 - 1 - 3 * 16) << 12 | 5) and list(pack[2:5]) == [0xf000] * 3   # ZRLs, no EOB
    back = port.unpack_image(port.geometry(8, 8, (1,), (1,)), pack, index)
    assert back[0] == 3000 - 4096 and back[1] == 2047 and back[63] == 5


def test_pack_batch_streams_layout():
    descs = [J.ImageDesc(40, 24, *J.SUBSAMPLINGS["420"]), J.ImageDesc(16, 16, *J.SUBSAMPLINGS["gray"], tq=(0,))]
    coef_len, _, _ = J.pack_batch(descs)
    coef = synth.batch_coefficients(descs, coef_len, synth.quality_tables(85), kinds=["natural"])
    pack, off, index = pack_batch_streams(descs, coef)
    assert off[0] == 0 and off[-1] == pack.size and index.size == coef_len // 64
    for i, d in enumerate(descs):
        p, ix = pack_from_quant(d, coef[d.coef_off:])
        assert np.array_equal(pack[off[i]:off[i + 1]], p)
        assert np.array_equal(index[d.coef_off // 64:d.coef_off // 64 + ix.size], ix)
