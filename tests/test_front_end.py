"""The CPU entropy front end (JFRONT_DECODE_CTX_VTBL) through the reference's five-slot
protocol, against what the reference's xjpeg produced for the same files (tests/golden/)."""
import numpy as np
import pytest

import jpeg_gpu_b200 as J
from golden_util import NAMES, load


@pytest.mark.parametrize("name", NAMES)
def test_header_quant_dct_match_xjpeg(name):
    jpg, z, g = load(name)
    with J.Decoder(jpg, impl="jfront") as dec:
        h = dec.decode_header()
        assert (h.width, h.height, h.bits, h.ncomps, h.restart_interval) == \
               (int(z["hdr_width"]), int(z["hdr_height"]), int(z["hdr_bits"]), int(z["hdr_ncomps"]),
                int(z["hdr_restart_interval"]))
        assert h.hsamp == list(z["hdr_hsamp"]) and h.vsamp == list(z["hdr_vsamp"]) and h.tq == list(z["hdr_tq"])
        assert h.hblocks == [p.hblocks for p in g.planes] and h.vblocks == [p.vblocks for p in g.planes]
        assert np.array_equal(h.qtabs, z["hdr_qtabs"]) and h.qvalid == list(z["hdr_qvalid"])
        quant = dec.decode_image("quant")["coef"]
        assert np.array_equal(quant, z["quant"])
        # steady-state protocol: reset -> header -> image (src/jpeg_gpu.c:1231-1237)
        dec.decode_reset()
        dec.decode_header()
        assert np.array_equal(dec.decode_image("dct")["coef"], z["dct"])


def test_subsampling_names():
    want = {"gray_48x40": "Mono", "c444_40x24": "4:4:4", "c422_56x24": "4:2:2", "c420_64x48": "4:2:0"}
    for name, s in want.items():
        jpg, _, _ = load(name)
        with J.Decoder(jpg, impl="jfront") as dec:
            assert dec.decode_header().subsamp == s


@pytest.mark.parametrize("name", ["c420_64x48", "gray_48x40", "c422_rst_33x17"])
def test_pack_stream_expands_to_quant(name):
    """PACK (src/xjpeg.c:484-496,513-519,531-535): DC word = dc & 0xfff, AC word =
    run<<12 | value & 0xfff, EOB word 0; index[] = first word of each block."""
    jpg, z, g = load(name)
    zz = J.synth.NATURAL
    with J.Decoder(jpg, impl="jfront") as dec:
        dec.decode_header()
        r = dec.decode_image("pack")
    pack, index = r["pack"].astype(np.int64) & 0xffff, r["index"]
    ioff = 0
    for p in g.planes:
        nblk = p.hblocks * p.vblocks
        want = z["quant"][p.coef_off:p.coef_off + 64 * nblk].reshape(nblk, 64)
        for b in range(nblk):
            i = int(index[ioff + b])
            blk = np.zeros(64, dtype=np.int64)
            v = pack[i] & 0xfff
            blk[0] = v - 4096 if v & 0x800 else v
            k, i = 0, i + 1
            while k < 63:
                w = pack[i]; i += 1
                if w == 0:
                    break
                k += (w >> 12) + 1
                v = w & 0xfff
                blk[zz[k]] = v - 4096 if v & 0x800 else v
            assert np.array_equal(blk, want[b]), (name, b)
        ioff += (p.hblocks << p.xdec) * p.cstride
    assert sum(r["packed"]) == pack.size


def test_yuv_and_rgb_are_not_the_front_ends_job():
    """Like the xjpeg backend for RGB (src/jpeg_wrap.c:335-339): EXIT_FAILURE, no crash."""
    jpg, _, _ = load("c420_64x48")
    with J.Decoder(jpg, impl="jfront") as dec:
        dec.decode_header()
        for out in ("yuv", "rgb"):
            with pytest.raises(J.DecodeError):
                dec.decode_image(out)


@pytest.mark.parametrize("mutate", ["truncate", "not_jpeg", "progressive", "garbage_scan"])
def test_malformed_input_fails_cleanly(mutate):
    jpg, _, _ = load("c420_64x48")
    if mutate == "truncate":
        data = jpg[:len(jpg) // 3]
    elif mutate == "not_jpeg":
        data = b"GIF89a" + jpg[6:]
    elif mutate == "progressive":
        data = jpg.replace(b"\xff\xc0", b"\xff\xc2", 1)
    else:
        sos = jpg.index(b"\xff\xda")
        rng = np.random.default_rng(1)
        data = jpg[:sos + 14] + bytes(rng.integers(0, 255, size=len(jpg) - sos - 16, dtype=np.uint8)) + b"\xff\xd9"
    dec = J.Decoder(data, impl="jfront")
    try:
        try:
            dec.decode_header()
        except J.DecodeError:
            return
        try:
            dec.decode_image("quant")       # must terminate without crashing; either outcome is fine
        except J.DecodeError:
            pass
    finally:
        dec.close()


def test_decode_image_before_header_is_an_error():
    jpg, _, _ = load("gray_48x40")
    with J.Decoder(jpg, impl="jfront") as dec:
        with pytest.raises(J.DecodeError):
            dec.decode_image("quant")
