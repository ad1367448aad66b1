"""GPU: PACK input.  k_unpack (jgpu_unpack.cu) against the reference's own PACK output for the
golden files and against the consumer restatement (oracle/oracle_pack.c,
res/horz_pack_yuv.fs.glsl:105-127); the packed host entry point against the oracle's RGB.
Bit-exact."""
import numpy as np
import pytest

import jpeg_gpu_b200 as J
from jpeg_gpu_b200 import synth
from jpeg_gpu_b200.batch import pack_batch_streams
from golden_util import NAMES, load
from util import compare_batch, make_batch, oracle_batch

pytestmark = pytest.mark.gpu


def _unpack_on_gpu(ctx, descs, coef_len, pack, pack_off, index, fill=0x5a5a):
    import torch
    dev = torch.device("cuda", ctx.device)
    d_pack = torch.from_numpy(pack.view(np.int16)).to(dev)
    d_off = torch.from_numpy(pack_off).to(dev)
    d_index = torch.from_numpy(index).to(dev)
    d_coef = torch.full((coef_len,), fill, dtype=torch.int16, device=dev)
    plan = ctx.plan(descs, rgb=True)
    plan.unpack(d_pack, d_off, d_index, d_coef)
    torch.cuda.synchronize(dev)
    plan.close()
    return d_coef.cpu().numpy()


@pytest.mark.parametrize("name", NAMES)
def test_unpack_reference_stream_gives_reference_planes(gpu_ctx, name):
    _, z, g = load(name)
    n = int(z["hdr_ncomps"])
    d = J.ImageDesc(int(z["hdr_width"]), int(z["hdr_height"]), [int(v) for v in z["hdr_hsamp"][:n]],
                    [int(v) for v in z["hdr_vsamp"][:n]], tq=[int(v) for v in z["hdr_tq"][:n]])
    coef_len, _, _ = J.pack_batch([d])
    got = _unpack_on_gpu(gpu_ctx, [d], coef_len, z["pack"], np.array([0, z["pack"].size], dtype=np.int64), z["index"])
    for p in g.planes:   # padding blocks are not coded and must be left alone
        nn = 64 * p.hblocks * p.vblocks
        assert np.array_equal(got[p.coef_off:p.coef_off + nn], z["quant"][p.coef_off:p.coef_off + nn])
        pad = got[p.coef_off + nn:p.coef_off + 64 * (p.hblocks << p.xdec) * p.cstride]
        assert np.all(pad == 0x5a5a)


def test_unpack_mixed_batch_matches_oracle_consumer(gpu_ctx, port):
    shapes = [(512, 512, "gray"), (1920, 1080, "420"), (70, 50, "444"), (640, 360, "422"), (33, 17, "420"),
              (256, 256, "411"), (48, 80, "440"), (1000, 563, "420"), (2048, 16, "420"), (1537, 9, "gray")]
    descs, coef_len, _, _ = make_batch(shapes, want_yuv=False)
    q = synth.quality_tables(85)
    coef = synth.batch_coefficients(descs, coef_len, q, kinds=["natural", "dense", "dc", "zero", "impulse"])
    pack, off, index = pack_batch_streams(descs, coef)
    got = _unpack_on_gpu(gpu_ctx, descs, coef_len, pack, off, index)
    for i, d in enumerate(descs):
        g = port.geometry(d.width, d.height, d.hsamp, d.vsamp)
        lay = d.query_layout()
        want = port.unpack_image(g, pack[off[i]:off[i + 1]], index[d.coef_off // 64:d.coef_off // 64 + lay.coef_len // 64])
        for p in g.planes:
            nn = 64 * p.hblocks * p.vblocks
            a = got[d.coef_off + p.coef_off:d.coef_off + p.coef_off + nn]
            assert np.array_equal(a, want[p.coef_off:p.coef_off + nn]), (i, shapes[i])
            assert np.array_equal(a, coef[d.coef_off + p.coef_off:d.coef_off + p.coef_off + nn]), (i, shapes[i])


def test_unpack_never_reads_past_the_stream(gpu_ctx):
    """Cut-off streams, wild index entries and overlong runs truncate the block; nothing faults."""
    d = J.ImageDesc(64, 64, (1,), (1,), tq=(0,))
    coef_len, _, _ = J.pack_batch([d])
    q = synth.quality_tables(85)
    coef = synth.batch_coefficients([d], coef_len, q, kinds=["dense"])
    pack, off, index = pack_batch_streams([d], coef)
    index = index.copy()
    index[3] = 2 ** 30        # far outside
    index[4] = -7             # negative
    cut = np.array([0, pack.size // 2], dtype=np.int64)
    bad = pack.copy()
    bad[index[0] + 1] = 0xf001  # run 15 ... repeated below so that j overruns 63
    bad[index[0] + 2:index[0] + 8] = 0xf001
    got = _unpack_on_gpu(gpu_ctx, [d], coef_len, bad, cut, index, fill=0)
    assert got.shape == (coef_len,)
    assert np.all(got[3 * 64:5 * 64] == 0)
    blk0 = got[:64]
    assert blk0[0] == coef[0] and np.count_nonzero(blk0[1:]) == 3   # 16, 32, 48 then overrun


@pytest.mark.parametrize("want_yuv", [False, True], ids=["fused-rgb", "generic-rgb+yuv"])
def test_host_packed_entry_matches_oracle(gpu_ctx, checker, want_yuv):
    shapes = [(1920, 1080, "420"), (70, 50, "444"), (640, 360, "422"), (512, 512, "gray"), (1000, 563, "420"),
              (3840, 2160, "420"), (48, 80, "440")] + [(1280, 720, "420")] * 9
    descs, coef_len, rgb_len, yuv_len = make_batch(shapes, want_yuv=want_yuv, n_sets=2)
    q = np.stack([synth.quality_tables(85), synth.quality_tables(40)])
    coef = synth.batch_coefficients(descs, coef_len, q, kinds=["natural", "dense", "impulse"])
    exp_rgb, exp_yuv = oracle_batch(checker, descs, coef, q, rgb_len, yuv_len, nthreads=8)
    pack, off, index = pack_batch_streams(descs, coef)
    rgb = np.zeros(rgb_len, dtype=np.uint8)
    yuv = np.zeros(yuv_len, dtype=np.uint8) if want_yuv else None
    gpu_ctx.decode_batch_host_packed(descs, pack, off, index, q, rgb, yuv)
    compare_batch(descs, rgb, yuv, exp_rgb, exp_yuv)
    # and again with the cached plan (steady state)
    rgb[:] = 0
    gpu_ctx.decode_batch_host_packed(descs, pack, off, index, q, rgb, yuv)
    compare_batch(descs, rgb, yuv, exp_rgb, exp_yuv)
