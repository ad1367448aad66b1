"""CUDA_DECODE_CTX_VTBL, the third decoder backend, driven through the reference's own
call protocol (src/jpeg_gpu.c:612-704,1231-1237) on real JPEG files, against what the
reference's xjpeg produced for them (tests/golden/): YUV planes bit-exact, RGB exact vs
the colour oracle; plus size-independent properties at the BASELINE sizes."""
import numpy as np
import pytest

import jpeg_gpu_b200 as J
from golden_util import NAMES, load
from jpeg_gpu_b200 import synth
from util import compare_batch, gpu_batch, make_batch, oracle_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def need_gpu(gpu_ctx):
    return gpu_ctx


@pytest.mark.parametrize("name", NAMES)
def test_cuda_backend_yuv_and_rgb(name):
    jpg, z, g = load(name)
    with J.Decoder(jpg, impl="cuda") as dec:
        h = dec.decode_header()
        assert (h.width, h.height) == (int(z["hdr_width"]), int(z["hdr_height"]))
        planes = dec.decode_image("yuv")["planes"]
        assert np.array_equal(np.concatenate([p.ravel() for p in planes]), z["yuv"])
        # steady state: reset -> header -> image, device buffers and plan are reused
        for _ in range(2):
            dec.decode_reset()
            dec.decode_header()
            px = dec.decode_image("rgb")["pixels"]
            assert np.array_equal(px.reshape(-1), z["rgb"])
        dec.decode_reset()
        dec.decode_header()
        assert np.array_equal(dec.decode_image("quant")["coef"], z["quant"])   # forwarded to the front end


@pytest.mark.parametrize("name", NAMES)
def test_cuda_backend_with_pack_upload(name):
    """cuda_decode_set_upload(JPEG_DECODE_PACK): the reference's `-o pack` choice -- the run/level
    stream crosses to the device and is expanded there; same planes, same pixels."""
    from jpeg_gpu_b200 import _capi
    jpg, z, g = load(name)
    assert _capi.lib().cuda_decode_set_upload(_capi.JPEG_DECODE_PACK) == 0
    try:
        with J.Decoder(jpg, impl="cuda") as dec:
            dec.decode_header()
            planes = dec.decode_image("yuv")["planes"]
            assert np.array_equal(np.concatenate([p.ravel() for p in planes]), z["yuv"])
            dec.decode_reset()
            dec.decode_header()
            assert np.array_equal(dec.decode_image("rgb")["pixels"].reshape(-1), z["rgb"])
    finally:
        assert _capi.lib().cuda_decode_set_upload(_capi.JPEG_DECODE_QUANT) == 0
    assert _capi.lib().cuda_decode_set_upload(_capi.JPEG_DECODE_RGB) == 1   # not an upload format


@pytest.mark.parametrize("name", NAMES)
def test_cuda_backend_with_entropy_decoding_on_the_device(name):
    """cuda_decode_set_entropy(1): decode_image(RGB) uploads the file as it is and decodes the
    Huffman scan on the GPU; same pixels and the same planes (bit-exact with xjpeg's YUV output),
    same protocol."""
    from jpeg_gpu_b200 import _capi
    jpg, z, g = load(name)
    assert _capi.lib().cuda_decode_set_entropy(1) == 0
    try:
        with J.Decoder(jpg, impl="cuda") as dec:
            for _ in range(2):
                dec.decode_header()
                assert np.array_equal(dec.decode_image("rgb")["pixels"].reshape(-1), z["rgb"])
                dec.decode_reset()
            dec.decode_header()
            planes = dec.decode_image("yuv")["planes"]     # also decoded on the device, planes read back
            assert np.array_equal(np.concatenate([p.ravel() for p in planes]), z["yuv"])
    finally:
        assert _capi.lib().cuda_decode_set_entropy(0) == 0
    assert _capi.lib().cuda_decode_set_entropy(7) == 1


def test_cuda_backend_error_convention():
    jpg, _, _ = load("c420_64x48")
    with J.Decoder(jpg, impl="cuda") as dec:
        with pytest.raises(J.DecodeError):      # decode_image before decode_header
            dec.decode_image("rgb")
    with J.Decoder(jpg[:200], impl="cuda") as dec:
        with pytest.raises(J.DecodeError):
            dec.decode_header()
            dec.decode_image("rgb")


def test_config2_1080p_420_batch1(gpu_ctx, checker):
    """BASELINE config 2: 1920x1080 4:2:0, batch 1: planes bit-exact, RGB exact."""
    shapes = [(1920, 1080, "420")]
    q = synth.quality_tables(85)
    descs, coef_len, rgb_len, yuv_len = make_batch(shapes, want_yuv=True)
    coef = synth.batch_coefficients(descs, coef_len, q)
    exp_rgb, exp_yuv = oracle_batch(checker, descs, coef, q, rgb_len, yuv_len, nthreads=8)
    got_rgb, got_yuv = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, yuv_len)            # generic (planes wanted)
    compare_batch(descs, got_rgb, got_yuv, exp_rgb, exp_yuv)
    descs2, _, rgb_len2, _ = make_batch(shapes, want_yuv=False)
    got_rgb, _ = gpu_batch(gpu_ctx, descs2, coef, q, rgb_len2, 0)                       # fused
    compare_batch(descs2, got_rgb, None, exp_rgb, None)


def test_full_size_batch_properties(gpu_ctx, checker):
    """At BASELINE's batch sizes the oracle is too slow to check every pixel, so: (1) the
    fused and the generic path agree on every byte of a 24-image 4K 4:2:0 batch; (2) three
    images of it are checked against the oracle; (3) identical inputs at different batch
    positions give identical outputs (no cross-image state)."""
    import torch
    n = 24
    shapes = [(3840, 2160, "420")] * n
    q = synth.quality_tables(85)
    descs, coef_len, rgb_len, _ = make_batch(shapes, want_yuv=False)
    dev = torch.device("cuda", gpu_ctx.device)
    d_coef = synth.torch_batch_coefficients(descs, coef_len, q, dev)
    lay = descs[0].query_layout()
    # image n-1 := image 0, so (3) has something to compare
    d_coef[descs[-1].coef_off:descs[-1].coef_off + lay.coef_len] = d_coef[:lay.coef_len]
    d_q = torch.from_numpy(q.astype(np.int16).reshape(-1)).to(dev)
    outs = []
    for generic in (False, True):
        d_rgb = torch.zeros(rgb_len, dtype=torch.uint8, device=dev)
        plan = gpu_ctx.plan(descs, rgb=True, force_generic=generic)
        plan.run(d_coef, d_q, d_rgb)
        torch.cuda.synchronize()
        plan.close()
        outs.append(d_rgb)
    assert torch.equal(outs[0], outs[1])
    first = outs[0][:lay.rgb_len]
    last = outs[0][descs[-1].rgb_off:descs[-1].rgb_off + lay.rgb_len]
    assert torch.equal(first, last)
    coef = d_coef.cpu().numpy()
    for i in (0, 11, n - 1):
        one, c_len, r_len, _ = make_batch(shapes[:1], want_yuv=False)
        exp, _ = oracle_batch(checker, one, coef[descs[i].coef_off:descs[i].coef_off + c_len], q, r_len, 0, nthreads=16)
        got = outs[0][descs[i].rgb_off:descs[i].rgb_off + lay.rgb_len].cpu().numpy()
        assert np.array_equal(got, exp[:lay.rgb_len]), i


def test_config4_shape_422(gpu_ctx, checker):
    """BASELINE config 4's image shape (3840x2160 4:2:2), a few images: fused vs oracle."""
    shapes = [(3840, 2160, "422")] * 2
    q = synth.quality_tables(85)
    descs, coef_len, rgb_len, _ = make_batch(shapes, want_yuv=False)
    coef = synth.batch_coefficients(descs, coef_len, q, kinds=["natural", "dense"])
    exp, _ = oracle_batch(checker, descs, coef, q, rgb_len, 0, nthreads=16)
    got, _ = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, 0)
    compare_batch(descs, got, None, exp, None)


def test_unaligned_rgb_offsets_take_the_slow_store_path(gpu_ctx, checker):
    shapes = [(64, 48, "420"), (48, 32, "444"), (80, 16, "422"), (32, 32, "gray")]
    q = synth.quality_tables(85)
    descs, coef_len, rgb_len, _ = make_batch(shapes, want_yuv=False)
    for i, d in enumerate(descs):       # knock every image off 16-byte alignment
        d.rgb_off += 1 + i
    rgb_len += 16
    coef = synth.batch_coefficients(descs, coef_len, q)
    exp, _ = oracle_batch(checker, descs, coef, q, rgb_len, 0)
    got, _ = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, 0)
    compare_batch(descs, got, None, exp, None)
    # bytes between images must be untouched (0xAB fill of gpu_batch)
    assert got[descs[0].rgb_off - 1] == 0xAB
