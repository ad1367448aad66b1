"""GPU parity: the CUDA path against the CPU oracle, bit for bit.

Y/Cb/Cr planes must equal what the reference's xjpeg + glj_real_idct8x8 produce
(src/xjpeg.c:565-584, src/dct.c:100-121); RGB must equal the colour oracle
(oracle/oracle_pipeline.c, from res/yuv.fs.glsl:11-23).  Tolerance: none.
"""
import numpy as np
import pytest

import jpeg_gpu_b200 as J
from jpeg_gpu_b200 import synth
from util import compare_batch, gpu_batch, make_batch, oracle_batch

pytestmark = pytest.mark.gpu

SIZES = [(8, 8), (16, 16), (70, 50), (64, 48), (129, 65), (512, 512), (1000, 563)]
KINDS = ["natural", "dense", "dc", "zero", "impulse"]


# (force_generic, want_yuv): RGB-only plans take the fused kernel (gray/444/422/420/440/411),
# plans that also want the planes (or 410, or force_generic) take the two-kernel generic path
PATHS = [(False, False), (False, True), (True, False), (True, True)]
PATH_IDS = ["fused-rgb", "auto-rgb+yuv", "generic-rgb", "generic-rgb+yuv"]


@pytest.mark.parametrize("force_generic,want_yuv", PATHS, ids=PATH_IDS)
@pytest.mark.parametrize("ss", ["gray", "444", "422", "420", "440", "411", "410"])
def test_parity_by_subsampling(gpu_ctx, checker, ss, force_generic, want_yuv):
    shapes = [(w, h, ss) for (w, h) in SIZES]
    descs, coef_len, rgb_len, yuv_len = make_batch(shapes, want_yuv=want_yuv)
    q = synth.quality_tables(85)
    coef = synth.batch_coefficients(descs, coef_len, q, kinds=KINDS)
    exp_rgb, exp_yuv = oracle_batch(checker, descs, coef, q, rgb_len, yuv_len)
    got_rgb, got_yuv = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, yuv_len, force_generic)
    compare_batch(descs, got_rgb, got_yuv, exp_rgb, exp_yuv)


@pytest.mark.parametrize("force_generic,want_yuv", PATHS, ids=PATH_IDS)
def test_parity_mixed_batch_two_table_sets(gpu_ctx, checker, force_generic, want_yuv):
    shapes = [(512, 512, "gray"), (1920, 1080, "420"), (70, 50, "444"), (640, 360, "422"), (33, 17, "420"),
              (256, 256, "411"), (48, 80, "440"), (1000, 563, "420"), (2048, 16, "420"), (16, 2048, "422"),
              (1600, 1200, "444"), (1537, 9, "gray"), (3840, 2160, "420")]
    descs, coef_len, rgb_len, yuv_len = make_batch(shapes, want_yuv=want_yuv, n_sets=2)
    q = np.stack([synth.quality_tables(85), synth.quality_tables(40)])
    coef = synth.batch_coefficients(descs, coef_len, q, kinds=KINDS)
    exp_rgb, exp_yuv = oracle_batch(checker, descs, coef, q, rgb_len, yuv_len)
    got_rgb, got_yuv = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, yuv_len, force_generic)
    compare_batch(descs, got_rgb, got_yuv, exp_rgb, exp_yuv)


def test_fused_many_tiles_per_cta(gpu_ctx, checker):
    """More tiles than resident CTAs, so the persistent loop and the smem ring wrap."""
    shapes = [(3840, 2160, "420")] * 3 + [(3840, 2160, "422")] * 2 + [(1920, 1080, "444")] * 2 + [(4096, 1024, "gray")]
    descs, coef_len, rgb_len, _ = make_batch(shapes, want_yuv=False)
    q = synth.quality_tables(85)
    coef = synth.batch_coefficients(descs, coef_len, q, kinds=["natural", "dense"])
    exp_rgb, _ = oracle_batch(checker, descs, coef, q, rgb_len, 0, nthreads=16)
    got_rgb, _ = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, 0)
    compare_batch(descs, got_rgb, None, exp_rgb, None)


@pytest.mark.parametrize("force_generic", [True, False])
def test_rgb_only_and_yuv_only(gpu_ctx, checker, force_generic):
    shapes = [(320, 240, "420"), (100, 100, "444"), (64, 64, "gray")]
    q = synth.quality_tables(85)
    descs, coef_len, rgb_len, yuv_len = make_batch(shapes, want_yuv=True)
    coef = synth.batch_coefficients(descs, coef_len, q)
    exp_rgb, exp_yuv = oracle_batch(checker, descs, coef, q, rgb_len, yuv_len)
    got_rgb, got_yuv = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, yuv_len, force_generic, want_rgb=False)
    assert got_rgb is None
    compare_batch(descs, None, got_yuv, exp_rgb, exp_yuv)
    descs2, coef_len2, rgb_len2, _ = make_batch(shapes, want_yuv=False)
    got_rgb, got_yuv = gpu_batch(gpu_ctx, descs2, coef, q, rgb_len2, 0, force_generic)
    assert got_yuv is None
    compare_batch(descs2, got_rgb, None, exp_rgb, None)


def test_dequantisation_wraps_like_a_short(gpu_ctx, checker):
    """src/xjpeg.c:501-503,524-527 store the int product into a short.  Inputs
    whose IDCT output stays inside int16 after the wrap must still match."""
    shapes = [(64, 64, "420"), (40, 24, "444")]
    descs, coef_len, rgb_len, yuv_len = make_batch(shapes)
    descs_rgb, _, rgb_len2, _ = make_batch(shapes, want_yuv=False)
    q = synth.quality_tables(85)
    rng = np.random.default_rng(7)
    coef = synth.batch_coefficients(descs, coef_len, q, kinds=["zero"])
    # sparse large coefficients: products exceed 32767 and wrap, outputs stay bounded
    idx = rng.choice(coef_len, size=coef_len // 40, replace=False)
    coef[idx] = rng.integers(-32768, 32768, size=idx.size).astype(np.int16)
    exp_rgb, exp_yuv = oracle_batch(checker, descs, coef, q, rgb_len, yuv_len)
    for fg in (True, False):
        got_rgb, got_yuv = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, yuv_len, fg)
        compare_batch(descs, got_rgb, got_yuv, exp_rgb, exp_yuv)
    got_rgb, _ = gpu_batch(gpu_ctx, descs_rgb, coef, q, rgb_len2, 0)   # fused kernel
    assert rgb_len2 == rgb_len
    compare_batch(descs_rgb, got_rgb, None, exp_rgb, None)


@pytest.mark.parametrize("force_generic", [True, False])
def test_sixteen_bit_quant_tables(gpu_ctx, checker, force_generic):
    """xjpeg accepts Pq=1 DQT segments (src/xjpeg.c:235-241): entries above 255."""
    shapes = [(96, 64, "420"), (64, 64, "444"), (80, 40, "422"), (64, 32, "gray")]
    descs, coef_len, rgb_len, _ = make_batch(shapes, want_yuv=False)
    rng = np.random.default_rng(11)
    q = synth.quality_tables(50).astype(np.uint16)
    q[0, 5] = 300; q[0, 63] = 1000; q[1, 1] = 256; q[1, 40] = 65535; q[1, 41] = 511
    coef = rng.integers(-40, 41, size=coef_len).astype(np.int16)
    coef[rng.random(coef_len) < 0.7] = 0
    exp_rgb, _ = oracle_batch(checker, descs, coef, q, rgb_len, 0)
    got_rgb, _ = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, 0, force_generic)
    compare_batch(descs, got_rgb, None, exp_rgb, None)


def test_host_batch_api(gpu_ctx, checker):
    shapes = [(1920, 1080, "420"), (512, 512, "gray"), (70, 50, "422")]
    descs, coef_len, rgb_len, yuv_len = make_batch(shapes)
    q = synth.quality_tables(85)
    coef = synth.batch_coefficients(descs, coef_len, q)
    exp_rgb, exp_yuv = oracle_batch(checker, descs, coef, q, rgb_len, yuv_len)
    rgb = np.zeros(rgb_len, dtype=np.uint8)
    yuv = np.zeros(yuv_len, dtype=np.uint8)
    gpu_ctx.decode_batch_host(descs, coef, q, rgb, yuv)       # pageable buffers
    compare_batch(descs, rgb, yuv, exp_rgb, exp_yuv)
    import torch
    p_coef = torch.from_numpy(coef).pin_memory()
    p_rgb = torch.zeros(rgb_len, dtype=torch.uint8).pin_memory()
    gpu_ctx.decode_batch_host(descs, p_coef, q, p_rgb, None)  # pinned buffers
    compare_batch(descs, p_rgb.numpy(), None, exp_rgb, None)


def test_grey_one_block_per_thread_variant(checker, monkeypatch):
    """JGPU_GRAY_TPB=1 (experiment, profiles/r1_ab_notes.md): k_gray_tpb keeps one block per thread
    (rows r / r+4, then columns c / c+4 in the packed lanes).  Same bits as the oracle, including
    cropped and unaligned images, 16-bit tables and the adversarial coefficient kinds."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    monkeypatch.setenv("JGPU_GRAY_TPB", "1")
    ctx = J.Context(0)           # the fused kernels are configured per context creation
    try:
        shapes = [(w, h, "gray") for (w, h) in SIZES] + [(4096, 1024, "gray"), (1537, 9, "gray"), (1920, 1080, "gray")]
        descs, coef_len, rgb_len, _ = make_batch(shapes, want_yuv=False)
        for q in (synth.quality_tables(85), synth.quality_tables(85) * 3):   # 8-bit and 16-bit entries
            q = q.astype(np.uint16)
            coef = synth.batch_coefficients(descs, coef_len, q, kinds=KINDS)
            exp_rgb, _ = oracle_batch(checker, descs, coef, q, rgb_len, 0, nthreads=8)
            got_rgb, _ = gpu_batch(ctx, descs, coef, q, rgb_len, 0)
            compare_batch(descs, got_rgb, None, exp_rgb, None)
    finally:
        ctx.close()
        monkeypatch.delenv("JGPU_GRAY_TPB")
        J.Context(0).close()     # back to the product configuration for the tests that follow


def test_sixteen_bit_tables_through_the_chunked_host_path(gpu_ctx, checker):
    """16-bit DQT entries AND a batch of several 48 MB chunks: jgpu_decode_batch_host runs the chunks
    of one cached plan on three streams.  The table preparation of a later chunk used to zero the
    16-bit flag under an earlier chunk's kernels (both instantiations then returned at once and the
    chunk's pixels were never written); it now only ever WRITES the flag's value."""
    import torch
    shapes = [(1920, 1080, "420")] * 40          # 40 x 6.3 MB of coefficients = 5 chunks
    descs, coef_len, rgb_len, _ = make_batch(shapes, want_yuv=False)
    q = synth.quality_tables(60).astype(np.uint16)
    q[0, 9] = 300; q[1, 2] = 700; q[1, 63] = 4000
    one = synth.image_coefficients(descs[0], synth.quality_tables(60), synth.SEED_BASE)
    lay = descs[0].query_layout()
    coef = torch.zeros(coef_len, dtype=torch.int16).pin_memory()
    for d in descs:
        coef[d.coef_off:d.coef_off + lay.coef_len] = torch.from_numpy(one)
    one_desc, c_len, r_len, _ = make_batch(shapes[:1], want_yuv=False)
    exp, _ = oracle_batch(checker, one_desc, one, q, r_len, 0, nthreads=8)
    for rep in range(3):
        rgb = torch.full((rgb_len,), 0x5A, dtype=torch.uint8).pin_memory()
        gpu_ctx.decode_batch_host(descs, coef, q, rgb, None)
        got = rgb.numpy()
        for i, d in enumerate(descs):
            assert np.array_equal(got[d.rgb_off:d.rgb_off + lay.rgb_len], exp[:lay.rgb_len]), (rep, i)


@pytest.mark.parametrize("ss", ["gray", "444", "422", "420", "440", "411"])
def test_planes_out_of_the_fused_kernel(gpu_ctx, checker, ss):
    """JGPU_OUT_YUV plans run the fused kernel too: the padded planes of xjpeg's YUV output
    (src/xjpeg.c:565-584), every byte of them, for sizes that leave half-tasks, half units and
    invisible MCU rows at the edges."""
    shapes = [(1920, 1080, ss), (520, 40, ss), (70, 50, ss), (8, 8, ss), (1000, 563, ss), (264, 16, ss), (2100, 24, ss)]
    q = synth.quality_tables(85)
    descs, coef_len, rgb_len, yuv_len = make_batch(shapes, want_yuv=True)
    coef = synth.batch_coefficients(descs, coef_len, q, kinds=["natural", "dense", "int16"])
    _, exp_yuv = oracle_batch(checker, descs, coef, q, rgb_len, yuv_len, nthreads=8)
    plan = gpu_ctx.plan(descs, rgb=False, yuv=True)
    assert plan.launches <= 3, "planes-only plans must take the fused path (prep + two instantiations per mode)"
    plan.close()
    _, got_yuv = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, yuv_len, want_rgb=False)
    compare_batch(descs, None, got_yuv, None, exp_yuv)
    # bytes between the images' planes stay untouched (0xCD fill of gpu_batch)
    starts = {d.yuv_off for d in descs}
    for d in descs:
        end = d.yuv_off + d.query_layout().data_len
        if end < yuv_len and end not in starts:
            assert got_yuv[end] == 0xCD


@pytest.mark.parametrize("ss", ["gray", "444", "422", "420", "440", "411", "410"])
def test_rows_of_every_alignment(gpu_ctx, checker, ss):
    """Widths 257..272 (and a few narrow ones): tight pixel rows of 3 x width bytes start at every 16-byte
    phase, so every case of the any-alignment store path (head bytes, head words, two 128-bit words, tail)
    runs, next to units cropped by the right edge and tiles cropped by the bottom edge; the bytes between
    the images must stay untouched."""
    shapes = [(256 + r, 17 + r % 9, ss) for r in range(1, 17)] + [(33 + r, 9, ss) for r in range(0, 16, 5)]
    descs, coef_len, rgb_len, _ = make_batch(shapes, want_yuv=False)
    q = synth.quality_tables(85)
    coef = synth.batch_coefficients(descs, coef_len, q, kinds=KINDS)
    exp_rgb, _ = oracle_batch(checker, descs, coef, q, rgb_len, 0, nthreads=8)
    got_rgb, _ = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, 0)
    compare_batch(descs, got_rgb, None, exp_rgb, None)
    starts = {d.rgb_off for d in descs}
    for d in descs:
        end = d.rgb_off + d.query_layout().rgb_len
        if end < rgb_len and end not in starts:
            assert (got_rgb[end:min(rgb_len, (end + 255) // 256 * 256)] == 0xAB).all(), (d.width, d.height)


@pytest.mark.parametrize("n_images", [1, 2, 7, 9, 150])
def test_task_counts_around_the_static_share(gpu_ctx, checker, n_images):
    """k_tk hands every warp pair its first four tasks statically and the rest from a counter: batches with
    fewer tasks than pairs, with a handful, and with a few more than the static share of one SM."""
    shapes = [(520, 40, "420"), (300, 24, "422"), (260, 8, "444"), (70, 50, "gray"), (530, 20, "411"), (100, 33, "410")]
    shapes = [shapes[i % len(shapes)] for i in range(n_images)]
    descs, coef_len, rgb_len, _ = make_batch(shapes, want_yuv=False)
    q = synth.quality_tables(85)
    coef = synth.batch_coefficients(descs, coef_len, q, kinds=["natural", "dense"])
    exp_rgb, _ = oracle_batch(checker, descs, coef, q, rgb_len, 0, nthreads=8)
    for _ in range(2):   # twice: the second run reuses the plan's task counters
        got_rgb, _ = gpu_batch(gpu_ctx, descs, coef, q, rgb_len, 0)
        compare_batch(descs, got_rgb, None, exp_rgb, None)


@pytest.mark.parametrize("kernel", ["mcu", "tk"])
def test_either_kernel_everywhere(checker, monkeypatch, kernel):
    """By default grey pixels and the planes of grey / 4:2:0 / 4:2:2 run k_mcu (one warp doing both halves of
    k_tk's pairs; faster where the colour warps have nothing to do, profiles/r2_notes.md) and everything else
    k_tk; JGPU_KERNEL forces one of them for every mode it has: same bits, pixels and planes."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    monkeypatch.setenv("JGPU_KERNEL", kernel)
    ctx = J.Context(0)           # the kernel choice is read per context creation
    try:
        shapes = [(w, h, ss) for ss in ("gray", "444", "422", "420", "440") for (w, h) in [(70, 50), (1000, 563), (1920, 1080)]]
        descs, coef_len, rgb_len, yuv_len = make_batch(shapes, want_yuv=False)
        q = synth.quality_tables(85)
        coef = synth.batch_coefficients(descs, coef_len, q, kinds=KINDS)
        exp_rgb, _ = oracle_batch(checker, descs, coef, q, rgb_len, 0, nthreads=8)
        got_rgb, _ = gpu_batch(ctx, descs, coef, q, rgb_len, 0)
        compare_batch(descs, got_rgb, None, exp_rgb, None)
        descs, coef_len, rgb_len, yuv_len = make_batch(shapes, want_yuv=True)
        _, exp_yuv = oracle_batch(checker, descs, coef, q, rgb_len, yuv_len, nthreads=8)
        plan = ctx.plan(descs, rgb=False, yuv=True)
        assert plan.launches <= 1 + 2 * 5, "planes-only plans take the fused path"
        plan.close()
        _, got_yuv = gpu_batch(ctx, descs, coef, q, rgb_len, yuv_len, want_rgb=False)
        compare_batch(descs, None, got_yuv, None, exp_yuv)
    finally:
        ctx.close()
        monkeypatch.delenv("JGPU_KERNEL")
        J.Context(0).close()     # back to the product's choice for the tests that follow
