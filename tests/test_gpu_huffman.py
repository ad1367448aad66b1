"""GPU: Huffman decoding on the device (jgpu_huff.cu) through jgpu_decode_jpegs_ex:
pixels must equal what the sequential reader + the same back end give (and the oracle's), file by
file, including files the device decoder has to hand back to the sequential reader."""
import io

import numpy as np
import pytest

import jpeg_gpu_b200 as J
import oracle
from golden_util import NAMES, load

pytestmark = pytest.mark.gpu


def _picture(w, h, seed, noise=40):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 5 + yy * 3) % 256, (yy * 7 + xx) % 256, (xx * 2 + yy * 9) % 256], -1)
    return np.clip(base + rng.integers(-noise, noise + 1, size=base.shape), 0, 255).astype(np.uint8)


def _jpeg(w, h, ss, rst=0, q=85, optimize=False, seed=1, noise=40):
    from PIL import Image
    img = _picture(w, h, seed, noise)
    bio = io.BytesIO()
    if ss == "L":
        Image.fromarray(img[..., 0]).save(bio, "JPEG", quality=q, restart_marker_blocks=rst, optimize=optimize)
    else:
        Image.fromarray(img).save(bio, "JPEG", quality=q, subsampling=ss, restart_marker_blocks=rst, optimize=optimize)
    return bio.getvalue()


def _expected_rgb(checker, jpg):
    with J.Decoder(jpg, impl="jfront") as dec:
        h = dec.decode_header()
        quant = dec.decode_image("quant")["coef"]
    g = oracle.geometry(h.width, h.height, h.hsamp, h.vsamp)
    rgb, _ = checker.decode_image(g, quant, h.qtabs, h.tq, nthreads=8)
    return rgb


def test_files_of_every_kind_match_the_sequential_reader_and_the_oracle(gpu_ctx, checker):
    pytest.importorskip("PIL")
    files = [_jpeg(1920, 1080, 2), _jpeg(1920, 1080, 2, rst=120), _jpeg(1000, 563, 1, rst=7, q=60, optimize=True),
             _jpeg(2048, 1536, 0, q=95), _jpeg(1537, 771, 2, rst=97, q=30), _jpeg(640, 480, "L", rst=80),
             _jpeg(333, 222, "L", q=98, optimize=True), _jpeg(3840, 2160, 2, noise=24), _jpeg(3840, 2160, 2, rst=240),
             _jpeg(17, 9, 2), _jpeg(64, 64, 2, rst=1)]
    files += [load(n)[0] for n in NAMES]
    got, gi = gpu_ctx.decode_jpegs(files, entropy="gpu")
    ref, ri = gpu_ctx.decode_jpegs(files, entropy="cpu")
    for k, (a, b) in enumerate(zip(gi, ri)):
        assert a.status == 0 and b.status == 0, k
        assert (a.rgb_off, a.rgb_len) == (b.rgb_off, b.rgb_len)
        assert np.array_equal(got[a.rgb_off:a.rgb_off + a.rgb_len], ref[b.rgb_off:b.rgb_off + b.rgb_len]), k
    # decoded on the device, in parallel pieces (not through the fallback)
    assert gi[0].tasks > 1000 and gi[7].tasks > 10000 and gi[8].tasks > 10000
    for k in (0, 2, 5, 9):
        want = _expected_rgb(checker, files[k])
        assert np.array_equal(got[gi[k].rgb_off:gi[k].rgb_off + gi[k].rgb_len].reshape(gi[k].shape), want), k


def test_a_batch_larger_than_one_chunk(gpu_ctx):
    """More coefficients than one 192 MB chunk: several groups on alternating streams, same
    pixels as the sequential reader; run twice (buffers and the cached plan are reused)."""
    pytest.importorskip("PIL")
    base = [_jpeg(3840, 2160, 2, seed=s, noise=16) for s in range(3)] + [_jpeg(1280, 720, 1, rst=40, seed=9)]
    files = [base[i % len(base)] for i in range(11)]
    ref, ri = gpu_ctx.decode_jpegs(files, entropy="cpu")
    for _ in range(2):
        got, gi = gpu_ctx.decode_jpegs(files, entropy="gpu")
        assert all(i.status == 0 and i.tasks > 1 for i in gi)
        assert np.array_equal(got, ref)


def test_damaged_scans_give_what_the_sequential_reader_gives(gpu_ctx, capfd):
    """Bytes flipped inside the scan, a cut file, a file with its restart markers renumbered: the
    device decoder must flag them and the outcome (pixels or rejection) must be the sequential
    reader's."""
    pytest.importorskip("PIL")
    good = _jpeg(640, 480, 2, rst=20)
    rng = np.random.default_rng(3)
    bad = []
    for k in range(6):
        b = bytearray(good)
        for pos in rng.integers(len(b) // 2, len(b) - 4, size=1 + k):
            b[pos] = int(rng.integers(0, 255))
        bad.append(bytes(b))
    bad.append(good[:len(good) * 3 // 5])
    renum = bytearray(good)
    for i in range(len(renum) - 1):
        if renum[i] == 0xFF and 0xD0 <= renum[i + 1] <= 0xD7:
            renum[i + 1] = 0xD0 + ((renum[i + 1] - 0xD0 + 3) & 7)
    bad.append(bytes(renum))
    files = [good] + bad + [good]
    got, gi = gpu_ctx.decode_jpegs(files, strict=False, entropy="gpu")
    ref, ri = gpu_ctx.decode_jpegs(files, strict=False, entropy="cpu")
    assert gi[0].status == 0 and gi[-1].status == 0 and gi[0].tasks > 1
    for k, (a, b) in enumerate(zip(gi, ri)):
        assert a.status == b.status, (k, a.status, b.status, a.message, b.message)
        if a.status == 0:
            assert np.array_equal(got[a.rgb_off:a.rgb_off + a.rgb_len], ref[b.rgb_off:b.rgb_off + b.rgb_len]), k
    capfd.readouterr()


def test_pixels_can_stay_on_the_device(gpu_ctx):
    """JGPU_JPEGS_DEVICE_OUT: the output buffer is device memory; same bytes as the host-output call."""
    pytest.importorskip("PIL")
    import torch
    files = [_jpeg(1920, 1080, 2, rst=120), _jpeg(333, 222, "L"), load("c444_40x24")[0], _jpeg(1280, 720, 1)]
    ref, ri = gpu_ctx.decode_jpegs(files, entropy="gpu")
    total, _ = J.probe_jpegs(files)
    dev = torch.zeros(total, dtype=torch.uint8, device="cuda:0")
    out, gi = gpu_ctx.decode_jpegs(files, dev, entropy="gpu")
    got = out.cpu().numpy()
    for a, b in zip(gi, ri):
        assert a.status == 0 and (a.rgb_off, a.rgb_len) == (b.rgb_off, b.rgb_len)
        assert np.array_equal(got[a.rgb_off:a.rgb_off + a.rgb_len], ref[b.rgb_off:b.rgb_off + b.rgb_len])
    # planes on the device, and a damaged file in the batch (it goes through the sequential reader after the
    # one block-decoder launch that ends a device-output call)
    bad = bytearray(files[0])
    bad[len(bad) // 2] ^= 0x5a
    files2 = files + [bytes(bad)]
    ref, ri = gpu_ctx.decode_jpegs(files2, entropy="gpu", out="yuv", strict=False)   # host output: per-group launches
    total, _ = J.probe_jpegs(files2, out="yuv")
    dev = torch.zeros(total, dtype=torch.uint8, device="cuda:0")
    out, gi = gpu_ctx.decode_jpegs(files2, dev, entropy="gpu", out="yuv", strict=False)
    got = out.cpu().numpy()
    for a, b in zip(gi, ri):
        assert a.status == b.status and (a.rgb_off, a.rgb_len) == (b.rgb_off, b.rgb_len)
        if a.status == 0:
            assert np.array_equal(got[a.rgb_off:a.rgb_off + a.rgb_len], ref[b.rgb_off:b.rgb_off + b.rgb_len])


def test_planes_instead_of_pixels(gpu_ctx):
    """JGPU_JPEGS_OUT_YUV: per file the padded Y|Cb|Cr planes, bit-exact with what the compiled
    reference's xjpeg_decode_image(YUV) wrote for the golden files."""
    files = [load(n)[0] for n in NAMES]
    buf, infos = gpu_ctx.decode_jpegs(files, entropy="gpu", out="yuv")
    for name, inf in zip(NAMES, infos):
        z = load(name)[1]
        assert inf.status == 0 and inf.rgb_len == z["yuv"].size, name
        assert np.array_equal(buf[inf.rgb_off:inf.rgb_off + inf.rgb_len], z["yuv"]), name


def test_many_small_files_in_one_batch(gpu_ctx):
    """Hundreds of small files: one group, one upload, a grid of (pieces x files) with mostly tiny
    files next to a larger one; pixels as from the sequential reader."""
    pytest.importorskip("PIL")
    base = [load(n)[0] for n in NAMES] + [_jpeg(640, 480, 2, rst=20), _jpeg(96, 64, 0)]
    files = [base[(7 * i) % len(base)] for i in range(400)]
    got, gi = gpu_ctx.decode_jpegs(files, entropy="gpu")
    ref, ri = gpu_ctx.decode_jpegs(files, entropy="cpu")
    assert all(i.status == 0 for i in gi)
    assert np.array_equal(got, ref)


def test_fuzzed_files_agree_with_the_host_thread_path(gpu_ctx, capfd):
    """tools/fuzz_jpegs_gpu.py in small: 240 files with flipped, deleted or cut-off bytes (and
    undamaged ones): status and pixels of the device entropy path == those of the host path."""
    pytest.importorskip("PIL")
    rng = np.random.default_rng(21)
    base = [load(n)[0] for n in NAMES] + [_jpeg(320, 240, 2, rst=20, seed=3), _jpeg(333, 222, 1, rst=5, seed=4),
                                          _jpeg(200, 120, 0, seed=5)]
    files = []
    for i in range(240):
        b = bytearray(base[i % len(base)])
        sos = bytes(b).index(b"\xff\xda")
        if i % 4 == 0:
            for pos in rng.integers(sos + 14, len(b) - 2, size=int(rng.integers(1, 4))):
                b[pos] = int(rng.integers(0, 256))
        elif i % 4 == 1:
            b = b[:int(rng.integers(sos + 20, len(b)))]
        elif i % 4 == 2:
            pos = int(rng.integers(sos + 14, len(b) - 4))
            del b[pos:pos + int(rng.integers(1, 4))]
        files.append(bytes(b))
    got, gi = gpu_ctx.decode_jpegs(files, strict=False, entropy="gpu")
    ref, ri = gpu_ctx.decode_jpegs(files, strict=False, entropy="cpu")
    on_device = 0
    for k, (a, b) in enumerate(zip(gi, ri)):
        assert a.status == b.status, (k, a.message, b.message)
        if a.status == 0:
            on_device += a.tasks > 1
            assert np.array_equal(got[a.rgb_off:a.rgb_off + a.rgb_len], ref[b.rgb_off:b.rgb_off + b.rgb_len]), k
    assert on_device > 100
    capfd.readouterr()


@pytest.mark.parametrize("knob", ["JGPU_HUFF_WRITE=scatter", "JGPU_BLOCKS_PER_GROUP=1", "JGPU_HUFF_WAVES=1"])
def test_a_b_knobs_give_the_same_bytes(gpu_ctx, monkeypatch, knob):
    """The write pass that stages blocks in shared memory (product) against the one that stores coefficient
    by coefficient (JGPU_HUFF_WRITE=scatter), the block decoder after every group against once at the end
    (pixels on the device), small groups: same pixels, same planes, files with and without restart markers,
    long codes (optimised tables) and blocks that straddle subsequences everywhere."""
    pytest.importorskip("PIL")
    import torch
    files = [_jpeg(1920, 1080, 2, rst=120), _jpeg(1000, 563, 1, optimize=True, q=95), _jpeg(333, 222, "L", rst=5),
             _jpeg(640, 480, 0, q=30), load("c444_40x24")[0], _jpeg(1280, 720, 2, q=98, noise=90)] * 3
    total, _ = J.probe_jpegs(files)
    ref, ri = gpu_ctx.decode_jpegs(files, entropy="gpu")
    ref_yuv, ry = gpu_ctx.decode_jpegs(files, entropy="gpu", out="yuv")
    name, value = knob.split("=")
    monkeypatch.setenv(name, value)
    ctx = J.Context(0)   # the knobs are read when a context is created / per call
    try:
        def same(got, gi, want, wi):   # image by image: the bytes between images belong to nobody
            for a, b in zip(gi, wi):
                assert a.status == 0 and (a.rgb_off, a.rgb_len) == (b.rgb_off, b.rgb_len)
                assert np.array_equal(got[a.rgb_off:a.rgb_off + a.rgb_len], want[b.rgb_off:b.rgb_off + b.rgb_len])
        dev = torch.zeros(total, dtype=torch.uint8, device="cuda:0")
        out, gi = ctx.decode_jpegs(files, dev, entropy="gpu")
        same(out.cpu().numpy(), gi, ref, ri)
        got, gi = ctx.decode_jpegs(files, entropy="gpu")
        same(got, gi, ref, ri)
        got_yuv, gy = ctx.decode_jpegs(files, entropy="gpu", out="yuv")
        same(got_yuv, gy, ref_yuv, ry)
    finally:
        ctx.close()
        monkeypatch.delenv(name)
        J.Context(0).close()


def test_files_longer_than_one_sweep_of_the_scan_kernel(gpu_ctx):
    """k_huff_scan takes 16 384 subsequences per sweep and carries the running count into the next one: 4K files
    of 4 and 8 MB (two and four sweeps; with and without restart markers, where the carry matters for the whole
    file) give the host-thread path's pixels, decoded on the device (tasks = subsequences)."""
    pytest.importorskip("PIL")
    files = [_jpeg(3840, 2160, 2, rst=240, q=85, noise=24), _jpeg(3840, 2160, 2, rst=0, q=85, noise=24),
             _jpeg(3840, 2160, 2, rst=0, q=96, noise=40)]
    got, gi = gpu_ctx.decode_jpegs(files, entropy="gpu")
    ref, ri = gpu_ctx.decode_jpegs(files, entropy="cpu")
    assert [i.status for i in gi] == [0, 0, 0]
    assert all(i.tasks > 16384 for i in gi), [i.tasks for i in gi]
    assert gi[2].tasks > 3 * 16384
    for a, b in zip(gi, ri):
        assert np.array_equal(got[a.rgb_off:a.rgb_off + a.rgb_len], ref[b.rgb_off:b.rgb_off + b.rgb_len])
