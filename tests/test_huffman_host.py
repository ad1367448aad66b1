"""GPU entropy decoder, checked on the CPU: the per-thread decoding loop, sinks and address
arithmetic of jgpu_huff_core.h (the code the kernels run) driven by a host emulation of the
kernels' choreography (tests/host_core/huff_host.cpp), against the sequential reader and the
reference reader's own QUANT planes (goldens, src/xjpeg.c:449-632)."""
import ctypes as C
import io
import os
import subprocess

import numpy as np
import pytest

import jpeg_gpu_b200 as J
from golden_util import NAMES, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "jpeg_gpu_b200", "csrc")


WRITE_MODES = {"scatter": 0, "staged": 1, "staged_descending": 2}


@pytest.fixture(scope="module", params=list(WRITE_MODES))
def emu(tmp_path_factory, request):
    """The emulation, once per write pass it can restate: k_huff_write (a store per coefficient) and
    k_huff_write_staged (whole blocks from a buffer, shared blocks as their non-zero coefficients; threads in
    ascending and in descending order, so that a shared block taken for a whole one cannot go unnoticed)."""
    d = tmp_path_factory.mktemp("huff_host")
    objs = []
    inc = ["-I", CSRC, "-I", os.path.join(ROOT, "include")]
    for src in ("jgpu_front.c", "jgpu_host.c", "jgpu_huff_prep.c"):
        o = str(d / (src + ".o"))
        subprocess.run(["gcc", "-std=c99", "-O2", "-fPIC", *inc, "-c", "-o", o, os.path.join(CSRC, src)], check=True)
        objs.append(o)
    so = str(d / "huff_host.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", *inc, "-o", so,
                    os.path.join(ROOT, "tests", "host_core", "huff_host.cpp"), *objs], check=True)
    lib = C.CDLL(so)
    lib.huff_emulate.restype = C.c_longlong
    lib.huff_emulate.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                 C.POINTER(C.c_uint), C.POINTER(C.c_int)]
    lib.huff_table_selfcheck.argtypes = [C.c_char_p, C.c_char_p]
    lib.huff_set_write_mode(WRITE_MODES[request.param])
    return lib


def emulate(lib, jpg, subseq_words=32, cta=256, passes=3):
    cap = 1 << 26
    coef = np.zeros(cap, dtype=np.int16)
    status = C.c_uint(0)
    stats = (C.c_int * 4)()
    n = lib.huff_emulate(jpg, len(jpg), coef.ctypes.data_as(C.c_void_p), cap, subseq_words, cta, passes,
                         C.byref(status), stats)
    return n, coef[:max(n, 0)], status.value, list(stats)


def sequential_quant(jpg):
    with J.Decoder(jpg, impl="jfront") as dec:
        dec.decode_header()
        return dec.decode_image("quant")["coef"]


def test_decoder_tables_agree_with_the_canonical_procedure(emu):
    """jgpu_huff_build_table + lookup on all 65536 windows == T.81 F.2.2.3, for the Annex K
    luminance AC table and for a skewed table with 16-bit codes."""
    k5_counts = bytes([0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d])
    k5_syms = bytes(range(162))
    assert emu.huff_table_selfcheck(k5_counts, k5_syms) == 0
    skew = bytes([1] * 15 + [2])
    assert emu.huff_table_selfcheck(skew, bytes(range(17))) == 0
    assert emu.huff_table_selfcheck(bytes([3] + [0] * 15), bytes(range(3))) == -1   # over-subscribed


def test_look_up_entries_carry_what_the_loop_assumes(emu):
    """JGPU_HUFF_ENTRY: T = code length + extra bits (at most 31, so one 32-bit window holds a symbol), s = extra
    bits, A = advance of the zig-zag index.  The loop ends a block when index + A >= 64 and stores a coefficient only
    when index + A <= 64; for that to be the reader's behaviour an end-of-block must land beyond every run
    (index + A > 63 + 16 for every index an AC symbol can meet) and a run past coefficient 63 must stay below it."""
    emu.huff_entry.restype = emu.huff_entry_field.restype = C.c_uint
    T, S, A = (lambda e: emu.huff_entry_field(e, 0)), (lambda e: emu.huff_entry_field(e, 1)), (lambda e: emu.huff_entry_field(e, 2))
    for length in range(1, 17):
        for sym in range(256):
            for ac in (0, 1):
                e = emu.huff_entry(length, sym, ac)
                assert 0 < e < 65536, "fits the 16-bit table, and 0 stays free for 'no entry'"
                assert T(e) == length + (sym & 15) <= 31 and S(e) == sym & 15
                if not ac:
                    assert A(e) == 1                      # DC: index 0 -> 1 whatever the high nibble says
                elif sym == 0:
                    assert A(e) == 96                     # end of block
                else:
                    assert A(e) == (sym >> 4) + 1         # run, then the coefficient (0xF0: sixteen zeros)
    eob, longest_run = 96, 16
    assert all(z + eob > 63 + longest_run for z in range(1, 64)), "an end of block is never taken for a run"
    assert all(z + a < 1 + eob for z in range(1, 64) for a in range(1, longest_run + 1)), "nor a run for an end of block"


@pytest.mark.parametrize("name", NAMES)
def test_golden_files_match_the_reference_readers_planes(emu, name):
    """Kernel geometry shrunk (1-word subsequences, 4-thread CTAs) so that these small files
    cross subsequence, CTA and launch boundaries; planes == xjpeg's own QUANT output."""
    jpg, z, _ = load(name)
    for (words, cta, passes) in ((1, 4, 400), (2, 8, 100), (4, 16, 12), (32, 256, 3)):
        n, coef, status, stats = emulate(emu, jpg, words, cta, passes)
        assert n == z["quant"].size and status == 0, (name, words, cta, status, stats)
        assert np.array_equal(coef, z["quant"].reshape(-1)), (name, words, cta)


def _photo_like(w, h, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 5 + yy * 3) % 256, (yy * 7 + xx) % 256, (xx * 2 + yy * 9) % 256], -1)
    return np.clip(base + rng.integers(-40, 41, size=base.shape), 0, 255).astype(np.uint8)


CASES = [(640, 480, 2, 0, 85), (640, 480, 2, 40, 85), (517, 389, 1, 7, 60), (512, 512, 0, 0, 95),
         (333, 222, "L", 0, 75), (333, 222, "L", 5, 30), (1920, 1080, 2, 0, 85), (1280, 720, 2, 80, 50)]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}x{c[1]}_{c[2]}_rst{c[3]}_q{c[4]}" for c in CASES])
def test_generated_files_match_the_sequential_reader(emu, case):
    Image = pytest.importorskip("PIL.Image")
    w, h, ss, rst, q = case
    img = _photo_like(w, h, w + h)
    bio = io.BytesIO()
    if ss == "L":
        Image.fromarray(img[..., 0]).save(bio, "JPEG", quality=q, restart_marker_blocks=rst)
    else:
        Image.fromarray(img).save(bio, "JPEG", quality=q, subsampling=ss, restart_marker_blocks=rst,
                                  optimize=(q == 60))
    jpg = bio.getvalue()
    want = sequential_quant(jpg)
    n, coef, status, stats = emulate(emu, jpg)
    assert n == want.size and status == 0, (status, stats)
    assert np.array_equal(coef, want)
    # the states settle in a handful of rounds: that is what makes the method parallel
    assert stats[2] <= 24, stats


def test_too_few_sync_launches_are_detected_not_decoded_wrong(emu):
    """With tiny CTAs and a single sync launch the states cannot cross CTA boundaries: the
    write pass must flag the file (the runtime then falls back to the sequential reader)."""
    jpg, z, _ = load("c420_q10_96x64")
    n, coef, status, stats = emulate(emu, jpg, 1, 2, 1)
    assert n > 0 and (status & 2), (status, stats)


def test_truncated_scan_is_flagged(emu):
    jpg = load("c420_64x48")[0]
    n, coef, status, stats = emulate(emu, jpg[:len(jpg) * 2 // 3])
    assert n > 0 and (status & 4), (status, stats)


def test_corrupted_scans_are_flagged_or_decode_like_the_sequential_reader(emu, capfd):
    """The write pass's verification, fuzzed: bytes flipped anywhere after the headers.  Whenever
    the emulated kernels report a clean status, the sequential reader must accept the file too and
    produce the same planes -- a damaged file may be handed to the fallback, never decoded wrong."""
    rng = np.random.default_rng(11)
    clean = flagged = 0
    for name in ("c420_64x48", "c420_rst_80x48", "c444_q100_24x24", "gray_48x40", "c422_rst_33x17"):
        jpg = load(name)[0]
        sos = jpg.index(b"\xff\xda")
        for trial in range(40):
            b = bytearray(jpg)
            for pos in rng.integers(sos + 14, len(b) - 2, size=int(rng.integers(1, 4))):
                b[pos] = int(rng.integers(0, 256))
            bad = bytes(b)
            words, cta = ((32, 256), (2, 8))[trial & 1]
            n, coef, status, stats = emulate(emu, bad, words, cta, 64)
            if n < 0:
                flagged += 1          # not eligible (e.g. a restart marker was hit): sequential reader's job
                continue
            if status != 0:
                flagged += 1
                continue
            clean += 1
            want = sequential_quant(bad)      # raises if the sequential reader rejects the file
            assert np.array_equal(coef, want), (name, trial)
    assert clean > 20 and flagged > 20, (clean, flagged)
    capfd.readouterr()


def _with_fill_bytes(jpg):
    """0xFF fill bytes in front of every RSTn marker (legal, T.81 B.1.1.2)."""
    out = bytearray()
    i = 0
    sos = jpg.index(b"\xff\xda")
    while i < len(jpg):
        if i > sos and jpg[i] == 0xFF and i + 1 < len(jpg) and 0xD0 <= jpg[i + 1] <= 0xD7:
            out += b"\xff\xff"
        out.append(jpg[i])
        i += 1
    return bytes(out)


def test_marker_corner_cases_of_the_preparation(emu):
    """Fill bytes before restart markers, more than eight restart intervals (the RSTn counter
    wraps), a restart interval longer than the image, one MCU per interval: planes as from the
    sequential reader, decoded by the emulated kernels (status 0, not through a fallback)."""
    Image = pytest.importorskip("PIL.Image")
    img = _photo_like(160, 96, 5)
    cases = []
    for ss, rst in ((2, 1), (2, 3), (1, 2), (0, 50), (2, 1000)):
        bio = io.BytesIO()
        Image.fromarray(img).save(bio, "JPEG", quality=80, subsampling=ss, restart_marker_blocks=rst)
        cases.append(bio.getvalue())
    cases += [_with_fill_bytes(c) for c in cases[:3]] + [_with_fill_bytes(load("c420_rst_80x48")[0])]
    for k, jpg in enumerate(cases):
        want = sequential_quant(jpg)
        for words, cta, passes in ((32, 256, 3), (1, 8, 200)):
            n, coef, status, stats = emulate(emu, jpg, words, cta, passes)
            assert n == want.size and status == 0, (k, words, status, stats)
            assert np.array_equal(coef, want), (k, words)
