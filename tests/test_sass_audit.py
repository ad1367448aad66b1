"""Static audit of the shipped sm_100a machine code (cuobjdump, no GPU needed).

Bit-exactness with src/dct.c forbids fused multiply-adds in the IDCT.  ptxas 12.9 fuses
`mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 despite the explicit .rn, so the core issues every
product as fma(a, b, -0.0) with -0.0 in a register loaded from __constant__ memory
(jgpu_idct_core.cuh).  This test checks that the SASS has that shape: no packed multiply at
all, every packed FMA adds that one register, no scalar FFMA in the kernels, and that the
kernels really use the Blackwell features the design rests on (FADD2/FFMA2, TMA, mbarrier)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jpeg_gpu_b200", "libjpeg_gpu_b200.so")


@pytest.fixture(scope="module")
def sass():
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            funcs[name].append(line)
    return funcs


def test_only_sm100a_code(sass):
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_idct_has_no_contracted_multiply_add(sass):
    kernels = {n: body for n, body in sass.items()
               if "k_fused" in n or "k_coef_to_planes" in n or "k_mcu" in n or "k_tk" in n}
    # generic + 5 modes x {8-bit, 16-bit tables} of k_fused, + 5 modes x 2 x {pixels, planes} of k_mcu, + 6 modes
    # x 2 x {pixels, pixels with rows of any alignment, planes} of k_tk (the product path)
    assert len(kernels) >= 11 + 20 + 36
    assert sum("k_tk" in n for n in kernels) == 36
    for name, body in kernels.items():
        text = "\n".join(body)
        assert "FMUL2" not in text, name
        ffma2 = [l for l in body if re.search(r"\bFFMA2\b", l)]
        fadd2 = [l for l in body if re.search(r"\bFADD2\b", l)]
        # 2 x 64 prescale products + 16 passes x 5 products = 208 packed products per block pair
        # (the fused kernel adds 4 per exchange word for the colour offsets, 16 or 32 in all)
        assert len(ffma2) in (208, 208 + 16, 208 + 32), (name, len(ffma2))
        assert len(fadd2) >= 16 * 29 + 8 + 64, (name, len(fadd2))
        # the addend of every packed FMA is a register pair loaded from c_negzero2 (bank 3),
        # or a MOV copy of such a pair (ptxas duplicates it now and then to dodge bank conflicts)
        negzero_regs = set(re.findall(r"LDC\.64 (R\d+), c\[0x3\]", text))
        assert negzero_regs, name
        copies = {d for d, src in re.findall(r"\bMOV (R\d+), (R\d+)(?:\.reuse)? ;", text) if src in negzero_regs}
        addends = []
        for l in ffma2:
            ops = l.split("FFMA2", 1)[1].split(";")[0].split(",")
            addends.append(ops[-1].strip().split(".")[0])
        assert set(addends) <= negzero_regs | copies, (name, set(addends) - negzero_regs - copies)
        assert sum(a in negzero_regs for a in addends) >= 0.9 * len(addends), name
        assert not re.search(r"\bFFMA\b", text), name        # no scalar contraction either


def test_colour_stage_has_no_fma(sass):
    for name, body in sass.items():
        if "k_planes_to_rgb" in name:
            assert not any(re.search(r"\bFFMA\b", l) for l in body), name


def test_product_kernel_splits_registers_between_its_warp_roles(sass):
    """k_tk: transform warps raise their register allowance, colour warps give theirs back."""
    tk = {n: "\n".join(body) for n, body in sass.items() if "k_tk" in n}
    assert tk
    for name, text in tk.items():
        assert "USETMAXREG.TRY_ALLOC" in text and "USETMAXREG.DEALLOC" in text, name
        t_part = text.split("USETMAXREG.DEALLOC")[0]
        # the transform keeps its 64 sample pairs in registers: no spills to speak of (4:4:0 at 176 registers: 6)
        assert len(re.findall(r"\b(STL|LDL)\b", t_part.split("USETMAXREG.TRY_ALLOC")[1])) <= 8, name


def test_fused_kernel_uses_tma_and_mbarriers(sass):
    fused = [body for n, body in sass.items() if "k_fused" in n or "k_mcu" in n or "k_tk" in n]
    assert len(fused) >= 50
    for body in fused:
        text = "\n".join(body)
        assert "UTMALDG" in text          # cp.async.bulk.tensor
        assert "UBLKCP" in text           # cp.async.bulk (tables, tile descriptors)
        assert "SYNCS" in text            # mbarrier arrive / try_wait
        assert "IDP.2A" in text           # packed-table dequantisation
        assert "VIADDMNMX" in text        # s16x2 add+clamp
        assert "STG.E.EF.128" in text or "STG.E.128" in text or re.search(r"STG\.E\.[A-Z.]*128", text)


def test_entropy_kernels_keep_their_shape(sass):
    """The entropy decoder's two hot kernels (jgpu_huff.cu): the sync pass must keep the register count
    that lets six CTAs of 256 threads share an SM and must not stage scan words in shared memory (they come through
    L1: LDG.E.CONSTANT in the loop); the write pass must flush finished blocks as whole words per lane after a vote
    (VOTE + SHFL + one STG.E per lane and block), not as 2-byte stores from the decoding loop; neither may spill."""
    sync = [body for n, body in sass.items() if "k_huff_sync" in n]
    staged = [body for n, body in sass.items() if "k_huff_write_staged" in n]
    assert len(sync) == 1 and len(staged) == 1
    for body in sync + staged:
        text = "\n".join(body)
        assert not re.search(r"\b(STL|LDL)\b", text), "spills in an entropy kernel"
    s = "\n".join(sync[0])
    assert "LDG.E.CONSTANT" in s and "SHF.L.W.U32.HI" in s, "scan words through L1, funnel-shift window"
    w = "\n".join(staged[0])
    assert "VOTE.ANY" in w and "SHFL.IDX" in w and "STS.U16" in w
    assert re.search(r"LDS\.128", w), "block record: one 16-byte shared load per block end"
    # 32-bit stores of whole words (full blocks) and predicated 16-bit stores (partial blocks) exist; nothing wider
    # than that goes to the coefficient planes from this kernel
    assert re.search(r"STG\.E\s+desc", w) and "STG.E.U16" in w
