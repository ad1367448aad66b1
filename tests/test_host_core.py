"""jgpu_idct_core.cuh built for the host (JGPU_CORE_HOST_EMULATION, two plain floats per
pair, -ffp-contract=off): the device core's operation order against the oracle, no GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

from golden_util import blocks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    so = tmp_path / "core_host.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I",
                    os.path.join(ROOT, "jpeg_gpu_b200", "csrc"), "-o", str(so),
                    os.path.join(ROOT, "tests", "host_core", "core_host.cpp")], check=True)
    return C.CDLL(str(so))


def test_fixed_point_colour_offsets_equal_the_binary32_definition_everywhere(tmp_path, port):
    """jgpu_colour_fixed.h (what the fused kernel executes) against the oracle's binary32
    definition (jgo_colour_offsets, res/yuv.fs.glsl:11-23) for ALL 65536 (Cb, Cr): exact."""
    import oracle
    lib = _build(tmp_path)
    got = np.zeros((256, 256, 3), dtype=np.int32)
    lib.core_colour_fixed_all(C.c_void_p(got.ctypes.data))
    cb, cr = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    ro, go, bo = oracle.np_colour_offsets(cb, cr)
    assert np.array_equal(got[..., 0], ro) and np.array_equal(got[..., 1], go) and np.array_equal(got[..., 2], bo)
    # and the C oracle itself, every input (the numpy mirror is only spot-checked elsewhere)
    for c in range(256):
        for r in range(0, 256):
            assert tuple(got[c, r]) == port.colour_offsets(c, r), (c, r)


def test_device_core_operation_order(tmp_path, port):
    lib = _build(tmp_path)
    coef, want = blocks()
    rng = np.random.default_rng(2)
    more = rng.integers(-2048, 2048, size=(20000, 8, 8), dtype=np.int16)
    for blk, exp in ((coef[:3500], want[:3500]), (more, port.idct_blocks(more))):
        out = np.zeros_like(blk)
        lib.core_idct_pairs(C.c_void_p(blk.ctypes.data), C.c_void_p(out.ctypes.data), C.c_longlong(blk.shape[0] // 2))
        assert np.array_equal(out, exp)


def test_constants_are_the_references(port):
    """Bit patterns hard-coded in jgpu_idct_core.cuh == the reference's double literals
    converted to float (src/dct.c:51,62-65,89-98)."""
    import re
    import ctypes
    out = np.zeros(12, dtype=np.float32)
    port.lib.jgo_idct_constants.argtypes = [ctypes.c_void_p]
    port.lib.jgo_idct_constants(out.ctypes.data_as(ctypes.c_void_p))
    bits = [int(v) for v in out.view(np.uint32)]
    src = open(os.path.join(ROOT, "jpeg_gpu_b200", "csrc", "jgpu_idct_core.cuh")).read()
    scale = [int(x, 16) for x in re.findall(r"k == \d \? (0x[0-9a-f]+)u", src)] + \
            [int(re.search(r": (0x[0-9a-f]+)u;\n\}", src).group(1), 16)]
    assert scale == bits[:8]
    for name, want in zip(["kSqrt2Bits", "k18477Bits", "k10823Bits", "k26131Bits"], bits[8:]):
        assert int(re.search(name + r" = (0x[0-9a-f]+)u", src).group(1), 16) == want
