"""The C ABI: the library loads on a CPU-only box, exports every symbol
include/jpeg_gpu_b200.h declares, and the interface structs in include/jgpu_ref_abi.h
have the layout of the reference's own headers (src/jpeg_wrap.h, src/image.h,
src/jpeg_info.h).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import jpeg_gpu_b200 as J
from jpeg_gpu_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("JGPU_REFERENCE_ROOT", "/root/reference")


def test_library_loads_and_exports_every_declared_symbol():
    lib = _capi.lib()
    header = open(os.path.join(ROOT, "include", "jpeg_gpu_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(jgpu_[a-z_0-9]+|cuda_decode_[a-z_]+)\s*\(", header))
    declared |= set(re.findall(r"extern const jpeg_decode_ctx_vtbl (\w+);", header))
    assert {"jgpu_create", "jgpu_plan_run", "jgpu_decode_batch_host", "jgpu_decode_image",
            "CUDA_DECODE_CTX_VTBL", "JFRONT_DECODE_CTX_VTBL"} <= declared
    bound = {name for name, _, _ in _capi.EXPORTS} | set(_capi.DATA_EXPORTS)
    assert declared == bound, (declared - bound, bound - declared)
    for name in declared:
        assert hasattr(lib, name), name
    for name in _capi.DATA_EXPORTS:
        vt = _capi.vtbl(name)
        for slot, _ in vt._fields_:
            assert C.cast(getattr(vt, slot), C.c_void_p).value, (name, slot)


def test_no_gpu_is_reported_not_crashed():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this test is for CPU-only boxes")
    assert _capi.lib().jgpu_device_count() == 0
    assert not _capi.lib().jgpu_create(0)
    assert "CUDA" in _capi.last_error() or "device" in _capi.last_error()
    with pytest.raises(RuntimeError):
        J.Context(0)


PROBE = r'''
#include <stddef.h>
#include <stdio.h>
%s
#define S(t) printf("sizeof " #t " %%zu\n", sizeof(t))
#define O(t, m) printf("offsetof " #t "." #m " %%zu\n", offsetof(t, m))
int main(void) {
  S(jpeg_quant); O(jpeg_quant, valid); O(jpeg_quant, bits); O(jpeg_quant, tbl);
  S(jpeg_component); O(jpeg_component, hblocks); O(jpeg_component, vblocks); O(jpeg_component, hsamp);
  O(jpeg_component, vsamp); O(jpeg_component, quant);
  S(jpeg_header); O(jpeg_header, bits); O(jpeg_header, width); O(jpeg_header, height); O(jpeg_header, ncomps);
  O(jpeg_header, subsamp); O(jpeg_header, restart_interval); O(jpeg_header, comp); O(jpeg_header, quant);
  S(jpeg_info); O(jpeg_info, size); O(jpeg_info, buf);
  S(image_plane); O(image_plane, bitdepth); O(image_plane, xdec); O(image_plane, ydec); O(image_plane, xstride);
  O(image_plane, ystride); O(image_plane, width); O(image_plane, height); O(image_plane, data);
  O(image_plane, coef); O(image_plane, cstride); O(image_plane, packed); O(image_plane, index);
  S(image); O(image, width); O(image, height); O(image, nplanes); O(image, plane); O(image, coef);
  O(image, packed); O(image, index); O(image, pixels);
  S(jpeg_decode_ctx_vtbl); O(jpeg_decode_ctx_vtbl, decode_alloc); O(jpeg_decode_ctx_vtbl, decode_header);
  O(jpeg_decode_ctx_vtbl, decode_image); O(jpeg_decode_ctx_vtbl, decode_reset); O(jpeg_decode_ctx_vtbl, decode_free);
  printf("enum %%d %%d %%d %%d %%d %%d\n", JPEG_DECODE_PACK, JPEG_DECODE_QUANT, JPEG_DECODE_DCT, JPEG_DECODE_YUV,
         JPEG_DECODE_RGB, JPEG_DECODE_OUT_MAX);
  printf("subsamp %%d %%d %%d %%d %%d %%d %%d\n", JPEG_SUBSAMP_UNKNOWN, JPEG_SUBSAMP_444, JPEG_SUBSAMP_422,
         JPEG_SUBSAMP_420, JPEG_SUBSAMP_440, JPEG_SUBSAMP_411, JPEG_SUBSAMP_MONO);
  return 0;
}
'''


def _probe(tmp_path, tag, include_line, flags):
    src = tmp_path / f"probe_{tag}.c"
    exe = tmp_path / f"probe_{tag}"
    src.write_text(PROBE % include_line)
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)] + flags, check=True)
    return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout


def test_interface_structs_match_reference_headers(tmp_path):
    ours = _probe(tmp_path, "ours", '#include "jgpu_ref_abi.h"', ["-I", os.path.join(ROOT, "include")])
    assert "sizeof image " in ours and "enum 0 1 2 3 4 5" in ours
    # ctypes mirror == C header
    sizes = dict(re.findall(r"sizeof (\w+) (\d+)", ours))
    for name in ("jpeg_quant", "jpeg_component", "jpeg_header", "jpeg_info", "image_plane", "image",
                 "jpeg_decode_ctx_vtbl"):
        assert C.sizeof(getattr(_capi, name)) == int(sizes[name]), name
    for (t, m, off) in re.findall(r"offsetof (\w+)\.(\w+) (\d+)", ours):
        assert getattr(getattr(_capi, t), m).offset == int(off), (t, m)
    if not os.path.exists(os.path.join(REF, "src", "jpeg_wrap.h")):
        pytest.skip("reference headers not present (GPU box): C-vs-ctypes checked only")
    theirs = _probe(tmp_path, "ref", '#include <stdlib.h>\n#include "jpeg_wrap.h"', ["-I", os.path.join(REF, "src")])
    assert ours == theirs
    # and the documented switch to the reference's own headers compiles
    both = _probe(tmp_path, "switch", '#include <stdlib.h>\n#define JGPU_USE_REFERENCE_HEADERS 1\n#include "jpeg_gpu_b200.h"',
                  ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(REF, "src")])
    assert both == theirs


def test_batch_structs_match_header(tmp_path):
    src = tmp_path / "p.c"
    src.write_text('#include <stdio.h>\n#include "jpeg_gpu_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(jgpu_image_desc), sizeof(jgpu_layout), sizeof(jgpu_plane_layout),'
                   'offsetof(jgpu_image_desc, coef_off), offsetof(jgpu_layout, plane));return 0;}')
    exe = tmp_path / "p"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    a = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert a == [C.sizeof(_capi.jgpu_image_desc), C.sizeof(_capi.jgpu_layout), C.sizeof(_capi.jgpu_plane_layout),
                 _capi.jgpu_image_desc.coef_off.offset, _capi.jgpu_layout.plane.offset]
