"""ctypes binding of include/jpeg_gpu_b200.h and include/jgpu_ref_abi.h.

The shared library is built in-tree by ``jpeg_gpu_b200/csrc/Makefile``
(``__graft_entry__.build()``).  There is no CPU fallback: if the library is
missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# JGPU_LIB_PATH selects a tuning variant built by csrc/Makefile (VARIANT=...); default is the product build
LIB_PATH = os.environ.get("JGPU_LIB_PATH") or os.path.join(HERE, "libjpeg_gpu_b200.so")

NCOMPS_MAX = 3
NQUANT_MAX = 4
NPLANES_MAX = 3

JPEG_DECODE_PACK, JPEG_DECODE_QUANT, JPEG_DECODE_DCT, JPEG_DECODE_YUV, JPEG_DECODE_RGB = range(5)
OUT_NAMES = {"pack": 0, "quant": 1, "dct": 2, "yuv": 3, "rgb": 4}
SUBSAMP_NAMES = ["Unknown", "4:4:4", "4:2:2", "4:2:0", "4:4:0", "4:1:1", "Mono"]

JGPU_OUT_RGB, JGPU_OUT_YUV, JGPU_FORCE_GENERIC = 1, 2, 4
JGPU_IMAGE_PINNED = 1


class cuda_decode_options(C.Structure):
    _fields_ = [("frontend", C.c_void_p), ("device", C.c_int), ("upload", C.c_int), ("entropy_on_device", C.c_int)]


# ---- include/jgpu_ref_abi.h --------------------------------------------------
class jpeg_quant(C.Structure):
    _fields_ = [("valid", C.c_int), ("bits", C.c_ubyte), ("tbl", C.c_ushort * 64)]


class jpeg_component(C.Structure):
    _fields_ = [("hblocks", C.c_int), ("vblocks", C.c_int), ("hsamp", C.c_int), ("vsamp", C.c_int),
                ("quant", C.POINTER(jpeg_quant))]


class jpeg_header(C.Structure):
    _fields_ = [("bits", C.c_int), ("width", C.c_int), ("height", C.c_int), ("ncomps", C.c_int),
                ("subsamp", C.c_int), ("restart_interval", C.c_int),
                ("comp", jpeg_component * NCOMPS_MAX), ("quant", jpeg_quant * NQUANT_MAX)]


class jpeg_info(C.Structure):
    _fields_ = [("size", C.c_int), ("buf", C.POINTER(C.c_ubyte))]


class image_plane(C.Structure):
    _fields_ = [("bitdepth", C.c_int), ("xdec", C.c_ubyte), ("ydec", C.c_ubyte), ("xstride", C.c_int),
                ("ystride", C.c_int), ("width", C.c_ushort), ("height", C.c_ushort),
                ("data", C.POINTER(C.c_ubyte)), ("coef", C.POINTER(C.c_short)), ("cstride", C.c_int),
                ("packed", C.c_int), ("index", C.POINTER(C.c_int))]


class image(C.Structure):
    _fields_ = [("width", C.c_ushort), ("height", C.c_ushort), ("nplanes", C.c_int),
                ("plane", image_plane * NPLANES_MAX), ("coef", C.POINTER(C.c_short)), ("packed", C.c_int),
                ("index", C.POINTER(C.c_int)), ("pixels", C.POINTER(C.c_ubyte))]


decode_alloc_func = C.CFUNCTYPE(C.c_void_p, C.POINTER(jpeg_info))
decode_header_func = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(jpeg_header))
decode_image_func = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(image), C.c_int)
decode_reset_func = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(jpeg_info))
decode_free_func = C.CFUNCTYPE(None, C.c_void_p)


class jpeg_decode_ctx_vtbl(C.Structure):
    _fields_ = [("decode_alloc", decode_alloc_func), ("decode_header", decode_header_func),
                ("decode_image", decode_image_func), ("decode_reset", decode_reset_func),
                ("decode_free", decode_free_func)]


# ---- include/jpeg_gpu_b200.h --------------------------------------------------
class jgpu_image_desc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("ncomps", C.c_int32),
                ("hsamp", C.c_int32 * 3), ("vsamp", C.c_int32 * 3), ("tq", C.c_int32 * 3),
                ("qtab_set", C.c_int32), ("reserved", C.c_int32),
                ("coef_off", C.c_int64), ("rgb_off", C.c_int64), ("yuv_off", C.c_int64)]


class jgpu_plane_layout(C.Structure):
    _fields_ = [("hblocks", C.c_int32), ("vblocks", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("xdec", C.c_int32), ("ydec", C.c_int32), ("cstride", C.c_int32), ("reserved", C.c_int32),
                ("coef_off", C.c_int64), ("data_off", C.c_int64)]


class jgpu_layout(C.Structure):
    _fields_ = [("nhmb", C.c_int32), ("nvmb", C.c_int32), ("hmax", C.c_int32), ("vmax", C.c_int32),
                ("coef_len", C.c_int64), ("coded_blocks", C.c_int64), ("data_len", C.c_int64),
                ("rgb_len", C.c_int64), ("plane", jgpu_plane_layout * 3)]


class jgpu_jpeg(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_int64)]


class jgpu_jpeg_info(C.Structure):
    _fields_ = [("status", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("ncomps", C.c_int32),
                ("hsamp0", C.c_int32), ("vsamp0", C.c_int32), ("restart_interval", C.c_int32), ("tasks", C.c_int32),
                ("rgb_off", C.c_int64), ("rgb_len", C.c_int64), ("message", C.c_char_p)]


EXPORTS = [
    # name, restype, argtypes  (every function include/jpeg_gpu_b200.h declares)
    ("cuda_decode_set_frontend", None, [C.c_void_p]),
    ("cuda_decode_set_device", None, [C.c_int]),
    ("cuda_decode_set_upload", C.c_int, [C.c_int]),
    ("cuda_decode_set_entropy", C.c_int, [C.c_int]),
    ("cuda_decode_get_options", None, [C.c_void_p]),
    ("cuda_decode_alloc_ex", C.c_void_p, [C.POINTER(jpeg_info), C.c_void_p]),
    ("jgpu_image_set_pinned", None, [C.c_int]),
    ("jgpu_image_init_ex", C.c_int, [C.POINTER(image), C.POINTER(jpeg_header), C.c_uint]),
    ("jgpu_host_is_pinned", C.c_int, [C.c_void_p]),
    ("jgpu_decode_image_packed", C.c_int, [C.c_void_p, C.POINTER(jpeg_header), C.POINTER(image), C.c_int64, C.c_int]),
    ("jgpu_image_init", C.c_int, [C.POINTER(image), C.POINTER(jpeg_header)]),
    ("jgpu_image_zero", None, [C.POINTER(image)]),
    ("jgpu_image_clear", None, [C.POINTER(image)]),
    ("jgpu_info_init", C.c_int, [C.POINTER(jpeg_info), C.c_char_p]),
    ("jgpu_info_clear", None, [C.POINTER(jpeg_info)]),
    ("jgpu_layout_query", C.c_int, [C.POINTER(jgpu_image_desc), C.POINTER(jgpu_layout)]),
    ("jgpu_last_error", C.c_char_p, []),
    ("jgpu_device_count", C.c_int, []),
    ("jgpu_create", C.c_void_p, [C.c_int]),
    ("jgpu_destroy", None, [C.c_void_p]),
    ("jgpu_plan_create", C.c_void_p, [C.c_void_p, C.POINTER(jgpu_image_desc), C.c_int, C.c_uint]),
    ("jgpu_plan_destroy", None, [C.c_void_p]),
    ("jgpu_plan_launches", C.c_int, [C.c_void_p]),
    ("jgpu_plan_bytes", C.c_int64, [C.c_void_p]),
    ("jgpu_plan_run", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("jgpu_decode_batch_host", C.c_int, [C.c_void_p, C.POINTER(jgpu_image_desc), C.c_int, C.c_uint, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    ("jgpu_decode_image", C.c_int, [C.c_void_p, C.POINTER(jpeg_header), C.POINTER(image), C.c_int]),
    ("jgpu_pack_bound", C.c_int64, [C.POINTER(jgpu_image_desc)]),
    ("jgpu_pack_from_quant", C.c_int64, [C.POINTER(jgpu_image_desc), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("jgpu_plan_unpack", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("jgpu_decode_batch_host_packed", C.c_int, [C.c_void_p, C.POINTER(jgpu_image_desc), C.c_int, C.c_uint,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                C.c_void_p, C.c_void_p]),
    ("jgpu_jpegs_probe", C.c_int64, [C.POINTER(jgpu_jpeg), C.c_int, C.POINTER(jgpu_jpeg_info)]),
    ("jgpu_jpegs_probe_ex", C.c_int64, [C.POINTER(jgpu_jpeg), C.c_int, C.c_uint, C.POINTER(jgpu_jpeg_info)]),
    ("jgpu_decode_jpegs", C.c_int, [C.c_void_p, C.POINTER(jgpu_jpeg), C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                    C.POINTER(jgpu_jpeg_info)]),
    ("jgpu_decode_jpegs_ex", C.c_int, [C.c_void_p, C.POINTER(jgpu_jpeg), C.c_int, C.c_int, C.c_uint, C.c_void_p,
                                       C.c_int64, C.POINTER(jgpu_jpeg_info)]),
    ("jgpu_host_alloc", C.c_void_p, [C.c_size_t]),
    ("jgpu_host_free", None, [C.c_void_p]),
]
DATA_EXPORTS = ["CUDA_DECODE_CTX_VTBL", "JFRONT_DECODE_CTX_VTBL"]

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads libjpeg_gpu_b200.so; raises LibraryMissing when it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C jpeg_gpu_b200/csrc`. There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, res, args in EXPORTS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def vtbl(name: str) -> jpeg_decode_ctx_vtbl:
    return jpeg_decode_ctx_vtbl.in_dll(lib(), name)


def last_error() -> str:
    return (lib().jgpu_last_error() or b"").decode(errors="replace")
