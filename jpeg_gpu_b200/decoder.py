"""Host-side mirror of the reference's decoder-plugin interface.

``Decoder`` drives a ``jpeg_decode_ctx_vtbl`` backend exported by the C
library exactly the way the reference's main() does (src/jpeg_gpu.c:612-704,
1231-1237): ``decode_alloc -> decode_header -> image_init -> decode_image``
and, in steady state, ``decode_reset -> decode_header -> decode_image``.
Backends: ``"cuda"`` = CUDA_DECODE_CTX_VTBL, ``"jfront"`` =
JFRONT_DECODE_CTX_VTBL (CPU entropy front end, pack/quant/dct only).
Errors follow the reference: a failing slot returns EXIT_FAILURE, which is
raised here as ``DecodeError``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _capi

BACKENDS = {"cuda": "CUDA_DECODE_CTX_VTBL", "jfront": "JFRONT_DECODE_CTX_VTBL"}


class DecodeError(RuntimeError):
    pass


@dataclass
class Header:
    """Python view of jpeg_header (src/jpeg_info.h:54-64)."""
    bits: int
    width: int
    height: int
    ncomps: int
    subsamp: str
    restart_interval: int
    hsamp: List[int]
    vsamp: List[int]
    hblocks: List[int]
    vblocks: List[int]
    tq: List[int]
    qtabs: np.ndarray      # (4, 64) uint16, natural order
    qvalid: List[int]


class Decoder:
    def __init__(self, data: bytes, impl: str = "cuda"):
        if impl not in BACKENDS:
            raise ValueError(f"Invalid decoder implementation: {impl}")
        self._impl = impl
        self._vt = _capi.vtbl(BACKENDS[impl])
        self._buf = (C.c_ubyte * len(data)).from_buffer_copy(data)  # owned by the caller, as in the reference
        self._info = _capi.jpeg_info(len(data), C.cast(self._buf, C.POINTER(C.c_ubyte)))
        self._hdr = _capi.jpeg_header()
        self._img: Optional[_capi.image] = None
        self._have_header = False
        self._dec = self._vt.decode_alloc(C.byref(self._info))
        if not self._dec:
            raise MemoryError("decode_alloc returned NULL")

    # -- the five slots ---------------------------------------------------------
    def decode_header(self) -> Header:
        if self._vt.decode_header(self._dec, C.byref(self._hdr)) != 0:
            raise DecodeError("decode_header failed")
        self._have_header = True
        h = self._hdr
        n = h.ncomps
        base = C.addressof(h.quant)
        tq = [(C.addressof(h.comp[i].quant.contents) - base) // C.sizeof(_capi.jpeg_quant) for i in range(n)]
        q = np.zeros((4, 64), dtype=np.uint16)
        for t in range(4):
            if h.quant[t].valid:
                q[t] = np.ctypeslib.as_array(h.quant[t].tbl)
        return Header(h.bits, h.width, h.height, n, _capi.SUBSAMP_NAMES[h.subsamp], h.restart_interval,
                      [h.comp[i].hsamp for i in range(n)], [h.comp[i].vsamp for i in range(n)],
                      [h.comp[i].hblocks for i in range(n)], [h.comp[i].vblocks for i in range(n)],
                      tq, q, [h.quant[t].valid for t in range(4)])

    def decode_image(self, out: str = "rgb") -> dict:
        """Returns numpy COPIES of the surface members `out` fills."""
        if not self._have_header:
            raise DecodeError("decode_image before decode_header")
        if self._img is None:
            self._img = _capi.image()
            # the CUDA backend reads pixels back into the surface: page-locked memory for it
            rc = _capi.lib().jgpu_image_init_ex(C.byref(self._img), C.byref(self._hdr),
                                                _capi.JGPU_IMAGE_PINNED if self._impl == "cuda" else 0)
            if rc != 0:
                self._img = None
                raise DecodeError("Error initializing image")
            # image_init leaves the buffers uninitialised (src/image.c:61-76); zero them once so
            # the layout's padding rows read as zeros
            _capi.lib().jgpu_image_zero(C.byref(self._img))
        img = self._img
        if out == "pack":
            _capi.lib().jgpu_image_zero(C.byref(img))
        if self._vt.decode_image(self._dec, C.byref(img), _capi.OUT_NAMES[out]) != 0:
            raise DecodeError(f"decode_image({out}) failed")
        res = {"width": img.width, "height": img.height, "nplanes": img.nplanes}
        blocks = sum(((img.plane[i].width >> 3) << img.plane[i].xdec) * img.plane[i].cstride
                     for i in range(img.nplanes))
        if out in ("quant", "dct"):
            res["coef"] = np.ctypeslib.as_array(img.coef, shape=(blocks * 64,)).copy()
        elif out == "pack":
            packed = [img.plane[i].packed for i in range(img.nplanes)]
            res["packed"] = packed
            res["pack"] = np.ctypeslib.as_array(img.coef, shape=(blocks * 64,))[:sum(packed)].copy()
            res["index"] = np.ctypeslib.as_array(img.index, shape=(blocks,)).copy()
        elif out == "yuv":
            res["planes"] = [np.ctypeslib.as_array(img.plane[i].data,
                                                   shape=(img.plane[i].height, img.plane[i].width)).copy()
                             for i in range(img.nplanes)]
        elif out == "rgb":
            ch = 1 if img.nplanes == 1 else 3
            px = np.ctypeslib.as_array(img.pixels, shape=(img.height * img.width * ch,)).copy()
            res["pixels"] = px.reshape(img.height, img.width) if ch == 1 else px.reshape(img.height, img.width, 3)
        return res

    def decode_reset(self, data: Optional[bytes] = None) -> None:
        if data is not None:
            self._buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
            self._info = _capi.jpeg_info(len(data), C.cast(self._buf, C.POINTER(C.c_ubyte)))
            if self._img is not None:
                _capi.lib().jgpu_image_clear(C.byref(self._img))
                self._img = None
        self._vt.decode_reset(self._dec, C.byref(self._info))
        self._have_header = False

    def close(self) -> None:
        if self._dec:
            self._vt.decode_free(self._dec)
            self._dec = None
        if self._img is not None:
            _capi.lib().jgpu_image_clear(C.byref(self._img))
            self._img = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
