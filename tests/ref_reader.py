"""The reference's OWN decoder table and surface functions, taken from oracle/_ref/libjgpu_ref.so
(the reference's src/jpeg_wrap.c, src/xjpeg.c, src/image.c compiled unmodified by oracle/Makefile),
and a driver that walks any five-slot table through the reference's call protocol
(src/jpeg_gpu.c:612-704 first frame, :1231-1237 steady state).

Test infrastructure: used to plug XJPEG_DECODE_CTX_VTBL (src/jpeg_wrap.c:352-358) into the CUDA
backend with cuda_decode_set_frontend(), which is the binding INTEGRATION.md section 2 documents.
"""
import ctypes as C
import hashlib
import os

import numpy as np

import oracle
from jpeg_gpu_b200 import _capi

BIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "big")
BIG_NAMES = ["gray_512x512", "c420_1920x1080"]


def load_big(name):
    with open(os.path.join(BIG, name + ".jpg"), "rb") as f:
        jpg = f.read()
    return jpg, np.load(os.path.join(BIG, name + ".npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class RefLib:
    """ctypes view of the compiled reference: its two tables and image_init/zero/clear."""

    def __init__(self):
        self.lib = C.CDLL(oracle.REF_SO)
        self.lib.image_init.restype = C.c_int
        self.lib.image_init.argtypes = [C.POINTER(_capi.image), C.POINTER(_capi.jpeg_header)]
        self.lib.image_zero.restype = None
        self.lib.image_zero.argtypes = [C.POINTER(_capi.image)]
        self.lib.image_clear.restype = None
        self.lib.image_clear.argtypes = [C.POINTER(_capi.image)]
        self.xjpeg = _capi.jpeg_decode_ctx_vtbl.in_dll(self.lib, "XJPEG_DECODE_CTX_VTBL")

    @property
    def xjpeg_address(self):
        return C.addressof(self.xjpeg)


class Session:
    """One decoder context of table `vt`, surface from `image_init`/`image_clear` (the reference's
    own, or the product's jgpu_image_*): alloc -> header -> image_init -> image ..., reset, free."""

    def __init__(self, vt, jpg, image_init, image_zero, image_clear):
        self.vt, self._init, self._zero, self._clear = vt, image_init, image_zero, image_clear
        self.buf = (C.c_ubyte * len(jpg)).from_buffer_copy(jpg)
        self.info = _capi.jpeg_info(len(jpg), C.cast(self.buf, C.POINTER(C.c_ubyte)))
        self.hdr = _capi.jpeg_header()
        self.img = None
        self.dec = vt.decode_alloc(C.byref(self.info))
        assert self.dec, "decode_alloc returned NULL"

    def header(self):
        rc = self.vt.decode_header(self.dec, C.byref(self.hdr))
        if rc == 0 and self.img is None:
            self.img = _capi.image()
            assert self._init(C.byref(self.img), C.byref(self.hdr)) == 0
            self._zero(C.byref(self.img))
        return rc

    def image(self, out):
        return self.vt.decode_image(self.dec, C.byref(self.img), _capi.OUT_NAMES[out])

    def reset(self):
        self.vt.decode_reset(self.dec, C.byref(self.info))

    # -- what the surface holds -------------------------------------------------
    def blocks(self):
        img = self.img
        return sum(((img.plane[i].width >> 3) << img.plane[i].xdec) * img.plane[i].cstride for i in range(img.nplanes))

    def planes(self):
        img = self.img
        return np.concatenate([np.ctypeslib.as_array(img.plane[i].data, shape=(img.plane[i].height * img.plane[i].width,))
                               for i in range(img.nplanes)]).copy()

    def pixels(self):
        img = self.img
        ch = 1 if img.nplanes == 1 else 3
        return np.ctypeslib.as_array(img.pixels, shape=(img.height * img.width * ch,)).copy()

    def coef(self):
        return np.ctypeslib.as_array(self.img.coef, shape=(self.blocks() * 64,)).copy()

    def close(self):
        if self.dec:
            self.vt.decode_free(self.dec)
            self.dec = None
        if self.img is not None:
            self._clear(C.byref(self.img))
            self.img = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
