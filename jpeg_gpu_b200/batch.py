"""Batch API: device-resident (torch tensors) and host-buffer decoding.

Thin wrappers over jgpu_plan_* / jgpu_decode_batch_host
(include/jpeg_gpu_b200.h).  PyTorch is used only for device memory and
streams; all work happens in the C/CUDA library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _capi

SUBSAMPLINGS = {
    "gray": ((1,), (1,)),
    "444": ((1, 1, 1), (1, 1, 1)),
    "422": ((2, 1, 1), (1, 1, 1)),
    "420": ((2, 1, 1), (2, 1, 1)),
    "440": ((1, 1, 1), (2, 1, 1)),
    "411": ((4, 1, 1), (1, 1, 1)),
    "410": ((4, 1, 1), (2, 1, 1)),
}


@dataclass
class PlaneLayout:
    hblocks: int
    vblocks: int
    width: int
    height: int
    xdec: int
    ydec: int
    cstride: int
    coef_off: int
    data_off: int


@dataclass
class Layout:
    nhmb: int
    nvmb: int
    hmax: int
    vmax: int
    coef_len: int
    coded_blocks: int
    data_len: int
    rgb_len: int
    planes: List[PlaneLayout]


@dataclass
class ImageDesc:
    """One image of a batch (jgpu_image_desc)."""
    width: int
    height: int
    hsamp: Sequence[int]
    vsamp: Sequence[int]
    tq: Sequence[int] = (0, 1, 1)
    qtab_set: int = 0
    coef_off: int = 0
    rgb_off: int = 0
    yuv_off: int = -1
    layout: Optional[Layout] = field(default=None, compare=False)

    def to_c(self) -> _capi.jgpu_image_desc:
        d = _capi.jgpu_image_desc()
        d.width, d.height, d.ncomps = self.width, self.height, len(self.hsamp)
        for i in range(len(self.hsamp)):
            d.hsamp[i], d.vsamp[i], d.tq[i] = self.hsamp[i], self.vsamp[i], self.tq[i]
        d.qtab_set, d.coef_off, d.rgb_off, d.yuv_off = self.qtab_set, self.coef_off, self.rgb_off, self.yuv_off
        return d

    def query_layout(self) -> Layout:
        """jgpu_layout_query: host-only, needs no GPU."""
        if self.layout is None:
            out = _capi.jgpu_layout()
            d = self.to_c()
            if _capi.lib().jgpu_layout_query(C.byref(d), C.byref(out)) != 0:
                raise ValueError(_capi.last_error())
            planes = [PlaneLayout(p.hblocks, p.vblocks, p.width, p.height, p.xdec, p.ydec, p.cstride,
                                  p.coef_off, p.data_off) for p in out.plane[:len(self.hsamp)]]
            self.layout = Layout(out.nhmb, out.nvmb, out.hmax, out.vmax, out.coef_len, out.coded_blocks,
                                 out.data_len, out.rgb_len, planes)
        return self.layout


def pack_batch(descs: List[ImageDesc], want_yuv: bool = False, align: int = 256):
    """Assigns back-to-back coef / rgb / yuv offsets.  Returns the buffer sizes
    (coef in int16 elements, rgb and yuv in bytes)."""
    coef = rgb = yuv = 0
    for d in descs:
        lay = d.query_layout()
        d.coef_off, d.rgb_off = coef, rgb
        coef += -(-lay.coef_len // 8) * 8
        rgb += -(-lay.rgb_len // align) * align
        if want_yuv:
            d.yuv_off = yuv
            yuv += -(-lay.data_len // align) * align
        else:
            d.yuv_off = -1
    return coef, rgb, yuv


def pack_from_quant(desc: ImageDesc, coef: np.ndarray):
    """Dense QUANT planes of one image (its image.coef) -> (pack uint16 words, index int32) as the
    reference's reader would write them for JPEG_DECODE_PACK (jgpu_pack_from_quant; host only)."""
    lay = desc.query_layout()
    coef = np.ascontiguousarray(coef, dtype=np.int16)
    assert coef.size >= lay.coef_len
    d = desc.to_c()
    cap = _capi.lib().jgpu_pack_bound(C.byref(d))
    pack = np.empty(cap, dtype=np.uint16)
    index = np.empty(lay.coef_len // 64, dtype=np.int32)
    n = _capi.lib().jgpu_pack_from_quant(C.byref(d), _addr(coef), _addr(pack), cap, _addr(index))
    if n < 0:
        raise RuntimeError(f"jgpu_pack_from_quant failed: {_capi.last_error()}")
    return pack[:n].copy(), index


def pack_batch_streams(descs: List[ImageDesc], coef: np.ndarray):
    """PACK form of a whole batch laid out by pack_batch(): returns (pack, pack_off int64[n+1],
    index int32[coef_len/64]) for jgpu_plan_unpack / jgpu_decode_batch_host_packed."""
    packs, offs = [], [0]
    total = max(d.coef_off + d.query_layout().coef_len for d in descs)
    index = np.zeros(-(-total // 64), dtype=np.int32)
    for d in descs:
        assert d.coef_off % 64 == 0, "PACK input needs coef_off to be a multiple of 64"
        lay = d.query_layout()
        p, ix = pack_from_quant(d, coef[d.coef_off:d.coef_off + lay.coef_len])
        packs.append(p)
        offs.append(offs[-1] + p.size)
        index[d.coef_off // 64:d.coef_off // 64 + ix.size] = ix
    return np.concatenate(packs), np.array(offs, dtype=np.int64), index


@dataclass
class JpegInfo:
    """jgpu_jpeg_info: what jgpu_jpegs_probe / jgpu_decode_jpegs report per file."""
    status: int
    width: int
    height: int
    ncomps: int
    hsamp0: int
    vsamp0: int
    restart_interval: int
    tasks: int
    rgb_off: int
    rgb_len: int
    message: Optional[str]

    @property
    def shape(self):
        return (self.height, self.width) if self.ncomps == 1 else (self.height, self.width, 3)


_JPEG_DT = np.dtype([("data", "<u8"), ("size", "<i8")])
_INFO_DT = np.dtype([("status", "<i4"), ("width", "<i4"), ("height", "<i4"), ("ncomps", "<i4"), ("hsamp0", "<i4"),
                     ("vsamp0", "<i4"), ("restart_interval", "<i4"), ("tasks", "<i4"), ("rgb_off", "<i8"),
                     ("rgb_len", "<i8"), ("message", "<u8")])
assert _JPEG_DT.itemsize == C.sizeof(_capi.jgpu_jpeg) and _INFO_DT.itemsize == C.sizeof(_capi.jgpu_jpeg_info)


def _jpeg_array(files: Sequence[bytes]):
    """The jgpu_jpeg array of a batch.  No copy: the C side reads the bytes objects' own buffers (kept
    alive by `keep`).  Filled column-wise -- a batch is hundreds of files and the call is milliseconds."""
    keep = [f if isinstance(f, bytes) else bytes(f) for f in files]
    n = len(keep)
    arr = (_capi.jgpu_jpeg * n)()
    if n:
        ptrs = (C.c_char_p * n)(*[f if f else b"\0" for f in keep])   # the buffers' addresses, taken in C
        view = np.frombuffer(arr, dtype=_JPEG_DT)
        view["data"] = np.frombuffer(ptrs, dtype=np.uint64)
        view["size"] = np.fromiter(map(len, keep), dtype=np.int64, count=n)
    return arr, keep


def _jpeg_infos(raw) -> List[JpegInfo]:
    if len(raw) == 0:
        return []
    v = np.frombuffer(raw, dtype=_INFO_DT)
    msgs = [C.cast(int(m), C.c_char_p).value.decode() if m else None for m in v["message"].tolist()] \
        if v["message"].any() else [None] * len(v)
    cols = [v[k].tolist() for k in ("status", "width", "height", "ncomps", "hsamp0", "vsamp0", "restart_interval",
                                   "tasks", "rgb_off", "rgb_len")]
    return [JpegInfo(*row, msg) for row, msg in zip(zip(*cols), msgs)]


def probe_jpegs(files: Sequence[bytes], out: str = "rgb"):
    """Headers only (host, no GPU): returns (output bytes needed, [JpegInfo]); out="yuv" sizes the
    buffer for the padded Y|Cb|Cr planes (JGPU_JPEGS_OUT_YUV) instead of pixels."""
    arr, keep = _jpeg_array(files)
    raw = (_capi.jgpu_jpeg_info * len(files))()
    total = _capi.lib().jgpu_jpegs_probe_ex(arr, len(files), {"rgb": 0, "yuv": 0x200}[out], raw)
    if total < 0:
        raise RuntimeError(f"jgpu_jpegs_probe failed: {_capi.last_error()}")
    del keep
    return int(total), _jpeg_infos(raw)


def _desc_array(descs: List[ImageDesc]):
    arr = (_capi.jgpu_image_desc * len(descs))()
    for i, d in enumerate(descs):
        arr[i] = d.to_c()
    return arr


class Context:
    """jgpu_ctx: one per device per host thread."""

    def __init__(self, device: int = 0):
        self._h = _capi.lib().jgpu_create(device)
        if not self._h:
            raise RuntimeError(f"jgpu_create({device}) failed: {_capi.last_error()}")
        self.device = device

    def close(self):
        if self._h:
            _capi.lib().jgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- device-resident ----------------------------------------------------------
    def plan(self, descs: List[ImageDesc], rgb: bool = True, yuv: bool = False, force_generic: bool = False):
        return Plan(self, descs, rgb, yuv, force_generic)

    # -- host buffers ---------------------------------------------------------------
    def decode_batch_host(self, descs: List[ImageDesc], coef: np.ndarray, qtabs: np.ndarray,
                          rgb: Optional[np.ndarray] = None, yuv: Optional[np.ndarray] = None,
                          force_generic: bool = False) -> None:
        """coef int16, qtabs uint16 (n_sets,4,64), rgb/yuv uint8: numpy arrays or
        anything exposing .ctypes / data_ptr (e.g. pinned torch tensors)."""
        flags = (_capi.JGPU_OUT_RGB if rgb is not None else 0) | (_capi.JGPU_OUT_YUV if yuv is not None else 0)
        if force_generic:
            flags |= _capi.JGPU_FORCE_GENERIC
        n_sets = int(np.prod(qtabs.shape)) // 256
        rc = _capi.lib().jgpu_decode_batch_host(self._h, _desc_array(descs), len(descs), flags, _addr(coef),
                                                _addr(qtabs), n_sets, _addr(rgb), _addr(yuv))
        if rc != 0:
            raise RuntimeError(f"jgpu_decode_batch_host failed: {_capi.last_error()}")


    ENTROPY = {"auto": 0, "cpu": 1, "gpu": 2}

    def decode_jpegs(self, files: Sequence[bytes], rgb=None, nthreads: int = 0, strict: bool = True,
                     entropy: str = "auto", out: str = "rgb"):
        """JPEG files in, RGB out (jgpu_decode_jpegs_ex).  entropy="gpu": Huffman decoding on the
        device (jgpu_huff.cu); "cpu": the multi-threaded host reader; "auto": $JGPU_ENTROPY, default
        gpu.  Returns (rgb uint8 buffer, [JpegInfo]); image i is
        rgb[info.rgb_off : info.rgb_off + info.rgb_len].reshape(info.shape)."""
        arr, keep = _jpeg_array(files)
        raw = (_capi.jgpu_jpeg_info * len(files))()
        out_flag = {"rgb": 0, "yuv": 0x200}[out]   # JGPU_JPEGS_OUT_YUV: padded Y|Cb|Cr planes per file
        if rgb is None:
            total = _capi.lib().jgpu_jpegs_probe_ex(arr, len(files), out_flag, raw)
            if total < 0:
                raise RuntimeError(f"jgpu_jpegs_probe failed: {_capi.last_error()}")
            rgb = np.zeros(max(int(total), 1), dtype=np.uint8)
        cap = rgb.numel() if hasattr(rgb, "numel") else rgb.size
        flags = self.ENTROPY[entropy] | out_flag
        if getattr(rgb, "is_cuda", False):
            flags |= 0x100   # JGPU_JPEGS_DEVICE_OUT: the pixels stay on the device
        rc = _capi.lib().jgpu_decode_jpegs_ex(self._h, arr, len(files), nthreads, flags, _addr(rgb), cap, raw)
        infos = _jpeg_infos(raw)
        del keep
        if rc != 0 and (strict or all(i.status for i in infos)):
            raise RuntimeError(f"jgpu_decode_jpegs failed: {_capi.last_error()}")
        return rgb, infos

    def decode_batch_host_packed(self, descs: List[ImageDesc], pack, pack_off: np.ndarray, index,
                                 qtabs: np.ndarray, rgb=None, yuv=None, force_generic: bool = False) -> None:
        """PACK words (uint16) + pack_off (int64, n+1) + index (int32) in, rgb / yuv out; host buffers."""
        flags = (_capi.JGPU_OUT_RGB if rgb is not None else 0) | (_capi.JGPU_OUT_YUV if yuv is not None else 0)
        if force_generic:
            flags |= _capi.JGPU_FORCE_GENERIC
        n_sets = int(np.prod(qtabs.shape)) // 256
        pack_off = np.ascontiguousarray(pack_off, dtype=np.int64)
        assert pack_off.size == len(descs) + 1
        rc = _capi.lib().jgpu_decode_batch_host_packed(self._h, _desc_array(descs), len(descs), flags, _addr(pack),
                                                       _addr(pack_off), _addr(index), _addr(qtabs), n_sets,
                                                       _addr(rgb), _addr(yuv))
        if rc != 0:
            raise RuntimeError(f"jgpu_decode_batch_host_packed failed: {_capi.last_error()}")


def _addr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


class Plan:
    """jgpu_plan: device work lists for one batch shape; run() only launches."""

    def __init__(self, ctx: Context, descs: List[ImageDesc], rgb=True, yuv=False, force_generic=False):
        self.ctx = ctx
        self.descs = descs
        self.flags = (_capi.JGPU_OUT_RGB if rgb else 0) | (_capi.JGPU_OUT_YUV if yuv else 0) | \
                     (_capi.JGPU_FORCE_GENERIC if force_generic else 0)
        self._h = _capi.lib().jgpu_plan_create(ctx._h, _desc_array(descs), len(descs), self.flags)
        if not self._h:
            raise RuntimeError(f"jgpu_plan_create failed: {_capi.last_error()}")
        self.launches = _capi.lib().jgpu_plan_launches(self._h)
        self.bytes = _capi.lib().jgpu_plan_bytes(self._h)

    def run(self, coef, qtabs, rgb=None, yuv=None, stream: Optional[int] = None) -> None:
        """coef/qtabs/rgb/yuv: CUDA torch tensors (int16 / uint16-as-int16 / uint8).
        stream: raw cudaStream_t; default = torch's current stream."""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(coef.device).cuda_stream
        n_sets = qtabs.numel() // 256
        rc = _capi.lib().jgpu_plan_run(self._h, _addr(coef), _addr(qtabs), n_sets, _addr(rgb), _addr(yuv),
                                       C.c_void_p(stream))
        if rc != 0:
            raise RuntimeError(f"jgpu_plan_run failed: {_capi.last_error()}")

    def unpack(self, pack, pack_off, index, coef, stream: Optional[int] = None) -> None:
        """PACK words / pack_off (int64, n+1) / index (int32) -> dense planes `coef`; CUDA torch tensors."""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(coef.device).cuda_stream
        rc = _capi.lib().jgpu_plan_unpack(self._h, _addr(pack), _addr(pack_off), _addr(index), _addr(coef),
                                          C.c_void_p(stream))
        if rc != 0:
            raise RuntimeError(f"jgpu_plan_unpack failed: {_capi.last_error()}")

    def close(self):
        if self._h:
            _capi.lib().jgpu_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
