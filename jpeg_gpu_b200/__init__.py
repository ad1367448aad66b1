"""jpeg_gpu_b200 — B200-native JPEG block-decode back end (coefficients -> RGB)
behind the decoder-plugin interface of negge/jpeg_gpu.

  Decoder        the reference's five-slot backend protocol (src/jpeg_wrap.h)
  Context/Plan   batch API over device-resident or host buffers
  ImageDesc      one image of a batch; layout as the reference's image_init
  synth          synthetic coefficient planes
  shard          multi-GPU sharding helpers

All compute happens in libjpeg_gpu_b200.so (hand-written sm_100a CUDA behind a
C ABI, include/jpeg_gpu_b200.h).  There is no CPU fallback.
"""
from . import _capi, shard, synth
from ._capi import LibraryMissing
from .batch import (Context, ImageDesc, JpegInfo, Layout, Plan, PlaneLayout, SUBSAMPLINGS, pack_batch,
                    pack_batch_streams, pack_from_quant, probe_jpegs)
from .decoder import DecodeError, Decoder, Header

__all__ = ["Decoder", "DecodeError", "Header", "Context", "Plan", "ImageDesc", "Layout", "PlaneLayout",
           "SUBSAMPLINGS", "pack_batch", "pack_batch_streams", "pack_from_quant", "probe_jpegs", "JpegInfo", "synth", "shard", "LibraryMissing", "_capi"]
