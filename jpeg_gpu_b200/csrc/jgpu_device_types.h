/* Device-side work descriptors shared by the kernels and the runtime that
 * builds them.  Plain PODs; built on the host by jgpu_plan_create and read by
 * the kernels from global memory. */
#ifndef JGPU_DEVICE_TYPES_H
#define JGPU_DEVICE_TYPES_H

#include <stdint.h>

namespace jgpu {

/* ---- generic path -------------------------------------------------------- */

/* One colour plane of one image: a run of nblocks coefficient blocks in raster
 * order (the reference layout is block-linear inside a plane, src/xjpeg.c:
 * 558-562) and the padded u8 plane they decode into. */
struct PlaneSeg {
  int64_t coef_off;   /* first int16 of the plane in the coef buffer */
  int64_t out_off;    /* first byte of the plane in the planes buffer */
  int32_t hblocks;    /* blocks per block row */
  int32_t nblocks;    /* hblocks*vblocks */
  int32_t pitch;      /* plane row stride in bytes = hblocks*8 */
  int32_t qidx;       /* 64-entry table index: qtab_set*4 + tq */
};

/* One CTA of k_coef_to_planes: `first` is the first block PAIR it owns. */
struct PairWork {
  int32_t seg;
  int32_t first;
};

/* One image for k_planes_to_rgb. */
struct ColourImage {
  int64_t plane_off[3]; /* byte offsets of Y, Cb, Cr in the planes buffer */
  int64_t rgb_off;      /* byte offset of the output in the rgb buffer */
  int32_t width, height;
  int32_t pitch[3];
  int32_t xdec[3], ydec[3];
  int32_t ncomps;
  int32_t groups_per_row; /* ceil(width/4) */
  int32_t reserved;
};

/* One CTA of k_planes_to_rgb: `first` is the first 4-pixel group it owns. */
struct ColourWork {
  int32_t img;
  int32_t first;
};

/* ---- PACK expansion (jgpu_unpack.cu) --------------------------------------- */

/* One colour plane of one image: its blocks are consecutive 128-byte rows of the
 * coefficient buffer starting at row block0, and index entry block0 + k belongs to
 * block k (src/image.c:85-95 gives the index the layout of the coefficients). */
struct UnpackSeg {
  int64_t block0;   /* (coef_off + plane.coef_off) / 64 */
  int32_t nblocks;  /* hblocks*vblocks */
  int32_t img;      /* selects pack_off[img], pack_off[img+1] */
};

/* One CTA of k_unpack: blocks [first, first + 256) of a segment. */
struct UnpackWork {
  int32_t seg;
  int32_t first;
};

/* ---- fused path ---------------------------------------------------------- */

/* Sampling classes the fused kernel is instantiated for (chroma 1x1). */
enum FusedMode : int32_t {
  kModeGray = 0, /* 1 component                 MCU  8x8,  1 block  */
  kMode444 = 1,  /* luma 1x1                    MCU  8x8,  3 blocks */
  kMode422 = 2,  /* luma 2x1                    MCU 16x8,  4 blocks */
  kMode420 = 3,  /* luma 2x2                    MCU 16x16, 6 blocks */
  kMode440 = 4,  /* luma 1x2                    MCU  8x16, 4 blocks */
  kMode411 = 5,  /* luma 4x1                    MCU 32x8,  6 blocks (k_tk only) */
  kNumFusedModes = 6
};

/* One image for the fused kernel.  block0[] are GLOBAL block indices: the
 * coefficient buffer viewed as rows of 64 int16 (128 bytes). */
struct FusedImage {
  int64_t rgb_off;
  int32_t block0[3];
  int32_t hblocks[3];
  int32_t width, height;
  int32_t qidx[3];     /* 64-entry table index: qtab_set*4 + tq */
  int32_t reserved;
};

/* One tile: a run of MCUs in one MCU row of one image. */
struct TileRef {
  int32_t img;
  int16_t mrow; /* MCU row */
  int16_t mx0;  /* first MCU (units of the mode's tile width, see jgpu_fused.cu) */
};

}  // namespace jgpu
#endif
