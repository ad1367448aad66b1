/* jgpu_kernels.cu — sm_100a kernels of the coefficient -> RGB path.
 *
 * Generic path (any sampling the reference layout supports, any size):
 *   k_coef_to_planes   dequantise + 8x8 IDCT + bias/clamp -> padded u8 planes
 *                      (replaces res/horz*.fs.glsl + res/vert.fs.glsl; output
 *                      is what XJPEG's JPEG_DECODE_YUV writes, src/xjpeg.c:565-584)
 *   k_planes_to_rgb    nearest chroma upsample + colour matrix + crop
 *                      (replaces res/unyuv.fs.glsl / ungrey.fs.glsl / yuv.fs.glsl)
 * Fused path: jgpu_fused.cu.
 */
#include "jgpu_kernels.cuh"
#include "jgpu_launch.h"

namespace jgpu {

/* One thread = two consecutive blocks of one plane (256 contiguous bytes of
 * coefficients); all 64 sample pairs stay in registers through both passes. */
__global__ void __launch_bounds__(kPairThreads)
k_coef_to_planes(const PlaneSeg *__restrict__ segs,
                 const PairWork *__restrict__ work,
                 const int16_t *__restrict__ coef,
                 const uint16_t *__restrict__ qtabs,
                 uint8_t *__restrict__ planes) {
  __shared__ int sq[64];
  const PairWork w = work[blockIdx.x];
  const PlaneSeg s = segs[w.seg];
  if (threadIdx.x < 64) sq[threadIdx.x] = qtabs[(int64_t)s.qidx * 64 + threadIdx.x];
  __syncthreads();

  const int b0 = (w.first + (int)threadIdx.x) * 2;
  if (b0 >= s.nblocks) return;
  const bool two = b0 + 1 < s.nblocks;
  const uint4 *src = reinterpret_cast<const uint4 *>(coef + s.coef_off + (int64_t)b0 * 64);

  pair32 m[8][8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    uint4 a = __ldg(src + r);
    uint4 b = two ? __ldg(src + 8 + r) : make_uint4(0, 0, 0, 0);
    load_row_pair(m[r], a, b, sq + 8 * r, sq + 8 * r, r);
    inv_pass8(m[r]);
  }
  column_pass(m);

  const int by0 = b0 / s.hblocks, bx0 = b0 - by0 * s.hblocks;
  const int b1 = b0 + 1;
  const int by1 = b1 / s.hblocks, bx1 = b1 - by1 * s.hblocks;
  uint8_t *dst0 = planes + s.out_off + (int64_t)by0 * 8 * s.pitch + bx0 * 8;
  uint8_t *dst1 = planes + s.out_off + (int64_t)by1 * 8 * s.pitch + bx1 * 8;
  const pair32 magic = p_make_bits(kMagicBits, kMagicBits);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int c = 0; c < 8; c++) p_split_bits(p_add_rm(m[k][c], magic), lo[c], hi[c]);
    uint2 ra = pack_row_u8(clamp_pair_u8(lo[0], lo[1]), clamp_pair_u8(lo[2], lo[3]),
                           clamp_pair_u8(lo[4], lo[5]), clamp_pair_u8(lo[6], lo[7]));
    *reinterpret_cast<uint2 *>(dst0 + (int64_t)k * s.pitch) = ra;
    if (two) {
      uint2 rb = pack_row_u8(clamp_pair_u8(hi[0], hi[1]), clamp_pair_u8(hi[2], hi[3]),
                             clamp_pair_u8(hi[4], hi[5]), clamp_pair_u8(hi[6], hi[7]));
      *reinterpret_cast<uint2 *>(dst1 + (int64_t)k * s.pitch) = rb;
    }
  }
}

/* One thread = four horizontally adjacent output pixels. */
__global__ void __launch_bounds__(kColourThreads)
k_planes_to_rgb(const ColourImage *__restrict__ imgs,
                const ColourWork *__restrict__ work,
                const uint8_t *__restrict__ planes, uint8_t *__restrict__ rgb) {
  const ColourWork w = work[blockIdx.x];
  const ColourImage im = imgs[w.img];
  const int item = w.first + (int)threadIdx.x;
  const int y = item / im.groups_per_row;
  if (y >= im.height) return;
  const int x0 = (item - y * im.groups_per_row) * 4;
  const int n = min(4, im.width - x0);
  const uint8_t *yrow = planes + im.plane_off[0] + (int64_t)y * im.pitch[0];
  const uint32_t y4 = *reinterpret_cast<const uint32_t *>(yrow + x0); /* pitch%8==0 */

  if (im.ncomps == 1) {
    uint8_t *o = rgb + im.rgb_off + (int64_t)y * im.width + x0;
    if (n == 4 && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {
      *reinterpret_cast<uint32_t *>(o) = y4;
    } else {
      for (int i = 0; i < n; i++) o[i] = (uint8_t)(y4 >> (8 * i));
    }
    return;
  }

  const uint8_t *brow = planes + im.plane_off[1] + (int64_t)(y >> im.ydec[1]) * im.pitch[1];
  const uint8_t *rrow = planes + im.plane_off[2] + (int64_t)(y >> im.ydec[2]) * im.pitch[2];
  uint8_t px[12];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    /* x0+i may run past width but never past the padded plane */
    const int x = x0 + i;
    int ro, go, bo;
    colour_offsets(brow[x >> im.xdec[1]], rrow[x >> im.xdec[2]], ro, go, bo);
    const int yy = (int)((y4 >> (8 * i)) & 0xffu);
    px[3 * i + 0] = (uint8_t)clamp255(yy + ro);
    px[3 * i + 1] = (uint8_t)clamp255(yy + go);
    px[3 * i + 2] = (uint8_t)clamp255(yy + bo);
  }
  uint8_t *o = rgb + im.rgb_off + ((int64_t)y * im.width + x0) * 3;
  if (n == 4 && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {
    uint32_t *o32 = reinterpret_cast<uint32_t *>(o);
    o32[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
    o32[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
    o32[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
  } else {
    for (int i = 0; i < 3 * n; i++) o[i] = px[i];
  }
}

cudaError_t launch_coef_to_planes(const PlaneSeg *segs, const PairWork *work,
                                  int ncta, const int16_t *coef,
                                  const uint16_t *qtabs, uint8_t *planes,
                                  cudaStream_t stream) {
  if (ncta <= 0) return cudaSuccess;
  k_coef_to_planes<<<ncta, kPairThreads, 0, stream>>>(segs, work, coef, qtabs, planes);
  return cudaGetLastError();
}

cudaError_t launch_planes_to_rgb(const ColourImage *imgs, const ColourWork *work,
                                 int ncta, const uint8_t *planes, uint8_t *rgb,
                                 cudaStream_t stream) {
  if (ncta <= 0) return cudaSuccess;
  k_planes_to_rgb<<<ncta, kColourThreads, 0, stream>>>(imgs, work, planes, rgb);
  return cudaGetLastError();
}

}  // namespace jgpu
