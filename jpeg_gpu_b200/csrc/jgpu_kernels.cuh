/* jgpu_kernels.cuh — device helpers shared by the generic and fused kernels:
 * coefficient unpack + dequantise, float -> clamped sample conversion, colour
 * offsets.  Everything here is bit-defined; see jgpu_idct_core.cuh for the
 * transform itself. */
#ifndef JGPU_KERNELS_CUH
#define JGPU_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include "jgpu_idct_core.cuh"

namespace jgpu {

/* 1.5 * 2^23: adding it to |x| < 2^22 leaves the integer part of x in the low
 * mantissa bits; with round-toward-minus-infinity that integer is floor(x). */
static constexpr uint32_t kMagicBits = 0x4b400000u;

/* Dequantise one coefficient the way the reference does (src/xjpeg.c:501-503,
 * 524-527): int product, stored into a `short` (wraps modulo 2^16), then the
 * int16 -> binary32 conversion of src/dct.c:107 (exact). */
JGPU_DEV float dequant_to_float(int coef, int q) {
  return (float)(short)(coef * q);
}

/* Low / high int16 of a packed coefficient word, sign-extended. */
JGPU_DEV int coef_lo(uint32_t w) { return (int)(short)(w & 0xffffu); }
JGPU_DEV int coef_hi(uint32_t w) { return ((int)w) >> 16; }

/* Loads one coefficient row (8 int16 = 16 bytes) of block A and of block B,
 * dequantises with the per-column table entries qa[c] / qb[c], applies the
 * two-step prescale and leaves the packed pairs in row[0..7]. */
JGPU_DEV void load_row_pair(pair32 (&row)[8], uint4 a, uint4 b, const int *qa,
                            const int *qb, int r) {
  const uint32_t wa[4] = {a.x, a.y, a.z, a.w};
  const uint32_t wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int c = 0; c < 8; c++) {
    int ca = (c & 1) ? coef_hi(wa[c >> 1]) : coef_lo(wa[c >> 1]);
    int cb = (c & 1) ? coef_hi(wb[c >> 1]) : coef_lo(wb[c >> 1]);
    pair32 y = p_make(dequant_to_float(ca, qa[c]), dequant_to_float(cb, qb[c]));
    row[c] = prescale(y, r, c);
  }
}

/* Two un-floored samples (same block, adjacent columns) -> two clamped u8
 * samples in the halves of a 32-bit word:
 *   floor            src/dct.c:118 ((short)floor(t): low 16 bits of the integer)
 *   +128, clamp      src/xjpeg.c:578, src/internal.h:36-37
 * `bits0/bits1` are the raw bits of RM(t + 1.5*2^23). */
JGPU_DEV uint32_t clamp_pair_u8(uint32_t bits0, uint32_t bits1) {
  uint32_t s = __byte_perm(bits0, bits1, 0x5410);      /* (short)floor, x2 */
  s = __viaddmin_s16x2(s, 0u, 0x007f007fu);            /* min(v+0,127)     */
  return __viaddmax_s16x2(s, 0x00800080u, 0u);         /* max(v+128,0)     */
}

/* Four clamped sample words (8 samples as 16-bit halves) -> 8 bytes. */
JGPU_DEV uint2 pack_row_u8(uint32_t p01, uint32_t p23, uint32_t p45, uint32_t p67) {
  uint2 r;
  r.x = __byte_perm(p01, p23, 0x6420);
  r.y = __byte_perm(p45, p67, 0x6420);
  return r;
}

/* Colour offsets of one chroma sample pair; DEFINITION in
 * oracle/oracle_pipeline.c (jgo_colour_offsets), matrix from
 * res/yuv.fs.glsl:11-15.  Returns integer offsets for R, G, B. */
JGPU_DEV void colour_offsets(int cb, int cr, int &ro, int &go, int &bo) {
  const float magic = __uint_as_float(kMagicBits);
  float cbf = (float)(cb - 128);
  float crf = (float)(cr - 128);
  float rc = __fmul_rn(1.402f, crf);
  float gc = __fadd_rn(__fmul_rn(-0.34414f, cbf), __fmul_rn(-0.71414f, crf));
  float bc = __fmul_rn(1.772f, cbf);
  /* round to nearest integer, ties to even: RN(x + 1.5*2^23) */
  ro = (int)(short)(__float_as_uint(__fadd_rn(rc, magic)) & 0xffffu);
  go = (int)(short)(__float_as_uint(__fadd_rn(gc, magic)) & 0xffffu);
  bo = (int)(short)(__float_as_uint(__fadd_rn(bc, magic)) & 0xffffu);
}

JGPU_DEV int clamp255(int v) { return min(max(v, 0), 255); }

}  // namespace jgpu
#endif
