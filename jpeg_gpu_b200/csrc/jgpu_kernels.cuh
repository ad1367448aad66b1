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

/* ---- packed-table variant (fused kernel) -------------------------------------
 * The product of one int16 coefficient with one 8-bit table entry is a single
 * IDP.2A: dp2a treats `w` as two int16 and `q` as four bytes,
 *     lo: w.h0*q.b0 + w.h1*q.b1        hi: w.h0*q.b2 + w.h1*q.b3
 * so with q = (q_even, 0, 0, q_odd) the two forms pick and multiply one half
 * each with no unpack instruction.  Table entries above 255 (16-bit DQT) are
 * split: c*q = c*(q&255) + ((c*(q>>8)) << 8), all modulo 2^16, which is the
 * reference's wrap into `short` anyway (src/xjpeg.c:501-503,524-527). */
JGPU_DEV int dp2a_lo_s16_u8(uint32_t w, uint32_t q) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(q), "r"(0));
  return d;
}
JGPU_DEV int dp2a_hi_s16_u8(uint32_t w, uint32_t q) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(q), "r"(0));
  return d;
}

/* A/B knob (measured, profiles/r1_ab_notes.md): (float)(short)a, (float)(short)b as one packed
 * pair without the conversion unit.  I2F.S16 issues on the XU pipe (one warp instruction per 8
 * cycles per sub-core; 32 % XU utilisation with 128 conversions per thread).  The alternative:
 * the low 16 bits of the product, sign bit flipped, under the exponent of 2^23 are the binary32
 * number 2^23 + 32768 + (short)a (ulp is 1 in that binade), and subtracting 2^23 + 32768 is
 * exact: one LOP3 per value plus one packed FADD2 per pair, same binary32 result.  It removes
 * the XU traffic but adds 64 issue slots per thread, and the kernel is issue-bound: 1.8 %
 * SLOWER on B200 (3.100 vs 3.045 ms, 4K 4:2:0 batch 256), so the default stays I2F. */
#ifndef JGPU_DEQ_MAGIC
#define JGPU_DEQ_MAGIC 0
#endif
JGPU_DEV pair32 s16_pair_to_float(int a, int b) {
#ifdef JGPU_CORE_HOST_EMULATION
  const uint32_t ua = ((uint32_t)a & 0xffffu) ^ 0x4b008000u;
  const uint32_t ub = ((uint32_t)b & 0xffffu) ^ 0x4b008000u;
#else
  /* (x & 0xffff) ^ k as ONE LOP3 (lut 0x6a = (a & b) ^ c); written in C, ptxas emits two */
  uint32_t ua, ub;
  const uint32_t k = 0x4b008000u;
  asm("lop3.b32 %0, %1, 0xffff, %2, 0x6a;" : "=r"(ua) : "r"(a), "r"(k));
  asm("lop3.b32 %0, %1, 0xffff, %2, 0x6a;" : "=r"(ub) : "r"(b), "r"(k));
#endif
  return p_sub(p_make_bits(ua, ub), p_make_bits(0x4b008000u, 0x4b008000u));
}

/* One coefficient row of block A and of block B (8 int16 each) with the packed
 * table rows qa / qb (4 words each; WIDE: plus the high-byte rows qah / qbh):
 * dequantise, convert, prescale. */
template <bool WIDE>
JGPU_DEV void load_row_pair_packed(pair32 (&row)[8], uint4 a, uint4 b, uint4 qa, uint4 qb,
                                   uint4 qah, uint4 qbh, int r) {
  const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
  const uint32_t ka[4] = {qa.x, qa.y, qa.z, qa.w}, kb[4] = {qb.x, qb.y, qb.z, qb.w};
  const uint32_t ha[4] = {qah.x, qah.y, qah.z, qah.w}, hb[4] = {qbh.x, qbh.y, qbh.z, qbh.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int a0 = dp2a_lo_s16_u8(wa[i], ka[i]), a1 = dp2a_hi_s16_u8(wa[i], ka[i]);
    int b0 = dp2a_lo_s16_u8(wb[i], kb[i]), b1 = dp2a_hi_s16_u8(wb[i], kb[i]);
    if (WIDE) {
      a0 += dp2a_lo_s16_u8(wa[i], ha[i]) << 8;
      a1 += dp2a_hi_s16_u8(wa[i], ha[i]) << 8;
      b0 += dp2a_lo_s16_u8(wb[i], hb[i]) << 8;
      b1 += dp2a_hi_s16_u8(wb[i], hb[i]) << 8;
    }
#if JGPU_DEQ_MAGIC == 1
    row[2 * i] = prescale(s16_pair_to_float(a0, b0), r, 2 * i);
    row[2 * i + 1] = prescale(s16_pair_to_float(a1, b1), r, 2 * i + 1);
#elif JGPU_DEQ_MAGIC == 2   /* every fourth pair without the conversion unit (A/B, profiles/r2_notes.md) */
    row[2 * i] = prescale((i & 1) ? s16_pair_to_float(a0, b0) : p_make((float)(short)a0, (float)(short)b0), r, 2 * i);
    row[2 * i + 1] = prescale(p_make((float)(short)a1, (float)(short)b1), r, 2 * i + 1);
#else
    row[2 * i] = prescale(p_make((float)(short)a0, (float)(short)b0), r, 2 * i);
    row[2 * i + 1] = prescale(p_make((float)(short)a1, (float)(short)b1), r, 2 * i + 1);
#endif
  }
}

/* The same for ONE block: rows r and r + 4 of it ride in the two lanes (k_gray_tpb), so the
 * first prescale factor differs between the lanes: (y*S[r])*S[c] and (y*S[r+4])*S[c]. */
template <bool WIDE>
JGPU_DEV void load_two_rows_packed(pair32 (&row)[8], uint4 a, uint4 b, uint4 qa, uint4 qb, uint4 qah, uint4 qbh, int r) {
  const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
  const uint32_t ka[4] = {qa.x, qa.y, qa.z, qa.w}, kb[4] = {qb.x, qb.y, qb.z, qb.w};
  const uint32_t ha[4] = {qah.x, qah.y, qah.z, qah.w}, hb[4] = {qbh.x, qbh.y, qbh.z, qbh.w};
  const pair32 sr = p_make_bits(scale_bits(r), scale_bits(r + 4));
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int a0 = dp2a_lo_s16_u8(wa[i], ka[i]), a1 = dp2a_hi_s16_u8(wa[i], ka[i]);
    int b0 = dp2a_lo_s16_u8(wb[i], kb[i]), b1 = dp2a_hi_s16_u8(wb[i], kb[i]);
    if (WIDE) {
      a0 += dp2a_lo_s16_u8(wa[i], ha[i]) << 8;
      a1 += dp2a_hi_s16_u8(wa[i], ha[i]) << 8;
      b0 += dp2a_lo_s16_u8(wb[i], hb[i]) << 8;
      b1 += dp2a_hi_s16_u8(wb[i], hb[i]) << 8;
    }
    row[2 * i] = p_mulc(p_mul(p_make((float)(short)a0, (float)(short)b0), sr), scale_bits(2 * i));
    row[2 * i + 1] = p_mulc(p_mul(p_make((float)(short)a1, (float)(short)b1), sr), scale_bits(2 * i + 1));
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
