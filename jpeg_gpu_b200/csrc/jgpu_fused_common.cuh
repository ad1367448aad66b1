/* jgpu_fused_common.cuh -- device helpers shared by the fused kernels (jgpu_fused.cu: warp-role
 * tiles; jgpu_mcu.cu: one MCU column per thread): PTX wrappers for mbarrier / TMA / shared memory,
 * the row pass out of swizzled TMA boxes, chroma narrowing, colour offsets, RGB packing. */
#ifndef JGPU_FUSED_COMMON_CUH
#define JGPU_FUSED_COMMON_CUH

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "jgpu_colour_fixed.h"
#include "jgpu_kernels.cuh"

namespace jgpu {

#ifndef JGPU_STORE_POLICY
#define JGPU_STORE_POLICY 1     /* 0: L1::no_allocate, 1: .cs (streaming; 6 % faster on B200) */
#endif
#ifndef JGPU_COLOUR_PACKED
#define JGPU_COLOUR_PACKED 0    /* 1: colour offsets of two samples per packed instruction (measured 2.5 % slower) */
#endif
#ifndef JGPU_COLOUR_INT
#define JGPU_COLOUR_INT 1       /* 1: colour offsets in fixed point (jgpu_colour_fixed.h), no I2F / FMUL / FADD */
#endif

constexpr int kBoxRows = 32;                 /* blocks per TMA box */
constexpr int kBoxBytes = kBoxRows * 128;    /* 4 KB */
constexpr int kQtabBytes = 64 * 4;           /* one packed table: 32 low-byte + 32 high-byte words */

/* ---- PTX wrappers ---------------------------------------------------------- */

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
/* HINT_NS > 0: suspend-time hint.  ptxas lowers it to TRYWAIT; NANOSLEEP.SYNCS <hint>; re-check, and a
 * phase that completes between the check and the sleep is not seen until the sleep times out (measured:
 * warps of jgpu_mcu.cu that lost the race slept the full 100 us, several times per launch,
 * profiles/r2_notes.md).  HINT_NS == 0: the plain form, whose wait is bounded by the hardware. */
template <uint32_t HINT_NS>
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  if (HINT_NS > 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(HINT_NS)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  return mbar_try_wait_hint<100000u>(bar, parity);   /* sleep, do not spin (jgpu_fused.cu) */
}
/* Bounded wait: a pipeline bug must trap, not hang the GPU. */
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++) {
    if (spins > (1u << 18)) __trap();
  }
}
template <uint32_t HINT_NS>
__device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait_hint<HINT_NS>(bar, parity); spins++) {
    if (spins > (HINT_NS ? (1u << 22) : (1u << 23))) __trap();   /* (a few seconds: a protocol bug must trap, not hang the box) */
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1,
                                            int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes,
                                          uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void named_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int threads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
/* Output pixels are written once and never read: keep them out of L1 (which, next to 216 KB
 * of shared memory, is only ~28 KB and holds the few spilled registers). */
__device__ __forceinline__ void stg128_stream(uint8_t *p, uint4 v) {
#if JGPU_STORE_POLICY == 1
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
#else
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
#endif
}

/* ---- per-thread stages ----------------------------------------------------- */

/* Row pass for one block pair out of two swizzled TMA boxes.  `row` is this
 * lane's row inside the boxes; chunk r of a 128-byte row sits at 16*(r ^ (row&7)).
 * qa / qb: the packed quantisation tables of block A / block B (k_prep_qtabs):
 * 32 words of low bytes, then 32 words of high bytes. */
template <bool WIDE>
__device__ __forceinline__ void pair_row_pass(pair32 (&m)[8][8], const uint8_t *box_a,
                                              const uint8_t *box_b, int row, const uint4 *qa,
                                              const uint4 *qb, const uint32_t (&park)[4]) {
  const uint8_t *ra = box_a + 128 * row, *rb = box_b + 128 * row;
  const int sw = row & 7;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int off = 16 * (r ^ sw);
    const uint4 a = *reinterpret_cast<const uint4 *>(ra + off);
    const uint4 b = *reinterpret_cast<const uint4 *>(rb + off);
    const uint4 z = make_uint4(0, 0, 0, 0);
    load_row_pair_packed<WIDE>(m[r], a, b, qa[r], qb[r], WIDE ? qa[8 + r] : z, WIDE ? qb[8 + r] : z, r);
    inv_pass8(m[r]);
    if (r == 0) {
      /* park row 0 in shared memory until the column pass asks for it, two pairs per chunk */
#pragma unroll
      for (int j = 0; j < 4; j++) {
        uint4 c;
        p_split_bits(m[0][2 * j], c.x, c.y);
        p_split_bits(m[0][2 * j + 1], c.z, c.w);
        sts128(park[j], c);
      }
    }
  }
}

/* Chroma sample pair (Cb in .lo, Cr in .hi, un-floored) -> one s16x2 word
 * (Cb-128 | Cr-128 << 16) of the CLAMPED samples: (short)floor as in src/dct.c:118, then
 * clamp(v+128, 0, 255) - 128 == clamp(v, -128, 127) (src/xjpeg.c:578). */
__device__ __forceinline__ uint32_t chroma_clamped(pair32 v) {
  const pair32 magic = p_make_bits(kMagicBits, kMagicBits);
  uint32_t cbits, rbits;
  p_split_bits(p_add_rm(v, magic), cbits, rbits);
  uint32_t s = __byte_perm(cbits, rbits, 0x5410);
  s = __viaddmin_s16x2(s, 0u, 0x007f007fu);
  return __viaddmax_s16x2(s, 0u, 0xff80ff80u);
}

/* One exchange word = two chroma samples as signed bytes (Cb0-128, Cr0-128, Cb1-128, Cr1-128)
 * -> raw bits of RN(offset + 1.5*2^23) for R, G, B of both samples (low 16 bits = the integer
 * colour offset).  Arithmetic is colour_offsets() of jgpu_kernels.cuh, i.e. the oracle's
 * jgo_colour_offsets, with the two samples riding in the two lanes of the packed binary32
 * instructions (each lane is one IEEE operation, products via fma(a, b, -0.0)). */
/* Byte I of w, sign-extended: one PRMT whose selector nibbles 1..3 carry the replicate-sign bit
 * (PTX prmt default mode; __byte_perm documents only three selector bits, so spell it in PTX). */
template <int I>
__device__ __forceinline__ int sext_byte(uint32_t w) {
  int d;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(d) : "r"(w), "n"((8 + I) * 0x1110 + I));
  return d;
}

/* PRMT selectors that turn two offset words into one s16x2 operand: (x, x) and (x, y).  The
 * binary32 forms leave the offset in the low half of the word, the fixed-point form in the high. */
constexpr uint32_t kSelRep = JGPU_COLOUR_INT ? 0x3232u : 0x1010u;
constexpr uint32_t kSelPair = JGPU_COLOUR_INT ? 0x7632u : 0x5410u;

__device__ __forceinline__ void chroma_offsets_bits2(uint32_t w, uint32_t (&r)[2], uint32_t (&g)[2],
                                                     uint32_t (&b)[2]) {
#if JGPU_COLOUR_INT
  /* sign-extending byte extracts (PRMT with the replicate-sign selector bit), then
   * jgpu_colour_offsets_fixed: bit-identical to the binary32 definition for every input */
  const int cb0 = sext_byte<0>(w), cr0 = sext_byte<1>(w), cb1 = sext_byte<2>(w), cr1 = ((int)w) >> 24;
  int r0, g0, b0, r1, g1, b1;
  jgpu_colour_offsets_fixed(cb0, cr0, &r0, &g0, &b0);
  jgpu_colour_offsets_fixed(cb1, cr1, &r1, &g1, &b1);
  r[0] = (uint32_t)r0; g[0] = (uint32_t)g0; b[0] = (uint32_t)b0;
  r[1] = (uint32_t)r1; g[1] = (uint32_t)g1; b[1] = (uint32_t)b1;
  return;
#endif
#if JGPU_COLOUR_PACKED == 0
  /* scalar form of the same arithmetic (A/B reference) */
  const float fm = __uint_as_float(kMagicBits);
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const float cbf = (float)(signed char)((w >> (16 * i)) & 0xffu);
    const float crf = (float)(signed char)((w >> (16 * i + 8)) & 0xffu);
    r[i] = __float_as_uint(__fadd_rn(__fmul_rn(1.402f, crf), fm));
    g[i] = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(-0.34414f, cbf), __fmul_rn(-0.71414f, crf)), fm));
    b[i] = __float_as_uint(__fadd_rn(__fmul_rn(1.772f, cbf), fm));
  }
  return;
#endif
  const pair32 cb = p_make((float)(signed char)(w & 0xffu), (float)(signed char)((w >> 16) & 0xffu));
  const pair32 cr = p_make((float)(signed char)((w >> 8) & 0xffu), (float)(signed char)(w >> 24));
  const pair32 magic = p_make_bits(kMagicBits, kMagicBits);
  const pair32 rc = p_mulc(cr, 0x3fb374bcu);                                   /*  1.402    */
  const pair32 gc = p_add(p_mulc(cb, 0xbeb0331eu), p_mulc(cr, 0xbf36d1e1u));  /* -0.34414, -0.71414 */
  const pair32 bc = p_mulc(cb, 0x3fe2d0e5u);                                   /*  1.772    */
  p_split_bits(p_add(rc, magic), r[0], r[1]);
  p_split_bits(p_add(gc, magic), g[0], g[1]);
  p_split_bits(p_add(bc, magic), b[0], b[1]);
}

/* Four pixels: Y pairs (ya, yb) + colour-offset words -> 12 RGB bytes. */
__device__ __forceinline__ void rgb4(uint32_t ya, uint32_t yb, uint32_t ra, uint32_t ga,
                                     uint32_t ba, uint32_t rb, uint32_t gb, uint32_t bb,
                                     uint32_t &w0, uint32_t &w1, uint32_t &w2) {
  const uint32_t lim = 0x00ff00ffu;
  const uint32_t Ra = __viaddmin_s16x2_relu(ya, ra, lim);
  const uint32_t Ga = __viaddmin_s16x2_relu(ya, ga, lim);
  const uint32_t Ba = __viaddmin_s16x2_relu(ya, ba, lim);
  const uint32_t Rb = __viaddmin_s16x2_relu(yb, rb, lim);
  const uint32_t Gb = __viaddmin_s16x2_relu(yb, gb, lim);
  const uint32_t Bb = __viaddmin_s16x2_relu(yb, bb, lim);
  const uint32_t t = __byte_perm(Ra, Ga, 0x6240);   /* R0 G0 R1 G1 */
  const uint32_t u = __byte_perm(Rb, Gb, 0x6240);   /* R2 G2 R3 G3 */
  const uint32_t x = __byte_perm(t, Ba, 0x0063);    /* G1 B1 .  .  */
  w0 = __byte_perm(t, Ba, 0x2410);                  /* R0 G0 B0 R1 */
  w1 = __byte_perm(x, u, 0x5410);                   /* G1 B1 R2 G2 */
  w2 = __byte_perm(u, Bb, 0x6324);                  /* B2 R3 G3 B3 */
}

/* Cold path: a row segment that is cropped or not 16-byte aligned. */
static __device__ __noinline__ void store_row_slow(uint8_t *dst, uint4 a, uint4 b, uint4 c, int nbytes) {
  const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
  for (int i = 0; i < nbytes; i++) dst[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
}

}  // namespace jgpu
#endif
