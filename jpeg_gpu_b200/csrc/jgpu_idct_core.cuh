/* jgpu_idct_core.cuh — the arithmetic contract of the hot path.
 *
 * 8x8 inverse DCT for TWO blocks at a time, bit-exact with the reference's
 * glj_real_idct8x8 (src/dct.c:100-121) / glj_real_idct8 (src/dct.c:21-87):
 * IEEE binary32, one round-to-nearest per operation, NO fused multiply-add,
 * denormals preserved, row pass before column pass, +0.5 on the vertical-DC
 * term between the passes, floor at the end.
 *
 * B200 mapping.  One thread owns a PAIR of blocks and keeps all 64 sample
 * pairs in registers for both passes, so there is no transpose and no
 * shared-memory traffic between the passes.  The two blocks ride in the two
 * lanes of Blackwell's packed binary32 instructions (PTX add/sub/fma
 * .rn.f32x2 -> SASS FADD2/FFMA2): per-lane results are the same IEEE values a
 * scalar FADD/FMUL gives, at half the issue slots.
 *
 * ptxas 12.9 fuses `mul.rn.f32x2` followed by `add/sub.rn.f32x2` into one
 * FFMA2 even though both carry an explicit .rn (scalar mul.rn/add.rn are not
 * fused) and even under --fmad=false.  That single rounding would break
 * bit-exactness, so every packed product here is issued as
 *     fma.rn.f32x2  p, a, b, NEGZERO      (a*b + (-0.0) == RN(a*b) exactly,
 *                                          including the sign of a zero product)
 * with NEGZERO read from __constant__ memory, which ptxas cannot fold and
 * cannot merge with the following add.  tests/test_sass_audit.py checks the
 * SASS, the GPU parity tests check the bits.
 *
 * The same source builds for the host with JGPU_CORE_HOST_EMULATION (two
 * plain floats per pair, compiled with -ffp-contract=off) so the operation
 * order can be checked against the oracle without a GPU (tests/host_core/).
 * That build is test-only and is not part of the product library.
 */
#ifndef JGPU_IDCT_CORE_CUH
#define JGPU_IDCT_CORE_CUH

#include <stdint.h>

namespace jgpu {

/* Constants of src/dct.c:51,62-65,89-98 after the double->float conversion the
 * reference performs, as exact bit patterns (printed from the oracle, see
 * tests/test_host_core.py::test_constants_are_the_references). */
#if defined(__CUDACC__)
#define JGPU_CONSTEXPR_HD __host__ __device__ constexpr
#else
#define JGPU_CONSTEXPR_HD constexpr
#endif
/* GLJ_REAL_IDCT8_SCALES[k] (src/dct.c:89-98) */
JGPU_CONSTEXPR_HD uint32_t scale_bits(int k) {
  return k == 0 ? 0x3eb504f3u : k == 1 ? 0x3efb14beu : k == 2 ? 0x3eec835eu
       : k == 3 ? 0x3ed4db31u : k == 4 ? 0x3eb504f3u : k == 5 ? 0x3e8e39dau
       : k == 6 ? 0x3e43ef15u : 0x3dc7c5c2u;
}
static constexpr uint32_t kSqrt2Bits = 0x3fb504f3u;  /* 1.41421356... */
static constexpr uint32_t k18477Bits = 0x3fec835eu;  /* 1.84775906... */
static constexpr uint32_t k10823Bits = 0x3f8a8bd4u;  /* 1.08239220... */
static constexpr uint32_t k26131Bits = 0x40273d75u;  /* 2.61312592... */

#ifdef JGPU_CORE_HOST_EMULATION
/* ---- host emulation of a packed pair (tests only) ---------------------- */
#define JGPU_DEV inline
struct pair32 { float lo, hi; };
static inline float bits_f32(uint32_t b) { float f; __builtin_memcpy(&f, &b, 4); return f; }
static inline uint32_t f32_bits(float f) { uint32_t b; __builtin_memcpy(&b, &f, 4); return b; }
JGPU_DEV pair32 p_make(float lo, float hi) { pair32 r = {lo, hi}; return r; }
JGPU_DEV pair32 p_add(pair32 a, pair32 b) { return p_make(a.lo + b.lo, a.hi + b.hi); }
JGPU_DEV pair32 p_sub(pair32 a, pair32 b) { return p_make(a.lo - b.lo, a.hi - b.hi); }
JGPU_DEV pair32 p_mulc(pair32 a, uint32_t cbits) {
  float c = bits_f32(cbits);
  return p_make(a.lo * c, a.hi * c);
}
JGPU_DEV pair32 p_mul(pair32 a, pair32 b) { return p_make(a.lo * b.lo, a.hi * b.hi); }
JGPU_DEV pair32 p_add_half(pair32 a) { return p_make(a.lo + 0.5f, a.hi + 0.5f); }
#else
/* ---- device: Blackwell packed binary32 --------------------------------- */
#define JGPU_DEV __device__ __forceinline__
typedef unsigned long long pair32;

/* (-0.0f, -0.0f); see the header comment. */
__constant__ unsigned long long c_negzero2 = 0x8000000080000000ULL;

JGPU_DEV pair32 p_make(float lo, float hi) {
  pair32 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
JGPU_DEV pair32 p_make_bits(uint32_t lo, uint32_t hi) {
  pair32 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
JGPU_DEV void p_split_bits(pair32 v, uint32_t &lo, uint32_t &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
JGPU_DEV pair32 p_add(pair32 a, pair32 b) {
  pair32 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
JGPU_DEV pair32 p_sub(pair32 a, pair32 b) {
  pair32 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
/* a*b rounded once, never merged with a neighbouring add. */
JGPU_DEV pair32 p_mul(pair32 a, pair32 b) {
  pair32 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c_negzero2));
  return r;
}
JGPU_DEV pair32 p_mulc(pair32 a, uint32_t cbits) {
  return p_mul(a, p_make_bits(cbits, cbits));
}
JGPU_DEV pair32 p_add_half(pair32 a) {
  return p_add(a, p_make_bits(0x3f000000u, 0x3f000000u));
}
/* a + b rounded toward -infinity in both lanes (FADD2.RM). */
JGPU_DEV pair32 p_add_rm(pair32 a, pair32 b) {
  pair32 r;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
#endif

/* One scaled 8-point inverse pass, in place: v[0..7] hold frequency-ordered
 * inputs on entry and the eight spatial outputs x0..x7 on exit.  Operation
 * and operand order are those of src/dct.c:21-87. */
JGPU_DEV void inv_pass8(pair32 (&v)[8]) {
  /* even half (src/dct.c:46-55) */
  pair32 s04 = p_add(v[0], v[4]);
  pair32 d04 = p_sub(v[0], v[4]);
  pair32 s26 = p_add(v[2], v[6]);
  pair32 r26 = p_sub(p_mulc(p_sub(v[2], v[6]), kSqrt2Bits), s26);
  pair32 e0 = p_add(s04, s26);
  pair32 e3 = p_sub(s04, s26);
  pair32 e1 = p_add(d04, r26);
  pair32 e2 = p_sub(d04, r26);
  /* odd half (src/dct.c:56-69) */
  pair32 s53 = p_add(v[5], v[3]);
  pair32 d53 = p_sub(v[5], v[3]);
  pair32 s17 = p_add(v[1], v[7]);
  pair32 d17 = p_sub(v[1], v[7]);
  pair32 o7 = p_add(s17, s53);
  pair32 m5 = p_mulc(p_sub(s17, s53), kSqrt2Bits);
  pair32 m8 = p_mulc(p_add(d17, d53), k18477Bits);
  pair32 m4 = p_sub(m8, p_mulc(d17, k10823Bits));
  pair32 m6 = p_sub(m8, p_mulc(d53, k26131Bits));
  pair32 o6 = p_sub(o7, m6);
  pair32 o5 = p_add(o6, m5);
  pair32 o4 = p_sub(o5, m4);
  /* output butterflies (src/dct.c:70-86) */
  v[0] = p_add(e0, o7);
  v[1] = p_sub(e1, o6);
  v[2] = p_add(e2, o5);
  v[3] = p_sub(e3, o4);
  v[4] = p_add(e3, o4);
  v[5] = p_sub(e2, o5);
  v[6] = p_add(e1, o6);
  v[7] = p_sub(e0, o7);
}

/* Two-step prescale of one dequantised sample pair at (row r, column c):
 * (y*S[r])*S[c], left to right (src/dct.c:105-110). */
JGPU_DEV pair32 prescale(pair32 y, int r, int c) {
  return p_mulc(p_mulc(y, scale_bits(r)), scale_bits(c));
}

/* Column pass over a register tile m[r][c] whose row r already holds the
 * eight outputs of the row pass on coefficient row r (src/dct.c:111): column
 * c of m is row c of the reference's `z`.  Adds 0.5 to the vertical-DC term
 * and transforms (src/dct.c:112-115).  On exit m[k][c] is the un-floored
 * sample at pixel row k, column c. */
JGPU_DEV void column_pass(pair32 (&m)[8][8]) {
#pragma unroll
  for (int c = 0; c < 8; c++) {
    pair32 v[8];
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = m[r][c];
    v[0] = p_add_half(v[0]);
    inv_pass8(v);
#pragma unroll
    for (int k = 0; k < 8; k++) m[k][c] = v[k];
  }
}

/* Same column pass, two columns at a time, handing each finished column pair
 * (8 rows x 2 adjacent columns) to `sink(j, left, right)`, j = 0..3.  Lets a
 * consumer narrow the samples as they are produced instead of keeping all 64
 * un-floored pairs live. */
template <typename Sink>
JGPU_DEV void column_pass_by_pairs(pair32 (&m)[8][8], Sink &&sink) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    pair32 u[8], v[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      u[r] = m[r][2 * j];
      v[r] = m[r][2 * j + 1];
    }
    u[0] = p_add_half(u[0]);
    v[0] = p_add_half(v[0]);
    inv_pass8(u);
    inv_pass8(v);
    sink(j, u, v);
  }
}

/* As column_pass_by_pairs, but row 0 of the tile is not in registers: `row0(j, a, b)` fetches
 * m[0][2j] and m[0][2j+1] from wherever the caller parked them after the row pass (shared
 * memory).  Parking one row keeps the live set under the register budget that 12 warps per
 * SM allow, so nothing spills to local memory. */
template <typename Row0, typename Sink>
JGPU_DEV void column_pass_by_pairs_parked(pair32 (&m)[8][8], Row0 &&row0, Sink &&sink) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    pair32 u[8], v[8];
    row0(j, u[0], v[0]);
#pragma unroll
    for (int r = 1; r < 8; r++) {
      u[r] = m[r][2 * j];
      v[r] = m[r][2 * j + 1];
    }
    u[0] = p_add_half(u[0]);
    v[0] = p_add_half(v[0]);
    inv_pass8(u);
    inv_pass8(v);
    sink(j, u, v);
  }
}

}  // namespace jgpu
#endif
