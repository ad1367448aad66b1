/* ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under jpeg_gpu_b200/ may link, load
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * CPU restatement of the reference's 8x8 inverse DCT
 *   /root/reference/src/dct.c:21-87   (glj_real_idct8, the 1-D scaled pass)
 *   /root/reference/src/dct.c:89-98   (GLJ_REAL_IDCT8_SCALES)
 *   /root/reference/src/dct.c:100-121 (glj_real_idct8x8, the 2-D driver)
 *
 * The reference computes in IEEE binary32 with one rounding per operation
 * (Makefile:20-21 builds with -std=c89 -O2, i.e. no FMA contraction).  This
 * file keeps the exact operation order and operand order of the reference and
 * MUST be built with -ffp-contract=off and without -ffast-math; see
 * oracle/Makefile.  Parity of this restatement is pinned against the compiled
 * reference (oracle/_ref) in tests/test_oracle_golden.py and against the
 * reference's own IEEE-1180 tolerances (test/dct.c:229-261) in
 * tests/test_ieee1180.py.
 */
#include <math.h>
#include "oracle.h"

/* The reference writes its constants as double literals cast/assigned to
 * float (dct.c:51,62-65,89-98); the same spelling is kept here so that the
 * decimal -> double -> float conversion is identical. */
#define K_SQRT2   ((float)1.4142135623730950488016887242097)
#define K_1_8477  ((float)1.8477590650225735122563663787936)
#define K_1_0823  ((float)1.0823922002923939687994464107328)
#define K_2_6131  ((float)2.6131259297527530557132863468544)

static const float kScale[8] = {
  0.35355339059327376220042218105242,
  0.49039264020161522456309111806712,
  0.46193976625564337806409159469839,
  0.41573480615127261853939418880895,
  0.35355339059327376220042218105242,
  0.27778511650980111237141540697427,
  0.19134171618254488586422999201520,
  0.097545161008064133924142434238511,
};

/* One scaled 8-point inverse pass: f[0..7] are frequency-ordered inputs, the
 * eight spatial outputs land at dst[0], dst[step], ... (dct.c:21-87). */
static void inv_pass8(const float f[8], float *dst, int step) {
  /* even half: scaled inverse 4-point DCT-II on f0,f4,f2,f6 (dct.c:46-55) */
  float s04 = f[0] + f[4];
  float d04 = f[0] - f[4];
  float s26 = f[2] + f[6];
  float r26 = (f[2] - f[6]) * K_SQRT2 - s26;
  float e0 = s04 + s26;
  float e3 = s04 - s26;
  float e1 = d04 + r26;
  float e2 = d04 - r26;
  /* odd half: scaled inverse 4-point DST-IV on f1,f7,f5,f3 (dct.c:56-69) */
  float s53 = f[5] + f[3];
  float d53 = f[5] - f[3];
  float s17 = f[1] + f[7];
  float d17 = f[1] - f[7];
  float o7 = s17 + s53;
  float m5 = (s17 - s53) * K_SQRT2;
  float m8 = (d17 + d53) * K_1_8477;
  float m4 = m8 - d17 * K_1_0823;
  float m6 = m8 - d53 * K_2_6131;
  float o6 = o7 - m6;
  float o5 = o6 + m5;
  float o4 = o5 - m4;
  /* output butterflies (dct.c:70-86) */
  dst[0 * step] = e0 + o7;
  dst[1 * step] = e1 - o6;
  dst[2 * step] = e2 + o5;
  dst[3 * step] = e3 - o4;
  dst[4 * step] = e3 + o4;
  dst[5 * step] = e2 - o5;
  dst[6 * step] = e1 + o6;
  dst[7 * step] = e0 - o7;
}

void jgo_idct8x8(short *x, int xstride, const short *y, int ystride) {
  float a[64];
  float b[64];
  int r, c;
  /* two-step prescale, left to right: (y*S[r])*S[c]  (dct.c:105-110) */
  for (r = 0; r < 8; r++) {
    for (c = 0; c < 8; c++) {
      a[r * 8 + c] = (float)y[r * ystride + c] * kScale[r] * kScale[c];
    }
  }
  /* pass 1: coefficient row r -> column r of b  (dct.c:111) */
  for (r = 0; r < 8; r++) inv_pass8(a + 8 * r, b + r, 8);
  /* pass 2: +0.5 on the first element of each row of b, then row r of b ->
   * column r of a  (dct.c:112-115) */
  for (r = 0; r < 8; r++) {
    b[8 * r] = b[8 * r] + 0.5f;
    inv_pass8(b + 8 * r, a + r, 8);
  }
  /* floor to short  (dct.c:116-120) */
  for (r = 0; r < 8; r++) {
    for (c = 0; c < 8; c++) {
      x[r * xstride + c] = (short)floor(a[r * 8 + c]);
    }
  }
}

void jgo_idct_constants(float out[12]) {
  int i;
  for (i = 0; i < 8; i++) out[i] = kScale[i];
  out[8] = K_SQRT2;
  out[9] = K_1_8477;
  out[10] = K_1_0823;
  out[11] = K_2_6131;
}
