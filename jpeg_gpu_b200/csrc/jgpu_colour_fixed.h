/* jgpu_colour_fixed.h — the colour offsets in integer arithmetic.
 *
 * The colour step is DEFINED in binary32 (oracle/oracle_pipeline.c jgo_colour_offsets, from
 * res/yuv.fs.glsl:11-23): per chroma sample, with cb' = Cb-128 and cr' = Cr-128 in [-128, 127],
 *     R offset = rne( fl(1.402 * cr') )
 *     G offset = rne( fl( fl(-0.34414 * cb') + fl(-0.71414 * cr') ) )
 *     B offset = rne( fl(1.772 * cb') )
 * (fl = one binary32 rounding, rne = to nearest integer, ties to even).  Each is a function of at
 * most 65536 inputs, and the fixed-point forms below reproduce ALL of them, including the two
 * exact ties of B at |cb'| = 125: tests/test_colour.py checks every input against the oracle.
 * On the GPU that is 4 IMAD + 1 shift per sample instead of 2 I2F + 4 FMUL + 4 FADD, and no
 * traffic on the conversion (XU) pipe.
 *
 * Results are returned SCALED: the offset sits in the upper 16 bits of each word (a PRMT then
 * picks it straight into an s16x2 operand).
 */
#ifndef JGPU_COLOUR_FIXED_H
#define JGPU_COLOUR_FIXED_H

#if defined(__CUDACC__)
#define JGPU_HD __host__ __device__ __forceinline__
#else
#define JGPU_HD static inline
#endif

#define JGPU_COL_KR 91879        /* 1.402   * 2^16, tuned */
#define JGPU_COL_KB 116130       /* 1.772   * 2^16 */
#define JGPU_COL_BIAS16 32767    /* half, minus one: resolves B's ties to even */
#define JGPU_COL_KG_CB (-1443428) /* -0.34414 * 2^22 */
#define JGPU_COL_KG_CR (-2995322) /* -0.71414 * 2^22 */
#define JGPU_COL_BIAS22 2097151

/* cbm, crm in [-128, 127].  offset = word >> 16 (arithmetic). */
JGPU_HD void jgpu_colour_offsets_fixed(int cbm, int crm, int *r16, int *g16, int *b16) {
  *r16 = crm * JGPU_COL_KR + JGPU_COL_BIAS16;
  *b16 = cbm * JGPU_COL_KB + JGPU_COL_BIAS16;
  *g16 = (cbm * JGPU_COL_KG_CB + (crm * JGPU_COL_KG_CR + JGPU_COL_BIAS22)) >> 6;
}

#endif
