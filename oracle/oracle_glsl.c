/* ORACLE — TEST INFRASTRUCTURE ONLY.  A COMPARATOR, not the ground truth.
 *
 * CPU emulation of the reference's OpenGL float path for `-o quant`, the three fragment-shader
 * passes the CUDA kernel replaces:
 *   pass 1  res/horz_quant_yuv.fs.glsl:81-99 / horz_quant_grey.fs.glsl:79-95
 *           y[i] = quant[j*8+i] * coef, where quant = GLJ_REAL_IDCT8X8_SCALES[k] * tbl[k] as ONE
 *           float (src/jpeg_gpu.c:34-67,1320-1338), then the 8-point pass of the shader
 *           (same factorisation and literals as src/dct.c:21-87) -> fp32 texture
 *   pass 2  res/vert.fs.glsl:79-102   y[0] += 0.5, the 8-point pass down the column,
 *           ivec4(x) -- conversion toward ZERO, not floor -- + 128 -> 16-bit integer texture
 *           (no clamp to 0..255 anywhere)
 *   pass 3  res/unyuv.fs.glsl:17-50   nearest-neighbour chroma (s>>xdec, t>>ydec), mat3 of
 *           res/unyuv.fs.glsl:12-16 on (y, u-128, v-128), /255 -> the framebuffer's unorm8
 *           conversion (clamp to [0,1], nearest);  res/ungrey.fs.glsl for one component
 * this path is numerically different from -- and a little worse than -- src/dct.c +
 * src/xjpeg.c:565-584, which is the ground truth the product is bit-exact with.  north_star
 * asks for "within +-1 LSB per channel where the GLSL float path is the comparator"; with the
 * shader's truncation every sample below the +128 bias comes out one higher than dct.c's floor,
 * so that bound holds against this emulation run with `floor_mode` = 1 (ivec4 replaced by
 * floor) and tests/test_glsl_comparator.py states the measured distances for both modes.
 *
 * Not restated: the shader's table-selection slip (`v>u_cstride`, res/horz_quant_yuv.fs.glsl:87-88,
 * the first block row of Cb is scaled with the luma table) -- tables go by component,
 * as src/xjpeg.c:443 has them.  What a GL driver does beyond the GLSL text (fusing a*b+c,
 * float -> unorm8 rounding) is implementation-defined; this file rounds once per operation
 * (build with -ffp-contract=off) and converts to the nearest integer.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define G_SQRT2   ((float)1.4142135623730950488016887242097)
#define G_1_8477  ((float)1.8477590650225735122563663787936)
#define G_1_0823  ((float)1.0823922002923939687994464107328)
#define G_2_6131  ((float)2.6131259297527530557132863468544)

/* GLJ_REAL_IDCT8_SCALES as doubles (src/dct.c:89-98); the 2-D table of src/jpeg_gpu.c:34-67 holds
 * the products S[j]*S[i] written out to 32 digits and stored as float, i.e. (float)(S[j]*S[i]) up
 * to the last bit of the double product (tests/test_glsl_comparator.py checks it against the
 * reference's source text where that is present). */
static const double kS[8] = {
  0.35355339059327376220042218105242, 0.49039264020161522456309111806712,
  0.46193976625564337806409159469839, 0.41573480615127261853939418880895,
  0.35355339059327376220042218105242, 0.27778511650980111237141540697427,
  0.19134171618254488586422999201520, 0.097545161008064133924142434238511,
};

void jgo_glsl_scales2d(float out[64]) {
  int j, i;
  for (j = 0; j < 8; j++) for (i = 0; i < 8; i++) out[j * 8 + i] = (float)(kS[j] * kS[i]);
}

/* glj_real_idct8 of the shaders (res/vert.fs.glsl:3-77): x = 8-point scaled inverse pass of y. */
static void glsl_idct8(float x[8], const float y[8]) {
  float t0 = y[0], u4 = y[1], t2 = y[2], u6 = y[3], t1 = y[4], u5 = y[5], t3 = y[6], u7 = y[7];
  float u0 = t0 + t1, u1 = t0 - t1, u3 = t2 + t3;
  float u2 = (t2 - t3) * G_SQRT2 - u3;
  float t4, t5, t6, t7, u8;
  t0 = u0 + u3; t3 = u0 - u3; t1 = u1 + u2; t2 = u1 - u2;
  t5 = u5 + u6; t6 = u5 - u6; t7 = u4 + u7; t4 = u4 - u7;
  u7 = t7 + t5;
  u5 = (t7 - t5) * G_SQRT2;
  u8 = (t4 + t6) * G_1_8477;
  u4 = u8 - t4 * G_1_0823;
  u6 = u8 - t6 * G_2_6131;
  t7 = u7; t6 = t7 - u6; t5 = t6 + u5; t4 = t5 - u4;
  x[0] = t0 + t7; x[7] = t0 - t7; x[6] = t1 + t6; x[1] = t1 - t6;
  x[2] = t2 + t5; x[5] = t2 - t5; x[4] = t3 + t4; x[3] = t3 - t4;
}

/* Passes 1 and 2 for one plane: block-linear int16 coefficients -> the integer texture of pass 2
 * (one int per sample, padded plane, NOT clamped). */
static void glsl_plane(const jgo_plane *p, const short *coef, const unsigned short *tbl, int floor_mode, int *out) {
  float quant[64], s2d[64];
  int k, bx, by, i, j;
  jgo_glsl_scales2d(s2d);
  for (k = 0; k < 64; k++) quant[k] = s2d[k] * (float)tbl[k];   /* src/jpeg_gpu.c:1323,1333 */
  for (by = 0; by < p->vblocks; by++) {
    for (bx = 0; bx < p->hblocks; bx++) {
      const short *b = coef + p->coef_off + 64ll * ((long long)by * p->hblocks + bx);
      float h[8][8];   /* pass 1: h[j][.] = row IDCT of coefficient row j */
      for (j = 0; j < 8; j++) {
        float y[8];
        for (i = 0; i < 8; i++) y[i] = quant[j * 8 + i] * (float)b[j * 8 + i];
        glsl_idct8(h[j], y);
      }
      for (i = 0; i < 8; i++) {   /* pass 2: column i */
        float y[8], x[8];
        for (j = 0; j < 8; j++) y[j] = h[j][i];
        y[0] += 0.5f;
        glsl_idct8(x, y);
        for (j = 0; j < 8; j++) {
          const int v = floor_mode ? (int)floorf(x[j]) : (int)x[j];   /* ivec4(): toward zero */
          out[(long long)(by * 8 + j) * p->width + bx * 8 + i] = v + 128;
        }
      }
    }
  }
}

static unsigned char unorm8(float c) {   /* framebuffer write of color = rgb/255.0 */
  if (!(c > 0.0f)) return 0;
  if (c >= 1.0f) return 255;
  return (unsigned char)lrintf(c * 255.0f);
}

/* The whole GL path for one image.  rgb: width*height*(1|3) bytes; samples (optional): the integer
 * texture of pass 2, planes back to back in padded size (g->data_len ints). */
int jgo_glsl_decode_image(const jgo_geom *g, const short *coef, const unsigned short *qtabs, const int *tq,
                          int floor_mode, unsigned char *rgb, int *samples) {
  int *tex = samples ? samples : (int *)malloc(sizeof(int) * (size_t)g->data_len);
  int c, s, t;
  if (tex == NULL) return 1;
  for (c = 0; c < g->ncomps; c++) {
    glsl_plane(&g->plane[c], coef, qtabs + 64 * tq[c], floor_mode, tex + g->plane[c].data_off);
  }
  if (rgb != NULL) {
    const jgo_plane *py = &g->plane[0];
    for (t = 0; t < g->height; t++) {
      for (s = 0; s < g->width; s++) {
        const float y = (float)tex[py->data_off + (long long)t * py->width + s];
        if (g->ncomps == 1) {
          rgb[(long long)t * g->width + s] = unorm8(y / 255.0f);   /* res/ungrey.fs.glsl */
        } else {
          const jgo_plane *pu = &g->plane[1], *pv = &g->plane[2];
          const float u = (float)tex[pu->data_off + (long long)(t >> pu->ydec) * pu->width + (s >> pu->xdec)] - 128.0f;
          const float v = (float)tex[pv->data_off + (long long)(t >> pv->ydec) * pv->width + (s >> pv->xdec)] - 128.0f;
          /* mat3 columns (1,1,1), (0,-0.34414,1.772), (1.402,-0.71414,0): M*vec3 summed left to right */
          const float r = 1.0f * y + 0.0f * u + 1.402f * v;
          const float gg = 1.0f * y + -0.34414f * u + -0.71414f * v;
          const float b = 1.0f * y + 1.772f * u + 0.0f * v;
          unsigned char *o = rgb + 3 * ((long long)t * g->width + s);
          o[0] = unorm8(r / 255.0f);
          o[1] = unorm8(gg / 255.0f);
          o[2] = unorm8(b / 255.0f);
        }
      }
    }
  }
  if (!samples) free(tex);
  return 0;
}

int jgo_glsl_decode_image_flat(int width, int height, int ncomps, const int *hsamp, const int *vsamp,
                               const short *coef, const unsigned short *qtabs, const int *tq, int floor_mode,
                               unsigned char *rgb, int *samples) {
  jgo_geom g;
  if (jgo_geometry(width, height, ncomps, hsamp, vsamp, &g)) return 1;
  return jgo_glsl_decode_image(&g, coef, qtabs, tq, floor_mode, rgb, samples);
}
