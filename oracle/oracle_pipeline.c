/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_idct.c header).
 *
 * CPU restatement of everything around the IDCT on the coefficient -> RGB
 * path: geometry, dequantisation, bias/clamp/plane store, nearest-neighbour
 * chroma upsample and the colour matrix.  Built twice by oracle/Makefile:
 *   - into oracle/_build/liboracle.so with JGO_IDCT_FN = jgo_idct8x8 (our
 *     restatement, cpu_baseline kind "port");
 *   - into oracle/_ref/libjgpu_ref.so with JGO_IDCT_FN = glj_real_idct8x8, the
 *     reference's own object compiled from /root/reference/src/dct.c
 *     (cpu_baseline kind "reference").
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#ifndef JGO_IDCT_FN
#define JGO_IDCT_FN jgo_idct8x8
#define JGO_IDCT_NAME "port:oracle/oracle_idct.c"
#else
void JGO_IDCT_FN(short *x, int xstride, const short *y, int ystride);
#define JGO_IDCT_NAME "reference:src/dct.c"
#endif

const char *jgo_idct_name(void) { return JGO_IDCT_NAME; }

/* src/internal.c:49-67 (glj_ilog): number of bits needed to hold v. */
static int bit_length(unsigned v) {
  int n = 0;
  while (v) { n++; v >>= 1; }
  return n;
}

/* Geometry: src/xjpeg.c:400-407 (nhmb/nvmb), src/jpeg_wrap.c:301-308
 * (hblocks/vblocks), src/image.c:24-97 (plane size, xdec/ydec, cstride, the
 * running coef pointer). */
int jgo_geometry(int width, int height, int ncomps, const int *hsamp,
                 const int *vsamp, jgo_geom *g) {
  int i;
  long long coef = 0, data = 0;
  if (width <= 0 || height <= 0 || (ncomps != 1 && ncomps != 3)) return 1;
  memset(g, 0, sizeof(*g));
  g->width = width;
  g->height = height;
  g->ncomps = ncomps;
  for (i = 0; i < ncomps; i++) {
    if (hsamp[i] < 1 || hsamp[i] > 4 || vsamp[i] < 1 || vsamp[i] > 4) return 1;
    if (hsamp[i] > g->hmax) g->hmax = hsamp[i];
    if (vsamp[i] > g->vmax) g->vmax = vsamp[i];
  }
  g->nhmb = (width + 8 * g->hmax - 1) / (8 * g->hmax);
  g->nvmb = (height + 8 * g->vmax - 1) / (8 * g->vmax);
  for (i = 0; i < ncomps; i++) {
    jgo_plane *p = &g->plane[i];
    p->hsamp = hsamp[i];
    p->vsamp = vsamp[i];
    p->hblocks = g->nhmb * hsamp[i];
    p->vblocks = g->nvmb * vsamp[i];
    p->width = p->hblocks * 8;
    p->height = p->vblocks * 8;
    p->xdec = bit_length(g->hmax) - bit_length(hsamp[i]);
    p->ydec = bit_length(g->vmax) - bit_length(vsamp[i]);
    p->cstride = (p->vblocks + ((1 << p->xdec) - 1)) >> p->xdec;
    p->coef_off = coef;
    p->data_off = data;
    coef += ((long long)p->width << (p->xdec + 3)) * p->cstride;
    data += (long long)p->width * p->height;
  }
  g->coef_len = coef;
  g->data_len = data;
  g->rgb_len = (long long)width * height * (ncomps == 1 ? 1 : 3);
  return 0;
}

/* out[0..7]  = width,height,ncomps,hmax,vmax,nhmb,nvmb,(unused)
 * out[8..10] = coef_len,data_len,rgb_len
 * out[16+12*p ..] = hsamp,vsamp,hblocks,vblocks,width,height,xdec,ydec,
 *                   cstride,coef_off,data_off,(unused)                    */
int jgo_geometry_flat(int width, int height, int ncomps, const int *hsamp,
                      const int *vsamp, long long *out) {
  jgo_geom g;
  int i;
  if (jgo_geometry(width, height, ncomps, hsamp, vsamp, &g)) return 1;
  memset(out, 0, sizeof(long long) * (16 + 12 * JGO_MAX_PLANES));
  out[0] = g.width; out[1] = g.height; out[2] = g.ncomps; out[3] = g.hmax;
  out[4] = g.vmax; out[5] = g.nhmb; out[6] = g.nvmb;
  out[8] = g.coef_len; out[9] = g.data_len; out[10] = g.rgb_len;
  for (i = 0; i < g.ncomps; i++) {
    const jgo_plane *p = &g.plane[i];
    long long *o = out + 16 + 12 * i;
    o[0] = p->hsamp; o[1] = p->vsamp; o[2] = p->hblocks; o[3] = p->vblocks;
    o[4] = p->width; o[5] = p->height; o[6] = p->xdec; o[7] = p->ydec;
    o[8] = p->cstride; o[9] = p->coef_off; o[10] = p->data_off;
  }
  return 0;
}

/* Where block (bx,by) of a plane lives inside image.coef: src/xjpeg.c:550-563.
 * w0x8 is plane[0].width<<3, i.e. the shorts in one luma block row. */
static long long block_offset(const jgo_geom *g, const jgo_plane *p, int bx,
                              int by) {
  long long w0x8 = (long long)g->plane[0].width << 3;
  return p->coef_off + w0x8 * (by >> p->xdec) +
         (w0x8 >> p->xdec) * (by & ((1 << p->xdec) - 1)) + ((long long)bx << 6);
}

/* src/internal.h:36-37 (GLJ_CLAMP255) */
static unsigned char clamp255(int v) {
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

/* One block: dequantise into short (src/xjpeg.c:501-503,524-527: the int
 * product is stored to a short, i.e. wraps), inverse DCT (src/xjpeg.c:569),
 * +128 / clamp / store with the plane's row stride (src/xjpeg.c:570-583). */
static void decode_block(const short *coef, const unsigned short *q,
                         unsigned char *dst, int pitch) {
  short blk[64];
  int k, r;
  for (k = 0; k < 64; k++) blk[k] = (short)(coef[k] * q[k]);
  JGO_IDCT_FN(blk, 8, blk, 8);
  for (r = 0; r < 8; r++) {
    for (k = 0; k < 8; k++) dst[r * pitch + k] = clamp255(blk[r * 8 + k] + 128);
  }
}

int jgo_coef_to_yuv(const jgo_geom *g, const short *coef,
                    const unsigned short *qtabs, const int *tq,
                    unsigned char *planes) {
  int i, bx, by;
  for (i = 0; i < g->ncomps; i++) {
    const jgo_plane *p = &g->plane[i];
    const unsigned short *q = qtabs + 64 * tq[i];
    unsigned char *base = planes + p->data_off;
    for (by = 0; by < p->vblocks; by++) {
      for (bx = 0; bx < p->hblocks; bx++) {
        decode_block(coef + block_offset(g, p, bx, by), q,
                     base + (long long)by * 8 * p->width + bx * 8, p->width);
      }
    }
  }
  return 0;
}

/* Colour stage.  The xjpeg backend has no RGB output (src/jpeg_wrap.c:321-342),
 * so the only colour formula the reference owns is the fragment shader
 * res/yuv.fs.glsl:11-23 (same matrix in res/unyuv.fs.glsl:12-16), applied to
 * the CLAMPED u8 planes with nearest-neighbour chroma (s>>xdec, t>>ydec):
 *     R = Y + 1.402 (Cr-128)
 *     G = Y - 0.34414 (Cb-128) - 0.71414 (Cr-128)
 *     B = Y + 1.772 (Cb-128)
 * followed by GL's float -> unorm8 conversion (clamp, round to nearest).
 * This oracle DEFINES the arithmetic as: each product is one binary32
 * multiply, the two G products are added in binary32, each of the three
 * chroma offsets is rounded to the nearest integer (ties to even), and the
 * integer offset is added to the integer Y and clamped to 0..255.  Because Y
 * is an integer this equals rounding the exact sum Y+offset to nearest.
 * Against a GL implementation the honest claim is +-1 LSB; that tolerance is
 * tested separately (tests/test_colour.py). */
void jgo_colour_offsets(int cb, int cr, int out[3]) {
  float cbf = (float)cb - 128.0f;
  float crf = (float)cr - 128.0f;
  float rc = 1.402f * crf;
  float g1 = -0.34414f * cbf;
  float g2 = -0.71414f * crf;
  float gc = g1 + g2;
  float bc = 1.772f * cbf;
  out[0] = (int)lrintf(rc);
  out[1] = (int)lrintf(gc);
  out[2] = (int)lrintf(bc);
}

int jgo_yuv_to_rgb(const jgo_geom *g, const unsigned char *planes,
                   unsigned char *rgb) {
  int x, y;
  const jgo_plane *py = &g->plane[0];
  if (g->ncomps == 1) {
    /* grey: one byte per pixel, the libjpeg convention the reference's RGB
     * surface follows for 1-component files (src/jpeg_wrap.c:215-220). */
    for (y = 0; y < g->height; y++) {
      memcpy(rgb + (long long)y * g->width,
             planes + py->data_off + (long long)y * py->width, g->width);
    }
    return 0;
  }
  {
    const jgo_plane *pb = &g->plane[1];
    const jgo_plane *pr = &g->plane[2];
    for (y = 0; y < g->height; y++) {
      const unsigned char *yrow = planes + py->data_off + (long long)y * py->width;
      const unsigned char *brow =
          planes + pb->data_off + (long long)(y >> pb->ydec) * pb->width;
      const unsigned char *rrow =
          planes + pr->data_off + (long long)(y >> pr->ydec) * pr->width;
      unsigned char *o = rgb + (long long)y * g->width * 3;
      for (x = 0; x < g->width; x++) {
        int ofs[3];
        int yy = yrow[x];
        jgo_colour_offsets(brow[x >> pb->xdec], rrow[x >> pr->xdec], ofs);
        o[3 * x + 0] = clamp255(yy + ofs[0]);
        o[3 * x + 1] = clamp255(yy + ofs[1]);
        o[3 * x + 2] = clamp255(yy + ofs[2]);
      }
    }
  }
  return 0;
}

/* ---- batch driver: the CPU baseline ------------------------------------- */

#define DESC_STRIDE 16
/* desc[i*16+..]: 0 width, 1 height, 2 ncomps, 3-5 hsamp, 6-8 vsamp, 9-11 tq,
 * 12 coef_off (shorts), 13 rgb_off (bytes), 14 qtab set index (x 4 x 64),
 * 15 planes_off (bytes, or -1 for none). */

typedef struct batch_job {
  int n;
  const long long *desc;
  const short *coef;
  const unsigned short *qtabs;
  unsigned char *rgb;
  unsigned char *planes;
  jgo_geom *geom;
  long long *first_item;  /* prefix sum of MCU rows per image */
  long long total_items;
  long long next;         /* atomic work counter */
  int error;
  short *gofs;            /* [256*256] green offsets */
  short rofs[256], bofs[256];
} batch_job;

static void decode_mcu_row(const batch_job *job, int img, int mrow,
                           unsigned char *strip) {
  const jgo_geom *g = &job->geom[img];
  const long long *d = job->desc + (long long)img * DESC_STRIDE;
  const short *coef = job->coef + d[12];
  const unsigned short *qt = job->qtabs + d[14] * (JGO_MAX_QTABS * 64);
  unsigned char *sp[JGO_MAX_PLANES];
  int i, bx, sby, x, y;
  long long off = 0;
  for (i = 0; i < g->ncomps; i++) {
    const jgo_plane *p = &g->plane[i];
    sp[i] = strip + off;
    off += (long long)p->width * p->vsamp * 8;
    for (sby = 0; sby < p->vsamp; sby++) {
      int by = mrow * p->vsamp + sby;
      for (bx = 0; bx < p->hblocks; bx++) {
        decode_block(coef + block_offset(g, p, bx, by), qt + 64 * d[9 + i],
                     sp[i] + (long long)sby * 8 * p->width + bx * 8, p->width);
      }
    }
    if (job->planes && d[15] >= 0) {
      memcpy(job->planes + d[15] + p->data_off +
                 (long long)mrow * p->vsamp * 8 * p->width,
             sp[i], (size_t)p->width * p->vsamp * 8);
    }
  }
  {
    int rows = g->vmax * 8;
    int y0 = mrow * rows;
    unsigned char *out = job->rgb + d[13];
    if (y0 + rows > g->height) rows = g->height - y0;
    for (y = 0; y < rows; y++) {
      const unsigned char *yrow = sp[0] + (long long)y * g->plane[0].width;
      if (g->ncomps == 1) {
        memcpy(out + (long long)(y0 + y) * g->width, yrow, g->width);
      } else {
        const jgo_plane *pb = &g->plane[1];
        const jgo_plane *pr = &g->plane[2];
        const unsigned char *brow = sp[1] + (long long)(y >> pb->ydec) * pb->width;
        const unsigned char *rrow = sp[2] + (long long)(y >> pr->ydec) * pr->width;
        unsigned char *o = out + (long long)(y0 + y) * g->width * 3;
        for (x = 0; x < g->width; x++) {
          int yy = yrow[x];
          int cb = brow[x >> pb->xdec];
          int cr = rrow[x >> pr->xdec];
          o[3 * x + 0] = clamp255(yy + job->rofs[cr]);
          o[3 * x + 1] = clamp255(yy + job->gofs[cb * 256 + cr]);
          o[3 * x + 2] = clamp255(yy + job->bofs[cb]);
        }
      }
    }
  }
}

static void *batch_worker(void *arg) {
  batch_job *job = (batch_job *)arg;
  unsigned char *strip = NULL;
  size_t strip_cap = 0;
  int img = 0;
  for (;;) {
    long long item = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);
    size_t need = 0;
    int i;
    if (item >= job->total_items) break;
    while (img + 1 < job->n && job->first_item[img + 1] <= item) img++;
    while (job->first_item[img] > item) img--;
    for (i = 0; i < job->geom[img].ncomps; i++) {
      need += (size_t)job->geom[img].plane[i].width * job->geom[img].plane[i].vsamp * 8;
    }
    if (need > strip_cap) {
      free(strip);
      strip = (unsigned char *)malloc(need);
      strip_cap = need;
      if (!strip) { job->error = 1; break; }
    }
    decode_mcu_row(job, img, (int)(item - job->first_item[img]), strip);
  }
  free(strip);
  return NULL;
}

int jgo_decode_batch(int n, const long long *desc, const short *coef,
                     const unsigned short *qtabs, unsigned char *rgb,
                     unsigned char *planes, int nthreads) {
  batch_job job;
  pthread_t *th;
  int i, cb, cr, rc = 0;
  if (n <= 0) return 0;
  if (nthreads < 1) nthreads = 1;
  memset(&job, 0, sizeof(job));
  job.n = n; job.desc = desc; job.coef = coef; job.qtabs = qtabs;
  job.rgb = rgb; job.planes = planes;
  job.geom = (jgo_geom *)calloc((size_t)n, sizeof(jgo_geom));
  job.first_item = (long long *)calloc((size_t)n + 1, sizeof(long long));
  job.gofs = (short *)malloc(65536 * sizeof(short));
  th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  if (!job.geom || !job.first_item || !job.gofs || !th) { rc = 1; goto done; }
  for (i = 0; i < n; i++) {
    const long long *d = desc + (long long)i * DESC_STRIDE;
    int hs[3], vs[3], k;
    for (k = 0; k < 3; k++) { hs[k] = (int)d[3 + k]; vs[k] = (int)d[6 + k]; }
    if (jgo_geometry((int)d[0], (int)d[1], (int)d[2], hs, vs, &job.geom[i])) {
      rc = 1;
      goto done;
    }
    job.first_item[i + 1] = job.first_item[i] + job.geom[i].nvmb;
  }
  job.total_items = job.first_item[n];
  for (cb = 0; cb < 256; cb++) {
    for (cr = 0; cr < 256; cr++) {
      int ofs[3];
      jgo_colour_offsets(cb, cr, ofs);
      job.gofs[cb * 256 + cr] = (short)ofs[1];
      if (cb == 0) job.rofs[cr] = (short)ofs[0];
      if (cr == 0) job.bofs[cb] = (short)ofs[2];
    }
  }
  for (i = 1; i < nthreads; i++) pthread_create(&th[i], NULL, batch_worker, &job);
  batch_worker(&job);
  for (i = 1; i < nthreads; i++) pthread_join(th[i], NULL);
  rc = job.error;
done:
  free(job.geom); free(job.first_item); free(job.gofs); free(th);
  return rc;
}

/* ---- bare-block helpers -------------------------------------------------- */

void jgo_idct_blocks(short *blocks, long long nblocks) {
  long long b;
  for (b = 0; b < nblocks; b++) JGO_IDCT_FN(blocks + 64 * b, 8, blocks + 64 * b, 8);
}

void jgo_dequant_idct_blocks(short *blocks, long long nblocks,
                             const unsigned short *q) {
  long long b;
  int k;
  for (b = 0; b < nblocks; b++) {
    short *blk = blocks + 64 * b;
    for (k = 0; k < 64; k++) blk[k] = (short)(blk[k] * q[k]);
    JGO_IDCT_FN(blk, 8, blk, 8);
  }
}

/* ---- IEEE-1180 input generator ------------------------------------------
 * Follows the procedure of the reference's unit test (test/dct.c:62-146):
 * LCG randx*1103515245+12345, pixels in [low,high]*sign, double-precision
 * forward DCT, round + clamp to [-2048,2047], double-precision inverse as the
 * comparator, clamp to [-256,255].  state is the LCG state (in/out) so a
 * caller can chain the ranges exactly as test/dct.c:237-256 does. */
static double basis8[8][8];
static int basis8_ready;

static void basis8_init(void) {
  int j, i;
  if (basis8_ready) return;
  for (j = 0; j < 8; j++) {
    long double cj = j == 0 ? sqrtl(0.125L) : 0.5L;
    for (i = 0; i < 8; i++) {
      basis8[j][i] = (double)(cj * cosl((2 * i + 1) * j * 3.14159265358979323846264338327950288L / 16));
    }
  }
  basis8_ready = 1;
}

static int lcg_next(unsigned *state, int low, int high) {
  double x;
  *state = *state * 1103515245U + 12345U;
  x = ((int)*state & 0x7ffffffe) / ((double)0x7fffffff) * (high - low + 1);
  return (int)x + low;
}

void jgo_ieee1180_gen(unsigned *state, int low, int high, int sign,
                      long long nblocks, short *coef, short *ref) {
  long long b;
  int j, i, k;
  basis8_init();
  for (b = 0; b < nblocks; b++) {
    double px[64], t[64], f[64];
    for (k = 0; k < 64; k++) px[k] = (short)(lcg_next(state, low, high) * sign);
    /* forward: columns then rows, as test/dct.c:104-109 */
    for (i = 0; i < 8; i++)
      for (j = 0; j < 8; j++) {
        double s = 0;
        for (k = 0; k < 8; k++) s += basis8[j][k] * px[k * 8 + i];
        t[i * 8 + j] = s;
      }
    for (i = 0; i < 8; i++)
      for (j = 0; j < 8; j++) {
        double s = 0;
        for (k = 0; k < 8; k++) s += basis8[j][k] * t[k * 8 + i];
        f[i * 8 + j] = s;
      }
    for (k = 0; k < 64; k++) {
      int v = (int)floor(f[k] + 0.5);
      v = v < -2048 ? -2048 : (v > 2047 ? 2047 : v);
      coef[b * 64 + k] = (short)v;
      f[k] = v;
    }
    /* inverse in double: rows to columns, then back, test/dct.c:111-116 */
    for (i = 0; i < 8; i++)
      for (j = 0; j < 8; j++) {
        double s = 0;
        for (k = 0; k < 8; k++) s += basis8[k][j] * f[i * 8 + k];
        t[j * 8 + i] = s;
      }
    for (i = 0; i < 8; i++)
      for (j = 0; j < 8; j++) {
        double s = 0;
        for (k = 0; k < 8; k++) s += basis8[k][j] * t[i * 8 + k];
        px[j * 8 + i] = s;
      }
    for (k = 0; k < 64; k++) {
      int v = (int)floor(px[k] + 0.5);
      ref[b * 64 + k] = (short)(v < -256 ? -256 : (v > 255 ? 255 : v));
    }
  }
}
