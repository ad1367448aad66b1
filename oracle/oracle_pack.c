/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_idct.c header).
 *
 * CPU restatement of the reference's PACK (zero-run packed) coefficient format:
 *   jgo_unpack_image  the consumer, res/horz_pack_yuv.fs.glsl:105-127 (the same loop is in
 *                     horz_pack_grey.fs.glsl): what every fragment does before the row
 *                     transform, done once per block into the dense layout of
 *                     src/xjpeg.c:550-563;
 *   jgo_pack_image    the producer, src/xjpeg.c:484-496 (DC word, index), 513-519 (AC word),
 *                     531-535 (EOB word), driven by dense coefficients instead of a bitstream,
 *                     in the scan order of src/xjpeg.c:462-472.
 * The product's implementations are jpeg_gpu_b200/csrc/jgpu_unpack.cu and jgpu_pack.c.
 */
#include <string.h>
#include "oracle.h"

/* res/horz_pack_yuv.fs.glsl:3-12 */
static const int DE_ZIG_ZAG[64] = {
   0,  1,  8, 16,  9,  2,  3, 10,
  17, 24, 32, 25, 18, 11,  4,  5,
  12, 19, 26, 33, 40, 48, 41, 34,
  27, 20, 13,  6,  7, 14, 21, 28,
  35, 42, 49, 56, 57, 50, 43, 36,
  29, 22, 15, 23, 30, 37, 44, 51,
  58, 59, 52, 45, 38, 31, 39, 46,
  53, 60, 61, 54, 47, 55, 62, 63
};

/* res/horz_pack_yuv.fs.glsl:112,124: 12-bit two's complement -> int */
static int sext12(unsigned p) {
  p &= 0xfffu;
  return (int)(p | ((p & 0x800u) == 0x800u ? ~0xfffu : 0u));
}

/* index has the layout of image.index (src/image.c:85-95): plane p's entries start at
 * plane[p].coef_off/64 and entry by*hblocks + bx belongs to block (bx, by)
 * (src/xjpeg.c:491 with ystride>>3 == hblocks).  coef receives g->coef_len shorts. */
int jgo_unpack_image(const jgo_geom *g, const unsigned short *pack, long long pack_len,
                     const int *index, short *coef) {
  int c, by, bx;
  for (c = 0; c < g->ncomps; c++) {
    const jgo_plane *pl = &g->plane[c];
    for (by = 0; by < pl->vblocks; by++) {
      for (bx = 0; bx < pl->hblocks; bx++) {
        long long blk = pl->coef_off / 64 + (long long)by * pl->hblocks + bx;
        short *b = coef + 64 * blk;
        long long i = index[blk];
        int j = 0;
        unsigned p;
        memset(b, 0, 64 * sizeof(short));
        if (i < 0 || i >= pack_len) return 1;
        p = pack[i++];                                  /* :108-109 */
        b[0] = (short)sext12(p);                        /* :112 */
        while (j < 63) {                                /* :114 */
          int len;
          if (i >= pack_len) return 1;
          p = pack[i++];                                /* :115-116 */
          if (p == 0) break;                            /* :117-119 */
          len = (int)((p >> 12) & 0xf) + 1;             /* :120 */
          j += len;                                     /* :123 */
          if (j > 63) return 2;
          b[DE_ZIG_ZAG[j]] = (short)sext12(p);          /* :121-124 */
        }
      }
    }
  }
  return 0;
}

/* Returns the number of words written (or -1).  packed[c] counts the words of component c
 * (image_plane.packed, src/xjpeg.c:492,515,532). */
long long jgo_pack_image(const jgo_geom *g, const short *coef, unsigned short *pack,
                         long long pack_cap, int *index, int *packed) {
  long long n = 0;
  int mby, mbx, c, sby, sbx;
  memset(index, 0, (size_t)(g->coef_len / 64) * sizeof(int));
  for (c = 0; c < g->ncomps; c++) packed[c] = 0;
  for (mby = 0; mby < g->nvmb; mby++) {
    for (mbx = 0; mbx < g->nhmb; mbx++) {
      for (c = 0; c < g->ncomps; c++) {
        const jgo_plane *pl = &g->plane[c];
        for (sby = 0; sby < pl->vsamp; sby++) {
          for (sbx = 0; sbx < pl->hsamp; sbx++) {
            int by = mby * pl->vsamp + sby, bx = mbx * pl->hsamp + sbx;
            long long blk = pl->coef_off / 64 + (long long)by * pl->hblocks + bx;
            const short *b = coef + 64 * blk;
            int j = 0, k, last = 0;
            if (n + 64 > pack_cap) return -1;
            index[blk] = (int)n;                                        /* :491 */
            packed[c]++;
            pack[n++] = (unsigned short)(b[0] & 0xfff);                 /* :493 */
            for (k = 1; k < 64; k++) if (b[DE_ZIG_ZAG[k]]) last = k;
            /* a baseline encoder codes each non-zero AC coefficient as (run, size) with
             * run < 16 and a ZRL symbol (0xf0) per 16 skipped zeros; the reader turns each
             * symbol into one word (:509-519) */
            while (j < last) {
              int run = 0, v;
              k = j + 1;
              while (b[DE_ZIG_ZAG[k]] == 0) { k++; run++; }
              while (run > 15) {
                packed[c]++;
                pack[n++] = (unsigned short)((0xf << 12) | 0);
                run -= 16;
                j += 16;
              }
              v = b[DE_ZIG_ZAG[k]];
              packed[c]++;
              pack[n++] = (unsigned short)(((run & 0xf) << 12) | (v & 0xfff));  /* :516 */
              j = k;
            }
            if (j < 63) {                                               /* :529-535, :542 */
              packed[c]++;
              pack[n++] = 0;
            }
          }
        }
      }
    }
  }
  return n;
}

/* ctypes-friendly entry points: geometry from the frame parameters */
int jgo_unpack_image_flat(int width, int height, int ncomps, const int *hsamp, const int *vsamp,
                          const unsigned short *pack, long long pack_len, const int *index,
                          short *coef) {
  jgo_geom g;
  if (jgo_geometry(width, height, ncomps, hsamp, vsamp, &g)) return 10;
  return jgo_unpack_image(&g, pack, pack_len, index, coef);
}

long long jgo_pack_image_flat(int width, int height, int ncomps, const int *hsamp, const int *vsamp,
                              const short *coef, unsigned short *pack, long long pack_cap,
                              int *index, int *packed) {
  jgo_geom g;
  if (jgo_geometry(width, height, ncomps, hsamp, vsamp, &g)) return -10;
  return jgo_pack_image(&g, coef, pack, pack_cap, index, packed);
}
