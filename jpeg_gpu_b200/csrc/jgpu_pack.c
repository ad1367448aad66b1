/* jgpu_pack.c — host side of the PACK (zero-run packed) coefficient format.
 *
 * The reference's reader can stop before de-zigzagging and hand the GPU the
 * run/level stream instead of dense planes (JPEG_DECODE_PACK,
 * src/xjpeg.c:484-496,513-519,531-535); res/horz_pack_*.fs.glsl expands it per
 * fragment.  Here the expansion is the device kernel k_unpack (jgpu_unpack.cu);
 * this file holds the host utility that turns dense QUANT planes into the
 * stream, for callers (tests, benches, other readers) that hold planes.
 *
 * Word format (src/xjpeg.c:491,516,533):
 *   DC   dc & 0xfff
 *   AC   run << 12 | value & 0xfff       (ZRL: run 15, value 0 = 0xf000)
 *   EOB  0, absent when the block's last coded coefficient is number 63
 * Blocks appear in scan order: MCU by MCU, component by component, the
 * vsamp x hsamp blocks of a component row-major (src/xjpeg.c:462-472).
 */
#include <stdlib.h>
#include <string.h>

#include "jgpu_internal.h"

/* zig-zag position -> natural (row-major) position, ITU-T T.81 figure A.6 */
static const unsigned char kNatural[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

int64_t jgpu_pack_bound(const jgpu_image_desc *desc) {
  jgpu_layout lay;
  if (!desc || jgpu_layout_query(desc, &lay) != EXIT_SUCCESS) return -1;
  return 64 * lay.coded_blocks;
}

int64_t jgpu_pack_from_quant(const jgpu_image_desc *desc, const int16_t *coef, uint16_t *pack,
                             int64_t pack_cap, int32_t *index) {
  jgpu_layout lay;
  int64_t n = 0;
  int mbx, mby, c, sby, sbx;
  if (!desc || !coef || !pack || !index) {
    jgpu_fail("jgpu_pack_from_quant: NULL argument");
    return -1;
  }
  if (jgpu_layout_query(desc, &lay) != EXIT_SUCCESS) return -1;
  memset(index, 0, (size_t)(lay.coef_len / 64) * sizeof(int32_t));
  for (mby = 0; mby < lay.nvmb; mby++) {
    for (mbx = 0; mbx < lay.nhmb; mbx++) {
      for (c = 0; c < desc->ncomps; c++) {
        const jgpu_plane_layout *pl = &lay.plane[c];
        for (sby = 0; sby < desc->vsamp[c]; sby++) {
          for (sbx = 0; sbx < desc->hsamp[c]; sbx++) {
            const int by = mby * desc->vsamp[c] + sby, bx = mbx * desc->hsamp[c] + sbx;
            const int64_t blk = pl->coef_off / 64 + (int64_t)by * pl->hblocks + bx;
            const int16_t *b = coef + 64 * blk;
            int last = 0, k, run = 0;
            if (n + 64 > pack_cap) {
              jgpu_fail("jgpu_pack_from_quant: pack buffer too small");
              return -1;
            }
            if (n > 0x7fffffff) {
              jgpu_fail("jgpu_pack_from_quant: stream exceeds the 31-bit index of the format");
              return -1;
            }
            index[blk] = (int32_t)n;
            pack[n++] = (uint16_t)(b[0] & 0xfff);
            for (k = 63; k > 0 && b[kNatural[k]] == 0; k--) {}
            last = k;
            for (k = 1; k <= last; k++) {
              const int v = b[kNatural[k]];
              if (v == 0) {
                run++;
                continue;
              }
              while (run > 15) { /* ZRL, src/xjpeg.c:509-516 with symbol 0xf0 */
                pack[n++] = 0xf000;
                run -= 16;
              }
              pack[n++] = (uint16_t)((run << 12) | (v & 0xfff));
              run = 0;
            }
            if (last < 63) pack[n++] = 0; /* EOB */
          }
        }
      }
    }
  }
  return n;
}
