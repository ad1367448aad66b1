/* jgpu_huff_prep.c — host-side preparation for the GPU entropy decoder (jgpu_huff_core.h):
 * decoder tables from the canonical description a DHT segment carries. */
#include <string.h>
#include "jgpu_huff_core.h"

/* T.81 C.2 (code generation) and F.2.2.3 (decoding by code length), laid out for a decoder
 * that looks at a left-aligned 16-bit window: the reference builds its own lookahead table
 * from the same description, src/xjpeg.c:311-336. */
int jgpu_huff_build_table(jgpu_huff_table *t, const unsigned char counts[16], const unsigned char *symbols, int ac) {
  int code = 0, k = 0, len, i;
  memset(t, 0, sizeof(*t));
  for (len = 1; len <= 16; len++) {
    const int n = counts[len - 1];
    if (code + n > (1 << len)) return 1; /* over-subscribed */
    t->delta[len] = k - code;
    if (len <= JGPU_HUFF_LUT_BITS) {
      for (i = 0; i < n; i++) {
        const int first = (code + i) << (JGPU_HUFF_LUT_BITS - len);
        const int span = 1 << (JGPU_HUFF_LUT_BITS - len);
        int s;
        for (s = 0; s < span; s++) t->lut[first + s] = (uint16_t)JGPU_HUFF_ENTRY((unsigned)len, (unsigned)symbols[k + i], ac);
      }
    }
    k += n;
    if (k > 256) return 1;
    code += n;
    /* every 16-bit window below this value starts with a code of at most `len` bits */
    t->limit[len] = (uint32_t)code << (16 - len);
    code <<= 1;
  }
  memcpy(t->symbols, symbols, (size_t)k);
  return 0;
}

/* The placement of src/xjpeg.c:550-563 (block-linear inside a plane), split into the terms the
 * write pass adds up: block c of MCU (mbx, mby) sits at plane_off + ((mby*vs + dy) * hblocks +
 * mbx*hs + dx) * 64. */
void jgpu_huff_file_finish(jgpu_huff_file *f) {
  int c;
  for (c = 0; c < f->bpm; c++) {
    const int comp = f->blk_comp[c];
    f->blk_base[c] = f->plane_off[comp] + ((int64_t)f->blk_dy[c] * f->hblocks[comp] + f->blk_dx[c]) * 64;
    f->blk_xs[c] = f->hs[comp] * 64;
    f->blk_ys[c] = f->vs[comp] * f->hblocks[comp] * 64;
  }
}
