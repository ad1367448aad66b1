/* jgpu_huff_prep.c — host-side preparation for the GPU entropy decoder (jgpu_huff_core.h):
 * decoder tables from the canonical description a DHT segment carries. */
#include <string.h>
#include "jgpu_huff_core.h"

/* T.81 C.2 (code generation) and F.2.2.3 (decoding by code length), laid out for a decoder
 * that looks at a left-aligned 16-bit window: the reference builds its own lookahead table
 * from the same description, src/xjpeg.c:311-336. */
int jgpu_huff_build_table(jgpu_huff_table *t, const unsigned char counts[16], const unsigned char *symbols) {
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
        for (s = 0; s < span; s++) t->lut[first + s] = (uint16_t)((len << 8) | symbols[k + i]);
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
