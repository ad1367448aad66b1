// TEST-ONLY host build of jgpu_idct_core.cuh (JGPU_CORE_HOST_EMULATION): checks
// the operation order of the device core against the oracle without a GPU.
// Not part of the product library; built by tests/test_host_core.py.
#define JGPU_CORE_HOST_EMULATION 1
#include <cmath>
#include "jgpu_idct_core.cuh"

extern "C" void core_idct_pairs(const short *in, short *out, long long npairs) {
  using namespace jgpu;
  for (long long p = 0; p < npairs; p++) {
    const short *a = in + 128 * p, *b = a + 64;
    pair32 m[8][8];
    for (int r = 0; r < 8; r++) {
      for (int c = 0; c < 8; c++) {
        m[r][c] = prescale(p_make((float)a[r * 8 + c], (float)b[r * 8 + c]), r, c);
      }
      inv_pass8(m[r]);
    }
    column_pass(m);
    for (int k = 0; k < 8; k++)
      for (int c = 0; c < 8; c++) {
        out[128 * p + k * 8 + c] = (short)std::floor(m[k][c].lo);
        out[128 * p + 64 + k * 8 + c] = (short)std::floor(m[k][c].hi);
      }
  }
}

// The fixed-point colour offsets the fused kernel uses (jgpu_colour_fixed.h), for every
// (Cb, Cr) in 0..255: out[(cb*256 + cr)*3 + {0,1,2}] = R, G, B offsets.
#include "jgpu_colour_fixed.h"
extern "C" void core_colour_fixed_all(int *out) {
  for (int cb = 0; cb < 256; cb++)
    for (int cr = 0; cr < 256; cr++) {
      int r, g, b;
      jgpu_colour_offsets_fixed(cb - 128, cr - 128, &r, &g, &b);
      int *o = out + (cb * 256 + cr) * 3;
      o[0] = r >> 16;
      o[1] = g >> 16;
      o[2] = b >> 16;
    }
}
