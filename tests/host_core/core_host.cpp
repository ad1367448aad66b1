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
