// Does VIADDMNMX.S16x2 wrap its intermediate sum?  nvcc -gencode arch=compute_100a,code=sm_100a -o viadd_probe viadd_probe.cu
#include <cstdio>
#include <cstdint>
__global__ void k(uint32_t *o) {
  const uint32_t a[4] = {0x7fff7fffu, 0x7f907f90u, 0x80008000u, 0x00640064u};
  for (int i = 0; i < 4; i++) {
    o[3 * i + 0] = __viaddmin_s16x2_relu(a[i], 0x00800080u, 0x00ff00ffu);
    o[3 * i + 1] = __viaddmax_s16x2(a[i], 0x00800080u, 0u);
    o[3 * i + 2] = __viaddmin_s16x2(a[i], 0xff80ff80u, 0x007f007fu);
  }
}
int main() {
  uint32_t *d, h[12];
  cudaMalloc(&d, sizeof(h));
  k<<<1, 1>>>(d);
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char *names[4] = {"0x7fff", "0x7f90", "0x8000", "0x0064"};
  for (int i = 0; i < 4; i++) printf("a=%s: relu(min(a+128,255))=%08x  max(a+128,0)=%08x  min(a-128,127)=%08x\n", names[i], h[3 * i], h[3 * i + 1], h[3 * i + 2]);
  return 0;
}
