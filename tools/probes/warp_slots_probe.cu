/* Which hardware warp slot (%warpid) does warp w of a CTA get when several CTAs share an SM?
 * The scheduler partition of a warp is believed to be %warpid % 4; the fused kernel's role
 * assignment depends on it.  Usage: warp_slots_probe <threads> <ctas_per_sm> */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
struct Rec { unsigned smid, warpid, cta, warp; };
__global__ void k(Rec *out, long long spin) {
  extern __shared__ char dummy[];
  unsigned smid, wid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  if ((threadIdx.x & 31) == 0) {
    Rec r = {smid, wid, blockIdx.x, threadIdx.x >> 5};
    out[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = r;
  }
  long long t0 = clock64();
  while (clock64() - t0 < spin) { }
  if (dummy[0] == 77 && spin < 0) out[0].smid = 0;
}
int main(int argc, char **argv) {
  int threads = argc > 1 ? atoi(argv[1]) : 192, per_sm = argc > 2 ? atoi(argv[2]) : 2;
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int smem = (220 * 1024 / per_sm) & ~1023;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int grid = sms * per_sm, warps = threads / 32;
  Rec *d, *h = (Rec *)malloc(sizeof(Rec) * grid * warps);
  cudaMalloc(&d, sizeof(Rec) * grid * warps);
  k<<<grid, threads, smem>>>(d, 2000000);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, d, sizeof(Rec) * grid * warps, cudaMemcpyDeviceToHost);
  printf("threads=%d ctas_per_sm=%d grid=%d\n", threads, per_sm, grid);
  for (int sm = 0; sm < 3; sm++) {
    printf("SM %d:", sm);
    for (int i = 0; i < grid * warps; i++)
      if ((int)h[i].smid == sm) printf("  cta%u.w%u->slot%u(p%u)", h[i].cta, h[i].warp, h[i].warpid, h[i].warpid & 3);
    printf("\n");
  }
  /* histogram of warps per partition, over all SMs */
  int worst[4] = {0, 0, 0, 0};
  for (int sm = 0; sm < sms; sm++) {
    int c[4] = {0, 0, 0, 0};
    for (int i = 0; i < grid * warps; i++) if ((int)h[i].smid == sm) c[h[i].warpid & 3]++;
    for (int p = 0; p < 4; p++) if (c[p] > worst[p]) worst[p] = c[p];
    if (sm < 3) printf("SM %d warps per partition: %d %d %d %d\n", sm, c[0], c[1], c[2], c[3]);
  }
  return 0;
}
