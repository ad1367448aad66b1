// Where do the CTAs of a one-CTA-per-SM persistent grid land?  Each CTA records its SM id and
// its start / end times (globaltimer); the host prints how many CTAs each SM hosted and whether
// they overlapped in time.  Usage: cta_placement_probe [smem_bytes] [threads] [grid] [spin_us]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
__global__ void probe(unsigned long long *out, long long spin_ns) {
  extern __shared__ unsigned char sm[];
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (threadIdx.x == 0) sm[0] = 1;
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while ((long long)(t1 - t0) < spin_ns);
  if (threadIdx.x == 0) { out[3 * blockIdx.x] = smid; out[3 * blockIdx.x + 1] = t0; out[3 * blockIdx.x + 2] = t1; }
}
int main(int argc, char **argv) {
  int smem = argc > 1 ? atoi(argv[1]) : 231168, threads = argc > 2 ? atoi(argv[2]) : 384;
  int grid = argc > 3 ? atoi(argv[3]) : 148; long long spin = (argc > 4 ? atoll(argv[4]) : 200) * 1000;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  unsigned long long *d; cudaMalloc(&d, 24 * grid);
  for (int rep = 0; rep < 3; rep++) {
    probe<<<grid, threads, smem>>>(d, spin);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<unsigned long long> h(3 * grid); cudaMemcpy(h.data(), d, 24 * grid, cudaMemcpyDeviceToHost);
    unsigned long long tmin = ~0ull, tmax = 0; std::vector<int> per(256, 0);
    for (int i = 0; i < grid; i++) { per[h[3*i]]++; tmin = std::min(tmin, h[3*i+1]); tmax = std::max(tmax, h[3*i+2]); }
    int used = 0, multi = 0, late = 0;
    for (int s = 0; s < 256; s++) { used += per[s] > 0; multi += per[s] > 1; }
    for (int i = 0; i < grid; i++) late += (h[3*i+1] - tmin) > (unsigned long long)spin / 2;
    printf("rep %d: smem %d threads %d grid %d: %d SMs used, %d SMs hosted >1 CTA, %d CTAs started late, span %.1f us (spin %.1f us)\n",
           rep, smem, threads, grid, used, multi, late, (tmax - tmin) / 1e3, spin / 1e3);
  }
  return 0;
}
