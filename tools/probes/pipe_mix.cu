// Probe: issue rate of packed-FP32 (FMA pipe) and PRMT / VIADDMNMX (ALU pipe) streams, alone and mixed, by warps per scheduler.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix pipe_mix.cu && ./pipe_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x6240;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t vmm(uint32_t a, uint32_t b) { uint32_t r; asm volatile("vmin.s32.s32.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ float fadd1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <int MODE>
__global__ void k(u64 *out, int iters) {
  u64 f[8]; uint32_t p[8]; float s[8];
  for (int i = 0; i < 8; i++) { f[i] = threadIdx.x + i; p[i] = threadIdx.x * 3 + i; s[i] = (float)i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 8; rep++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MODE == 0 || MODE == 2 || MODE == 4) f[i] = fadd2(f[i], f[(i + 1) & 7]);
        if (MODE == 1 || MODE == 2) p[i] = prmt(p[i], p[(i + 1) & 7]);
        if (MODE == 3 || MODE == 4) p[i] = imad(p[i], p[(i + 1) & 7], p[(i + 2) & 7]);
        if (MODE == 5 || MODE == 6) s[i] = fadd1(s[i], s[(i + 1) & 7]);
        if (MODE == 6) p[i] = prmt(p[i], p[(i + 1) & 7]);
      }
    }
  }
  long long t1 = clock64();
  u64 acc = 0;
  for (int i = 0; i < 8; i++) acc += f[i] + p[i] + (u64)s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (u64)(t1 - t0);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1 << 20] = (u64)(t1 - t0);
}
template <int MODE> void run(const char *name, int per_iter) {
  u64 *out; cudaMalloc(&out, ((1 << 20) + 8) * 8);
  for (int warps = 4; warps <= 16; warps *= 2) {
    int iters = 2000;
    k<MODE><<<1, warps * 32>>>(out, iters); cudaDeviceSynchronize();
    k<MODE><<<1, warps * 32>>>(out, iters); cudaDeviceSynchronize();
    u64 cyc; cudaMemcpy(&cyc, out + (1 << 20), 8, cudaMemcpyDeviceToHost);
    double inst = (double)iters * 64 * per_iter;   // per warp
    printf("%-28s warps/scheduler %d: %.3f cycles per warp-instruction, scheduler IPC %.3f\n", name, warps / 4, cyc / inst, inst * (warps / 4) / cyc);
  }
  cudaFree(out);
}
int main() {
  run<0>("FADD2 only", 1); run<1>("PRMT only", 1); run<2>("FADD2 + PRMT alternating", 2); run<3>("IMAD only", 1); run<4>("FADD2 + IMAD alternating", 2);
  run<5>("FADD only", 1); run<6>("FADD + PRMT alternating", 2);
  return 0;
}
