/* jgpu_unpack.cu — PACK (zero-run packed) stream -> dense QUANT planes on the device.
 *
 * What res/horz_pack_yuv.fs.glsl:105-127 / horz_pack_grey.fs.glsl do per fragment (fetch the
 * block's index, sign-extend the 12-bit DC word, walk run/level words until the end-of-block
 * word or coefficient 63, de-zigzag), done once per block; the output is the coefficient
 * layout of src/xjpeg.c:550-563 that the fused kernel reads.
 *
 * thread = one block.  It walks its words (2-byte loads; neighbouring blocks' words are
 * neighbours in the stream, so the lines are shared through L1/L2) and scatters the values
 * into a 128-byte row of shared memory; the warp then writes its 32 rows = 4 KB of contiguous
 * global memory with 128-bit stores.  Rows are stored with the 16-byte chunks XOR-swizzled by
 * the row number so that both the per-thread zero fill and the transposed read-out are free
 * of bank conflicts.
 *
 * Bytes per block: ~20 in (words + index), 128 out.  HBM-write bound; it exists to shrink
 * what crosses PCIe (3 B/px -> ~0.5 B/px for 4:2:0), not to speed up the device side.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "jgpu_launch.h"

namespace jgpu {

/* zig-zag position -> natural position (ITU-T T.81 figure A.6; the reference's DE_ZIG_ZAG,
 * res/horz_pack_yuv.fs.glsl:3-12) */
__constant__ uint8_t c_natural[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

__device__ __forceinline__ int sext12(uint32_t w) { return ((int)(w << 20)) >> 20; }

__global__ void __launch_bounds__(kUnpackThreads)
k_unpack(const UnpackSeg *__restrict__ segs, const UnpackWork *__restrict__ work,
         const uint16_t *__restrict__ pack, const int64_t *__restrict__ pack_off,
         const int32_t *__restrict__ index, int16_t *__restrict__ coef) {
  __shared__ __align__(16) uint8_t rows[kUnpackThreads * 128];
  __shared__ uint8_t natural[64];
  const UnpackWork w = work[blockIdx.x];
  const UnpackSeg s = segs[w.seg];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) natural[threadIdx.x] = c_natural[threadIdx.x];
  __syncthreads();

  uint8_t *const wrows = rows + warp * (32 * 128);
  uint8_t *const mine = wrows + lane * 128;
  const int sw = lane & 7;
  const int b = w.first + (int)threadIdx.x;
#pragma unroll
  for (int c = 0; c < 8; c++) *reinterpret_cast<uint4 *>(mine + 16 * (c ^ sw)) = make_uint4(0, 0, 0, 0);
  if (b < s.nblocks) {
    const int64_t lo = pack_off[s.img], hi = pack_off[s.img + 1];
    int64_t i = lo + (int64_t)(uint32_t)index[s.block0 + b];
    if (i < hi) {
      /* natural position n lives at chunk n>>3 (swizzled), element n&7 */
      *reinterpret_cast<int16_t *>(mine + 16 * (0 ^ sw)) = (int16_t)sext12(pack[i]);
      i++;
      int j = 0;
      while (j < 63 && i < hi) {
        const uint32_t p = pack[i++];
        if (p == 0) break;
        j += (int)(p >> 12) + 1;
        if (j > 63) break;   /* corrupt stream: the reference would index outside the block */
        const int n = natural[j];
        *reinterpret_cast<int16_t *>(mine + 16 * ((n >> 3) ^ sw) + 2 * (n & 7)) = (int16_t)sext12(p);
      }
    }
  }
  __syncwarp();
  /* the warp's rows are contiguous in global memory: chunk q of the warp is chunk q&7 of row q>>3 */
  const int first_row = w.first + 32 * warp;
  int16_t *const out = coef + (s.block0 + first_row) * 64;
#pragma unroll
  for (int it = 0; it < 8; it++) {
    const int q = 32 * it + lane, r = q >> 3, c = q & 7;
    if (first_row + r < s.nblocks) {
      const uint4 v = *reinterpret_cast<const uint4 *>(wrows + 128 * r + 16 * (c ^ (r & 7)));
      *reinterpret_cast<uint4 *>(out + 8 * q) = v;
    }
  }
}

cudaError_t launch_unpack(const UnpackSeg *segs, const UnpackWork *work, int ncta,
                          const uint16_t *pack, const int64_t *pack_off, const int32_t *index,
                          int16_t *coef, cudaStream_t stream) {
  if (ncta <= 0) return cudaSuccess;
  k_unpack<<<ncta, kUnpackThreads, 0, stream>>>(segs, work, pack, pack_off, index, coef);
  return cudaGetLastError();
}

}  // namespace jgpu
