/* jgpu_fused.cu — the fused coefficient -> RGB kernel for sm_100a.
 *
 * One kernel does what the reference needs three render passes for
 * (res/horz_quant_*.fs.glsl -> res/vert.fs.glsl -> res/unyuv.fs.glsl /
 * ungrey.fs.glsl, driven by src/jpeg_gpu.c:1341-1363): dequantise, both IDCT
 * passes, bias/clamp, nearest-neighbour chroma upsample, colour matrix, crop,
 * RGB8 store.  Per coefficient sample the only DRAM traffic is the 2-byte read
 * and, per pixel, the 3-byte write.
 *
 * Work decomposition
 *   tile   = a run of MCUs of one MCU row of one image (512 px wide for colour
 *            modes).  CTAs are persistent and take tiles round-robin.
 *   thread = one PAIR of 8x8 blocks, 64 sample pairs in registers through both
 *            IDCT passes (jgpu_idct_core.cuh); the pair is two horizontally
 *            adjacent luma blocks, or the Cb and the Cr block of one MCU.
 *   warp   = 32 pairs of one block row: "Y warps" (one per luma block row of
 *            the MCU row) and "C warps" (chroma).
 *
 * Data movement
 *   HBM -> smem: TMA tensor loads (cp.async.bulk.tensor) issued by one thread,
 *     completion on an mbarrier, kStages-deep ring.  The coefficient buffer is
 *     described ONCE as rows of 128 bytes (one block per row); a box is 32 rows
 *     = 32 blocks, written with the 128-byte swizzle so that lane L reading
 *     16-byte chunk r of row L is bank-conflict free.  Luma uses a 3-D view
 *     (64, parity, pair) of the same memory so that one box gathers the 32
 *     even (or the 32 odd) blocks of 64 consecutive blocks: lane L then owns
 *     blocks 2L and 2L+1, i.e. 16 adjacent output pixels.
 *   C warps -> Y warps: per chroma sample the three integer colour offsets,
 *     already laid out as the s16x2 operands the Y threads need, through a
 *     padded shared-memory exchange buffer and one named barrier.
 *   regs -> HBM: each Y thread holds 16 adjacent pixels of a row = 48 bytes =
 *     three 128-bit stores; a warp covers 1536 contiguous bytes per row.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "jgpu_internal.h"
#include "jgpu_kernels.cuh"
#include "jgpu_launch.h"

namespace jgpu {

constexpr int kBoxRows = 32;                 /* blocks per TMA box */
constexpr int kBoxBytes = kBoxRows * 128;    /* 4 KB */
constexpr int kStageTail = 1024;             /* tables + header, keeps boxes 1 KB aligned */
constexpr int kQtabBytes = 64 * 4;           /* one table as int32 */

struct __align__(16) StageHeader {
  long long rgb_base;   /* byte offset in the rgb buffer of the tile's top-left pixel */
  int32_t width_left;   /* visible pixels from the tile's left edge to the image's right edge */
  int32_t rows_left;    /* visible rows from the tile's top row to the image's bottom */
  int32_t pitch;        /* bytes per output row */
  int32_t flags;        /* bit 0: output rows are 16-byte aligned */
  int32_t pad[2];
};

template <int HS, int VS, bool GRAY>
struct Cfg {
  static constexpr int kYWarps = GRAY ? 3 : VS;
  static constexpr int kCWarps = GRAY ? 0 : 2 / HS;
  static constexpr int kWarps = kYWarps + kCWarps;
  static constexpr int kThreads = 32 * kWarps;
  /* MCUs per tile */
  static constexpr int kTileMcus = GRAY ? 192 : (HS == 2 ? 32 : 64);
  static constexpr int kMcuW = GRAY ? 8 : 8 * HS;
  static constexpr int kMcuH = GRAY ? 8 : 8 * VS;
  static constexpr int kBoxes = 2 * kYWarps + 2 * kCWarps;
  static constexpr int kStageBytes = kBoxes * kBoxBytes + kStageTail;
  static constexpr int kTables = GRAY ? 1 : 3;
  /* exchange buffer: colour offsets of one chroma block, 8 rows */
  static constexpr int kExRow = HS == 2 ? 96 : 48;        /* bytes per chroma row */
  static constexpr int kExTask = 8 * kExRow + 16;          /* padded: odd multiple of 16 */
  static constexpr int kExRegion1 = 32 * kExTask + 64;     /* HS==1: odd MCUs live here */
  static constexpr int kExBytes = GRAY ? 0 : (HS == 2 ? 32 * kExTask : 2 * 32 * kExTask + 64 + 64);
  static constexpr int kChannels = GRAY ? 1 : 3;
  /* resident CTAs per SM the register budget is sized for (smem allows no more) */
  static constexpr int kMinCtas = kThreads <= 64 ? 4 : (kThreads <= 96 ? 3 : 2);
};

/* ---- PTX wrappers ---------------------------------------------------------- */

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
/* Bounded wait: a pipeline bug must trap, not hang the GPU. */
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++) {
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1,
                                            int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes,
                                          uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void named_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int threads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg128_stream(uint8_t *p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

/* ---- per-thread stages ----------------------------------------------------- */

/* Row pass for one block pair out of two swizzled TMA boxes.  `row` is this
 * lane's row inside the boxes; chunk r of a 128-byte row sits at 16*(r ^ (row&7)).
 * qa / qb: the int32 quantisation tables of block A / block B (natural order). */
__device__ __forceinline__ void pair_row_pass(pair32 (&m)[8][8], const uint8_t *box_a,
                                              const uint8_t *box_b, int row, const int *qa,
                                              const int *qb) {
  const uint8_t *ra = box_a + 128 * row, *rb = box_b + 128 * row;
  const int sw = row & 7;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int off = 16 * (r ^ sw);
    const uint4 a = *reinterpret_cast<const uint4 *>(ra + off);
    const uint4 b = *reinterpret_cast<const uint4 *>(rb + off);
    load_row_pair(m[r], a, b, qa + 8 * r, qb + 8 * r, r);
    inv_pass8(m[r]);
  }
}

/* Chroma sample pair (Cb in .lo, Cr in .hi, un-floored) -> the raw bits of
 * RN(offset + 1.5*2^23) for R, G, B (low 16 bits = the integer offset).
 * Arithmetic is colour_offsets() of jgpu_kernels.cuh, i.e. the oracle's. */
__device__ __forceinline__ void chroma_offsets_bits(pair32 v, uint32_t &rb, uint32_t &gb,
                                                    uint32_t &bb) {
  const pair32 magic = p_make_bits(kMagicBits, kMagicBits);
  uint32_t cbits, rbits;
  p_split_bits(p_add_rm(v, magic), cbits, rbits);
  uint32_t s = __byte_perm(cbits, rbits, 0x5410);   /* (short)floor of Cb | Cr */
  s = __viaddmin_s16x2(s, 0u, 0x007f007fu);         /* clamp(v+128,0,255)-128 == clamp(v,-128,127) */
  s = __viaddmax_s16x2(s, 0u, 0xff80ff80u);
  const float cbf = (float)(short)(s & 0xffffu);
  const float crf = (float)((int)s >> 16);
  const float fm = __uint_as_float(kMagicBits);
  const float rc = __fmul_rn(1.402f, crf);
  const float gc = __fadd_rn(__fmul_rn(-0.34414f, cbf), __fmul_rn(-0.71414f, crf));
  const float bc = __fmul_rn(1.772f, cbf);
  rb = __float_as_uint(__fadd_rn(rc, fm));
  gb = __float_as_uint(__fadd_rn(gc, fm));
  bb = __float_as_uint(__fadd_rn(bc, fm));
}

/* Four pixels: Y pairs (ya, yb) + colour-offset words -> 12 RGB bytes. */
__device__ __forceinline__ void rgb4(uint32_t ya, uint32_t yb, uint32_t ra, uint32_t ga,
                                     uint32_t ba, uint32_t rb, uint32_t gb, uint32_t bb,
                                     uint32_t &w0, uint32_t &w1, uint32_t &w2) {
  const uint32_t lim = 0x00ff00ffu;
  const uint32_t Ra = __viaddmin_s16x2_relu(ya, ra, lim);
  const uint32_t Ga = __viaddmin_s16x2_relu(ya, ga, lim);
  const uint32_t Ba = __viaddmin_s16x2_relu(ya, ba, lim);
  const uint32_t Rb = __viaddmin_s16x2_relu(yb, rb, lim);
  const uint32_t Gb = __viaddmin_s16x2_relu(yb, gb, lim);
  const uint32_t Bb = __viaddmin_s16x2_relu(yb, bb, lim);
  const uint32_t t = __byte_perm(Ra, Ga, 0x6240);   /* R0 G0 R1 G1 */
  const uint32_t u = __byte_perm(Rb, Gb, 0x6240);   /* R2 G2 R3 G3 */
  const uint32_t x = __byte_perm(t, Ba, 0x0063);    /* G1 B1 .  .  */
  w0 = __byte_perm(t, Ba, 0x2410);                  /* R0 G0 B0 R1 */
  w1 = __byte_perm(x, u, 0x5410);                   /* G1 B1 R2 G2 */
  w2 = __byte_perm(u, Bb, 0x6324);                  /* B2 R3 G3 B3 */
}

/* Stores `nbytes` (<= 48) of one output row segment. */
__device__ __forceinline__ void store_row(uint8_t *dst, const uint32_t (&w)[12], int nbytes,
                                          bool fast) {
  if (fast && nbytes == 48) {
    stg128_stream(dst, make_uint4(w[0], w[1], w[2], w[3]));
    stg128_stream(dst + 16, make_uint4(w[4], w[5], w[6], w[7]));
    stg128_stream(dst + 32, make_uint4(w[8], w[9], w[10], w[11]));
  } else {
#pragma unroll
    for (int i = 0; i < 12; i++) {
#pragma unroll
      for (int b = 0; b < 4; b++) {
        if (4 * i + b < nbytes) dst[4 * i + b] = (uint8_t)(w[i] >> (8 * b));
      }
    }
  }
}

/* ---- the kernel ------------------------------------------------------------ */

/* What every role needs to know about the launch; lives in registers. */
struct TileLoop {
  uint32_t smem0;      /* shared-space address of stage 0 (1 KB aligned) */
  uint8_t *smem_gen;   /* the same location as a generic pointer */
  uint32_t ex0;        /* exchange buffer */
  uint32_t bar0;       /* mbarriers, one per stage */
  int n_tiles;
};

/* Tile header of stage `s`, as the producer wrote it. */
struct TileView {
  long long rgb_base;
  int width_left, rows_left, pitch;
  bool fast;
};

template <typename C>
__device__ __forceinline__ TileView read_header(const TileLoop &L, int s) {
  const uint32_t a = L.smem0 + s * C::kStageBytes + C::kBoxes * kBoxBytes + 3 * kQtabBytes;
  const uint4 h0 = lds128(a), h1 = lds128(a + 16);
  TileView v;
  v.rgb_base = (long long)(((unsigned long long)h0.y << 32) | h0.x);
  v.width_left = (int)h0.z;
  v.rows_left = (int)h0.w;
  v.pitch = (int)h1.x;
  v.fast = (h1.y & 1u) != 0;
  return v;
}

/* C warps: Cb/Cr block pair -> colour offsets in the exchange buffer. */
template <int HS, int VS, bool GRAY, int STAGES>
__device__ __forceinline__ void chroma_role(const TileLoop &L, int cw, int lane) {
  using C = Cfg<HS, VS, GRAY>;
  int it = 0;
  for (int tile = blockIdx.x; tile < L.n_tiles; tile += gridDim.x, it++) {
    const int s = it % STAGES;
    mbar_wait(L.bar0 + 8 * s, (uint32_t)(it / STAGES) & 1u);
    const TileView tv = read_header<C>(L, s);
    const uint8_t *stp = L.smem_gen + s * C::kStageBytes;
    const int *qtp = reinterpret_cast<const int *>(stp + C::kBoxes * kBoxBytes);
    const int mcu = 32 * cw + lane;
    const bool active = (HS == 2 ? 16 : 8) * mcu < tv.width_left;
    pair32 m[8][8];
    if (active) {
      pair_row_pass(m, stp + (2 * C::kYWarps + cw) * kBoxBytes,
                    stp + (2 * C::kYWarps + C::kCWarps + cw) * kBoxBytes, lane, qtp + 64, qtp + 128);
    }
    named_sync(1, C::kThreads);   /* stage s consumed */
    if (active) {
      column_pass(m);
      const uint32_t base = HS == 2 ? L.ex0 + lane * C::kExTask
                                    : L.ex0 + (mcu & 1) * C::kExRegion1 + (mcu >> 1) * C::kExTask;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        uint32_t rb[8], gb[8], bb[8];
#pragma unroll
        for (int c = 0; c < 8; c++) chroma_offsets_bits(m[k][c], rb[c], gb[c], bb[c]);
        if (HS == 2) {
          /* each chroma sample serves a horizontal pixel pair: replicate */
#pragma unroll
          for (int v = 0; v < 6; v++) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int word = 4 * v + i, c = word / 3, ch = word % 3;
              const uint32_t src = ch == 0 ? rb[c] : (ch == 1 ? gb[c] : bb[c]);
              w[i] = __byte_perm(src, src, 0x1010);
            }
            sts128(base + k * 96 + 16 * v, make_uint4(w[0], w[1], w[2], w[3]));
          }
        } else {
#pragma unroll
          for (int v = 0; v < 3; v++) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int word = 4 * v + i, j = word / 3, ch = word % 3;
              const uint32_t s0 = ch == 0 ? rb[2 * j] : (ch == 1 ? gb[2 * j] : bb[2 * j]);
              const uint32_t s1 = ch == 0 ? rb[2 * j + 1] : (ch == 1 ? gb[2 * j + 1] : bb[2 * j + 1]);
              w[i] = __byte_perm(s0, s1, 0x5410);
            }
            sts128(base + k * 48 + 16 * v, make_uint4(w[0], w[1], w[2], w[3]));
          }
        }
      }
    }
    named_arrive(2, C::kThreads);   /* offsets of this tile are in the exchange buffer */
  }
}

/* Y warps: luma block pair -> 16 x 8 pixels of RGB (or grey). */
template <int HS, int VS, bool GRAY, int STAGES, typename Issue>
__device__ __forceinline__ void luma_role(const TileLoop &L, int wy, int lane, uint8_t *rgb,
                                          Issue &&issue) {
  using C = Cfg<HS, VS, GRAY>;
  int it = 0;
  for (int tile = blockIdx.x; tile < L.n_tiles; tile += gridDim.x, it++) {
    const int s = it % STAGES;
    mbar_wait(L.bar0 + 8 * s, (uint32_t)(it / STAGES) & 1u);
    const TileView tv = read_header<C>(L, s);
    const uint8_t *stp = L.smem_gen + s * C::kStageBytes;
    const int *qtp = reinterpret_cast<const int *>(stp + C::kBoxes * kBoxBytes);
    const int px_x = GRAY ? 512 * wy + 16 * lane : 16 * lane;
    const int px_y = GRAY ? 0 : 8 * wy;
    const bool active = px_x < tv.width_left && px_y < tv.rows_left;

    uint32_t ya[8][4], yb[8][4];
    {
      pair32 m[8][8];
      if (active) {
        pair_row_pass(m, stp + (2 * wy) * kBoxBytes, stp + (2 * wy + 1) * kBoxBytes, lane, qtp, qtp);
      }
      named_sync(1, C::kThreads);   /* every warp has consumed stage s */
      if (active) {
        /* (short)floor + 128, clamp, two columns at a time as the column pass
         * produces them: ya[k][j] / yb[k][j] = pixels 2j, 2j+1 of row k of
         * block A / block B as s16x2 */
        const pair32 magic = p_make_bits(kMagicBits, kMagicBits);
        column_pass_by_pairs(m, [&](int j, pair32 (&u)[8], pair32 (&v)[8]) {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            uint32_t ulo, uhi, vlo, vhi;
            p_split_bits(p_add_rm(u[k], magic), ulo, uhi);
            p_split_bits(p_add_rm(v[k], magic), vlo, vhi);
            ya[k][j] = clamp_pair_u8(ulo, vlo);
            yb[k][j] = clamp_pair_u8(uhi, vhi);
          }
        });
      }
    }
    /* refill stage s with the tile STAGES ahead (after the column pass: the 64
     * sample pairs are dead by now, so the producer's state costs no registers) */
    if (threadIdx.x == 0) {
      const int nt = tile + STAGES * gridDim.x;
      if (nt < L.n_tiles) issue(nt, s);
    }
    if (!GRAY) named_sync(2, C::kThreads);   /* colour offsets of this tile are ready */
    if (!active) continue;

    uint8_t *out = rgb + tv.rgb_base + (long long)px_y * tv.pitch + (long long)px_x * C::kChannels;
    const int vis_px = min(16, tv.width_left - px_x);
    const int vis_rows = min(8, tv.rows_left - px_y);
    if (GRAY) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (k < vis_rows) {
          uint4 v;
          v.x = __byte_perm(ya[k][0], ya[k][1], 0x6420);
          v.y = __byte_perm(ya[k][2], ya[k][3], 0x6420);
          v.z = __byte_perm(yb[k][0], yb[k][1], 0x6420);
          v.w = __byte_perm(yb[k][2], yb[k][3], 0x6420);
          uint8_t *dst = out + (long long)k * tv.pitch;
          if (tv.fast && vis_px == 16) {
            stg128_stream(dst, v);
          } else {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 16; i++) {
              if (i < vis_px) dst[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
            }
          }
        }
      }
    } else {
      uint32_t ca[12], cb[12];   /* offsets for block A / block B: 4 pixel pairs x (R,G,B) */
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (k % VS == 0) {        /* VS pixel rows share one chroma row */
          const int crow = (8 * wy + k) / VS;
          const uint32_t a = L.ex0 + lane * C::kExTask + crow * C::kExRow;
          const uint32_t b = HS == 2 ? a + 48 : a + C::kExRegion1;
#pragma unroll
          for (int v = 0; v < 3; v++) {
            const uint4 t0 = lds128(a + 16 * v), t1 = lds128(b + 16 * v);
            ca[4 * v] = t0.x; ca[4 * v + 1] = t0.y; ca[4 * v + 2] = t0.z; ca[4 * v + 3] = t0.w;
            cb[4 * v] = t1.x; cb[4 * v + 1] = t1.y; cb[4 * v + 2] = t1.z; cb[4 * v + 3] = t1.w;
          }
        }
        if (k < vis_rows) {
          uint32_t w[12];
          rgb4(ya[k][0], ya[k][1], ca[0], ca[1], ca[2], ca[3], ca[4], ca[5], w[0], w[1], w[2]);
          rgb4(ya[k][2], ya[k][3], ca[6], ca[7], ca[8], ca[9], ca[10], ca[11], w[3], w[4], w[5]);
          rgb4(yb[k][0], yb[k][1], cb[0], cb[1], cb[2], cb[3], cb[4], cb[5], w[6], w[7], w[8]);
          rgb4(yb[k][2], yb[k][3], cb[6], cb[7], cb[8], cb[9], cb[10], cb[11], w[9], w[10], w[11]);
          store_row(out + (long long)k * tv.pitch, w, 3 * vis_px, tv.fast);
        }
      }
    }
  }
}

template <int HS, int VS, bool GRAY, int STAGES>
__global__ void __launch_bounds__(Cfg<HS, VS, GRAY>::kThreads, Cfg<HS, VS, GRAY>::kMinCtas)
k_fused(const __grid_constant__ CUtensorMap tm_rows,   /* (64, rows)            */
        const __grid_constant__ CUtensorMap tm_pairs,  /* (64, parity, pairs)   */
        const FusedImage *__restrict__ images, const TileRef *__restrict__ tiles, int n_tiles,
        const int32_t *__restrict__ qint, uint8_t *__restrict__ rgb) {
  using C = Cfg<HS, VS, GRAY>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TileLoop L;
  /* dynamic smem is only guaranteed 16-byte aligned: round up to 1 KB for the swizzle */
  L.smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  L.smem_gen = smem_raw + (L.smem0 - smem_u32(smem_raw));
  L.ex0 = L.smem0 + STAGES * C::kStageBytes;
  L.bar0 = L.ex0 + ((C::kExBytes + 15) & ~15);
  L.n_tiles = n_tiles;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(L.bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  /* Producer side (thread 0 only): describe tile `t` in stage `s` and start its loads. */
  auto issue = [&](int t, int s) {
    const TileRef tr = tiles[t];
    const FusedImage im = images[tr.img];
    const uint32_t st = L.smem0 + s * C::kStageBytes;
    const uint32_t bar = L.bar0 + 8 * s;
    const int x0 = tr.mx0 * C::kMcuW, y0 = tr.mrow * C::kMcuH;
    StageHeader h;
    h.pitch = im.width * C::kChannels;
    h.rgb_base = im.rgb_off + ((long long)y0 * im.width + x0) * C::kChannels;
    h.width_left = im.width - x0;
    h.rows_left = im.height - y0;
    h.flags = (((reinterpret_cast<uintptr_t>(rgb) + (uintptr_t)h.rgb_base) & 15) == 0 &&
               (h.pitch & 15) == 0) ? 1 : 0;
    h.pad[0] = h.pad[1] = 0;
    *reinterpret_cast<StageHeader *>(L.smem_gen + s * C::kStageBytes + C::kBoxes * kBoxBytes +
                                     3 * kQtabBytes) = h;
    mbar_expect_tx(bar, C::kBoxes * kBoxBytes + C::kTables * kQtabBytes);
    /* luma boxes: even blocks then odd blocks of each Y warp's 64-block run */
#pragma unroll
    for (int wy = 0; wy < C::kYWarps; wy++) {
      int first;
      if (GRAY) {
        first = im.block0[0] + tr.mrow * im.hblocks[0] + tr.mx0 + 64 * wy;
      } else {
        first = im.block0[0] + (tr.mrow * VS + wy) * im.hblocks[0] + tr.mx0 * HS;
      }
      tma_load_3d(st + (2 * wy) * kBoxBytes, &tm_pairs, 0, first & 1, first >> 1, bar);
      tma_load_3d(st + (2 * wy + 1) * kBoxBytes, &tm_pairs, 0, (first + 1) & 1, (first + 1) >> 1, bar);
    }
#pragma unroll
    for (int cw = 0; cw < C::kCWarps; cw++) {
      const int fb = im.block0[1] + tr.mrow * im.hblocks[1] + tr.mx0 + 32 * cw;
      const int fr = im.block0[2] + tr.mrow * im.hblocks[2] + tr.mx0 + 32 * cw;
      tma_load_2d(st + (2 * C::kYWarps + cw) * kBoxBytes, &tm_rows, 0, fb, bar);
      tma_load_2d(st + (2 * C::kYWarps + C::kCWarps + cw) * kBoxBytes, &tm_rows, 0, fr, bar);
    }
#pragma unroll
    for (int c = 0; c < C::kTables; c++) {
      bulk_load(st + C::kBoxes * kBoxBytes + c * kQtabBytes, qint + (size_t)im.qidx[c] * 64,
                kQtabBytes, bar);
    }
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < n_tiles) issue(t, s);
    }
  }

  if (warp < C::kYWarps) {
    luma_role<HS, VS, GRAY, STAGES>(L, warp, lane, rgb, issue);
  } else {
    chroma_role<HS, VS, GRAY, STAGES>(L, warp - C::kYWarps, lane);
  }
}

/* ---- host side ------------------------------------------------------------- */

/* u16 tables -> int32 tables (one tiny launch per run) */
__global__ void k_prep_qtabs(const uint16_t *__restrict__ q, int32_t *__restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = q[i];
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

constexpr int kDefaultStages = 2;

template <int HS, int VS, bool GRAY>
size_t smem_bytes() {
  using C = Cfg<HS, VS, GRAY>;
  return 1024 + kDefaultStages * C::kStageBytes + ((C::kExBytes + 15) & ~15) + 8 * kDefaultStages + 16;
}

struct ModeInfo {
  int tile_mcus, mcu_w, mcu_h, threads;
  size_t smem;
  const void *func;
  int ctas_per_sm;
};

ModeInfo g_modes[kNumFusedModes];
bool g_configured = false;

template <int HS, int VS, bool GRAY>
cudaError_t configure_mode(int mode) {
  using C = Cfg<HS, VS, GRAY>;
  auto *f = &k_fused<HS, VS, GRAY, kDefaultStages>;
  ModeInfo &mi = g_modes[mode];
  mi.tile_mcus = C::kTileMcus;
  mi.mcu_w = C::kMcuW;
  mi.mcu_h = C::kMcuH;
  mi.threads = C::kThreads;
  mi.smem = smem_bytes<HS, VS, GRAY>();
  mi.func = reinterpret_cast<const void *>(f);
  cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.smem);
  if (e != cudaSuccess) return e;
  int n = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, f, C::kThreads, mi.smem);
  if (e != cudaSuccess) return e;
  mi.ctas_per_sm = std::max(n, 1);
  return cudaSuccess;
}

template <int HS, int VS, bool GRAY>
cudaError_t launch_mode(int grid, size_t smem, cudaStream_t stream, const CUtensorMap &tm_rows,
                        const CUtensorMap &tm_pairs, const FusedImage *images, const TileRef *tiles,
                        int n_tiles, const int32_t *qint, uint8_t *rgb) {
  using C = Cfg<HS, VS, GRAY>;
  k_fused<HS, VS, GRAY, kDefaultStages><<<grid, C::kThreads, smem, stream>>>(
      tm_rows, tm_pairs, images, tiles, n_tiles, qint, rgb);
  return cudaGetLastError();
}

}  // namespace

bool fused_available() { return true; }

cudaError_t fused_configure(int device) {
  (void)device;
  cudaError_t e;
  if ((e = configure_mode<1, 1, true>(kModeGray)) != cudaSuccess) return e;
  if ((e = configure_mode<1, 1, false>(kMode444)) != cudaSuccess) return e;
  if ((e = configure_mode<2, 1, false>(kMode422)) != cudaSuccess) return e;
  if ((e = configure_mode<2, 2, false>(kMode420)) != cudaSuccess) return e;
  if ((e = configure_mode<1, 2, false>(kMode440)) != cudaSuccess) return e;
  g_configured = true;
  return cudaSuccess;
}

struct FusedPlanImpl {
  int n = 0;
  int sm_count = 0;
  void *d_images = nullptr;
  void *d_tiles[kNumFusedModes] = {};
  int n_tiles[kNumFusedModes] = {};
  std::vector<int> first_tile[kNumFusedModes]; /* per mode, n+1 entries */
  void *d_qint = nullptr;
  int qint_cap = 0; /* tables */
  long long coef_rows = 0; /* 128-byte rows the batch touches */
  /* tensor maps, rebuilt when the coefficient pointer changes */
  const void *map_ptr = nullptr;
  CUtensorMap tm_rows, tm_pairs;
};

int fused_plan_build(FusedPlan &fp, const jgpu_image_desc *descs, const jgpu_layout *layouts,
                     const int *modes, int n, unsigned flags, int sm_count) {
  if (!g_configured) return jgpu_fail("fused kernel not configured");
  if (!encode_fn()) return jgpu_fail("cuTensorMapEncodeTiled is not available from this driver");
  FusedPlanImpl *p = new FusedPlanImpl();
  fp.impl = p;
  p->n = n;
  p->sm_count = sm_count;
  (void)flags;
  std::vector<FusedImage> images(n);
  std::vector<TileRef> tiles[kNumFusedModes];
  for (int m = 0; m < kNumFusedModes; m++) p->first_tile[m].assign(n + 1, 0);
  for (int i = 0; i < n; i++) {
    const jgpu_image_desc &d = descs[i];
    const jgpu_layout &lay = layouts[i];
    FusedImage &im = images[i];
    memset(&im, 0, sizeof(im));
    im.rgb_off = d.rgb_off;
    im.width = d.width;
    im.height = d.height;
    for (int c = 0; c < d.ncomps; c++) {
      im.block0[c] = (int32_t)((d.coef_off + lay.plane[c].coef_off) / 64);
      im.hblocks[c] = lay.plane[c].hblocks;
      im.qidx[c] = d.qtab_set * 4 + d.tq[c];
    }
    p->coef_rows = std::max<long long>(p->coef_rows, (d.coef_off + lay.coef_len + 63) / 64);
    const ModeInfo &mi = g_modes[modes[i]];
    for (int m = 0; m < kNumFusedModes; m++) p->first_tile[m][i] = (int)tiles[m].size();
    const int nh = modes[i] == kModeGray ? lay.plane[0].hblocks : lay.nhmb;
    const int nv = modes[i] == kModeGray ? lay.plane[0].vblocks : lay.nvmb;
    for (int r = 0; r < nv; r++) {
      for (int x = 0; x < nh; x += mi.tile_mcus) {
        /* tiles wholly to the right of / below the visible image carry no pixels */
        if (x * mi.mcu_w >= d.width || r * mi.mcu_h >= d.height) continue;
        TileRef t = {i, (int16_t)r, (int16_t)x};
        tiles[modes[i]].push_back(t);
      }
    }
  }
  for (int m = 0; m < kNumFusedModes; m++) p->first_tile[m][n] = (int)tiles[m].size();
  if (cudaMalloc(&p->d_images, sizeof(FusedImage) * n) != cudaSuccess ||
      cudaMemcpy(p->d_images, images.data(), sizeof(FusedImage) * n, cudaMemcpyHostToDevice) != cudaSuccess) {
    return jgpu_fail("fused plan: image table upload failed");
  }
  for (int m = 0; m < kNumFusedModes; m++) {
    p->n_tiles[m] = (int)tiles[m].size();
    if (tiles[m].empty()) continue;
    if (cudaMalloc(&p->d_tiles[m], sizeof(TileRef) * tiles[m].size()) != cudaSuccess ||
        cudaMemcpy(p->d_tiles[m], tiles[m].data(), sizeof(TileRef) * tiles[m].size(),
                   cudaMemcpyHostToDevice) != cudaSuccess) {
      return jgpu_fail("fused plan: tile list upload failed");
    }
  }
  return 0;
}

void fused_plan_release(FusedPlan &fp) {
  FusedPlanImpl *p = static_cast<FusedPlanImpl *>(fp.impl);
  if (!p) return;
  cudaFree(p->d_images);
  for (int m = 0; m < kNumFusedModes; m++) cudaFree(p->d_tiles[m]);
  cudaFree(p->d_qint);
  delete p;
  fp.impl = nullptr;
}

int fused_plan_launches(const FusedPlan &fp) {
  const FusedPlanImpl *p = static_cast<const FusedPlanImpl *>(fp.impl);
  int n = 1; /* table conversion */
  for (int m = 0; m < kNumFusedModes; m++) n += p->n_tiles[m] > 0;
  return n;
}

static int build_maps(FusedPlanImpl *p, const int16_t *coef) {
  if (p->map_ptr == coef) return 0;
  if (reinterpret_cast<uintptr_t>(coef) & 255) {
    return jgpu_fail("fused path: the coefficient buffer must be 256-byte aligned");
  }
  EncodeTiledFn enc = encode_fn();
  const cuuint64_t rows = (cuuint64_t)((p->coef_rows + 1) & ~1ll);
  {
    cuuint64_t dims[2] = {64, rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, kBoxRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p->tm_rows, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<int16_t *>(coef), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return jgpu_fail("cuTensorMapEncodeTiled(rows) failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[3] = {64, 2, rows / 2};
    cuuint64_t strides[2] = {128, 256};
    cuuint32_t box[3] = {64, 1, kBoxRows};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&p->tm_pairs, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<int16_t *>(coef), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return jgpu_fail("cuTensorMapEncodeTiled(pairs) failed (%d)", (int)r);
  }
  p->map_ptr = coef;
  return 0;
}

int fused_plan_launch(FusedPlan &fp, int i0, int i1, const int16_t *coef, const uint16_t *qtabs,
                      int n_sets, uint8_t *rgb, cudaStream_t stream) {
  FusedPlanImpl *p = static_cast<FusedPlanImpl *>(fp.impl);
  if (build_maps(p, coef)) return 1;
  const int n_tables = n_sets * 4;
  if (n_tables > p->qint_cap) {
    cudaFree(p->d_qint);
    p->d_qint = nullptr;
    if (cudaMalloc(&p->d_qint, (size_t)n_tables * 64 * 4) != cudaSuccess) {
      return jgpu_fail("fused path: table buffer allocation failed");
    }
    p->qint_cap = n_tables;
  }
  k_prep_qtabs<<<(n_tables * 64 + 255) / 256, 256, 0, stream>>>(qtabs, (int32_t *)p->d_qint, n_tables * 64);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return jgpu_fail("k_prep_qtabs launch failed (%s)", cudaGetErrorString(e));
  for (int m = 0; m < kNumFusedModes; m++) {
    const int t0 = p->first_tile[m][i0], t1 = p->first_tile[m][i1];
    if (t1 <= t0) continue;
    const ModeInfo &mi = g_modes[m];
    const int grid = std::min(t1 - t0, p->sm_count * mi.ctas_per_sm);
    const TileRef *tiles = static_cast<const TileRef *>(p->d_tiles[m]) + t0;
    const FusedImage *images = static_cast<const FusedImage *>(p->d_images);
    const int32_t *qint = static_cast<const int32_t *>(p->d_qint);
    switch (m) {
      case kModeGray: e = launch_mode<1, 1, true>(grid, mi.smem, stream, p->tm_rows, p->tm_pairs, images, tiles, t1 - t0, qint, rgb); break;
      case kMode444: e = launch_mode<1, 1, false>(grid, mi.smem, stream, p->tm_rows, p->tm_pairs, images, tiles, t1 - t0, qint, rgb); break;
      case kMode422: e = launch_mode<2, 1, false>(grid, mi.smem, stream, p->tm_rows, p->tm_pairs, images, tiles, t1 - t0, qint, rgb); break;
      case kMode420: e = launch_mode<2, 2, false>(grid, mi.smem, stream, p->tm_rows, p->tm_pairs, images, tiles, t1 - t0, qint, rgb); break;
      case kMode440: e = launch_mode<1, 2, false>(grid, mi.smem, stream, p->tm_rows, p->tm_pairs, images, tiles, t1 - t0, qint, rgb); break;
    }
    if (e != cudaSuccess) return jgpu_fail("fused kernel launch failed (%s)", cudaGetErrorString(e));
  }
  return 0;
}

}  // namespace jgpu
