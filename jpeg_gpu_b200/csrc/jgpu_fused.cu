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
 *   warp   = 32 pairs of one block row: "luma warps" (one per luma block row
 *            of the MCU row) and "chroma warps".
 *
 * Data movement
 *   HBM -> smem: TMA tensor loads (cp.async.bulk.tensor).  The coefficient
 *     buffer is described ONCE as rows of 128 bytes (one block per row); a box
 *     is 32 rows = 32 blocks, written with the 128-byte swizzle so that lane L
 *     reading 16-byte chunk r of row L is bank-conflict free.  Luma uses a 3-D
 *     view (64, parity, pair) of the same memory so that one box gathers the
 *     32 even (or the 32 odd) blocks of 64 consecutive blocks: lane L then owns
 *     blocks 2L and 2L+1, i.e. 16 adjacent output pixels.
 *     Every WARP owns its two boxes, its quantisation tables and its mbarrier:
 *     as soon as a warp has pulled its coefficients into registers (row pass)
 *     its lane 0 starts the loads of the warp's next tile, which land while the
 *     warp runs the column pass and the colour loop.  There is no CTA-wide
 *     barrier on the load path.
 *   tile descriptors: built on the host per plan (128 bytes per tile), fetched
 *     two tiles ahead into a 4-slot smem ring with cp.async.bulk.
 *   chroma warps -> luma warps: the clamped Cb/Cr samples (2 bytes per sample)
 *     through a double-buffered exchange area and named barriers, so chroma
 *     may run up to two tiles ahead of luma.
 *   regs -> HBM: each luma thread holds 16 adjacent pixels of a row = 48 bytes =
 *     three 128-bit stores; a warp covers 1536 contiguous bytes per row.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "jgpu_colour_fixed.h"
#include "jgpu_fused_common.cuh" /* PTX wrappers and per-thread stages, shared with jgpu_mcu.cu */
#include "jgpu_internal.h"
#include "jgpu_kernels.cuh"
#include "jgpu_launch.h"

namespace jgpu {

/* Build-time tuning knobs (A/B-tested on the GPU, see profiles/). */
#ifndef JGPU_FUSED_G
#define JGPU_FUSED_G 0          /* 32-pair column groups per tile; 0: per-mode default (mode_groups) */
#endif
/* Tile shapes of the other sampling modes, measured on 4K batches (profiles/r1_ab_notes.md): what
 * counts is how evenly 3840 (1920) pixels divide into tiles, not the warps per CTA. */
#ifndef JGPU_FUSED_GRAY_WARPS
#define JGPU_FUSED_GRAY_WARPS 2 /* warps of 64 blocks per grey tile: 1024 px (3 warps = 1536 px left 4K rows 2.5 tiles wide) */
#endif
#ifndef JGPU_FUSED_G444
#define JGPU_FUSED_G444 1       /* column groups, luma 1x1: 512 px tiles */
#endif
#ifndef JGPU_FUSED_G440
#define JGPU_FUSED_G440 2       /* luma 1x2: 1024 px tiles */
#endif
#ifndef JGPU_FUSED_G422
#define JGPU_FUSED_G422 2       /* luma 2x1: 1024 px tiles (3 groups = 1536 px: 2.5 tiles per 4K row) */
#endif
#ifndef JGPU_FUSED_MINCTAS
#define JGPU_FUSED_MINCTAS 0    /* 0: size registers for 12 warps per SM */
#endif

/* Column groups per tile, per sampling mode.  Every choice keeps 12 warps per SM (the most 168
 * registers per thread allow): 4:2:0 -> 4 luma + 2 chroma warps x 2 CTAs, 4:2:2 -> 2 + 2 x 3 CTAs, ... */
constexpr int mode_groups(int hs, int vs, bool gray) {
  return JGPU_FUSED_G > 0 ? JGPU_FUSED_G
         : gray           ? 1
         : hs == 2        ? (vs == 2 ? 2 : JGPU_FUSED_G422)
                          : (vs == 2 ? JGPU_FUSED_G440 : JGPU_FUSED_G444);
}

constexpr int kWarpBytes = 2 * kBoxBytes;        /* the two boxes of a warp, 1 KB aligned */
constexpr int kWarpTabBytes = 2 * kQtabBytes;    /* its table(s): luma one, chroma Cb then Cr */
constexpr int kMaxYWarps = 6;
constexpr int kMaxCWarps = 4;
constexpr int kDescSlots = 4;

/* One tile, as the host describes it (fused_plan_build): where its boxes start
 * in the coefficient buffer (global block indices: the buffer viewed as rows of
 * 64 int16) and where its pixels go. */
struct __align__(16) TileDesc {
  long long rgb_base;   /* byte offset in the rgb buffer of the tile's top-left pixel */
  int32_t width_left;   /* visible pixels from the tile's left edge to the image's right edge */
  int32_t rows_left;    /* visible rows from the tile's top row to the image's bottom */
  int32_t pitch;        /* bytes per output row */
  int32_t flags;        /* bit 0: output rows are 16-byte aligned (given an aligned base pointer) */
  int32_t qidx[3];      /* 64-entry table indices: qtab_set*4 + tq */
  int32_t pad0;
  int32_t yfirst[kMaxYWarps];     /* first block of each luma warp's 64-block run */
  int32_t cfirst[2][kMaxCWarps];  /* first Cb / Cr block of each chroma warp's 32-block run */
  int32_t pad1[8];
};
static_assert(sizeof(TileDesc) == 128, "TileDesc is copied with cp.async.bulk and read with 128-bit loads");

/* HS, VS: luma sampling factors (chroma is 1x1); G: 32-pair column groups per tile. */
template <int HS, int VS, bool GRAY, int G>
struct Cfg {
  static constexpr int kYWarps = (GRAY ? JGPU_FUSED_GRAY_WARPS : VS) * G;
  static constexpr int kCWarps = GRAY ? 0 : (2 / HS) * G;
  static constexpr int kWarps = kYWarps + kCWarps;
  static constexpr int kThreads = 32 * kWarps;
  static constexpr int kMcuW = GRAY ? 8 : 8 * HS;
  static constexpr int kMcuH = GRAY ? 8 : 8 * VS;
  /* MCUs per tile: every luma warp covers 64 blocks of one block row */
  static constexpr int kTileMcus = GRAY ? 64 * kYWarps : 64 * G / HS;
  static constexpr int kChannels = GRAY ? 1 : 3;
  /* exchange area: the clamped chroma samples of one MCU's Cb/Cr block pair: 8 rows x 8
   * samples x (Cb-128, Cr-128) as two signed bytes = 16 bytes per row, padded per task */
  static constexpr int kExTask = 8 * 16 + 16;                /* odd multiple of 16 bytes */
  static constexpr int kExRegion1 = 32 * G * kExTask + 64;   /* HS==1: odd MCUs live here */
  static constexpr int kExSlot = GRAY ? 0 : (HS == 2 ? 32 * G * kExTask : 2 * 32 * G * kExTask + 128);
  /* luma staging: each luma thread parks its 16x8 clamped samples (s16x2 words) here between
   * the column pass and the colour loop: 8 rows x 32 bytes, padded */
  static constexpr int kYsTask = 8 * 32 + 16;
  static constexpr int kYsBytes = 32 * kYWarps * kYsTask;
  /* shared-memory map (offsets from a 1 KB aligned base) */
  static constexpr int kOffWarp = 0;
  static constexpr int kOffTab = kWarps * kWarpBytes;
  static constexpr int kOffEx = kOffTab + kWarps * kWarpTabBytes;
  static constexpr int kOffYs = kOffEx + 2 * kExSlot;
  /* chroma threads park row 0 of their tile here between the passes (luma threads use their
   * staging rows): 4 chunks of 16 bytes per thread, padded */
  static constexpr int kParkTask = 64 + 16;
  static constexpr int kOffPark = kOffYs + kYsBytes;
  static constexpr int kOffDesc = kOffPark + 32 * kCWarps * kParkTask;
  static constexpr int kOffBar = kOffDesc + kDescSlots * (int)sizeof(TileDesc);
  static constexpr int kSmemBytes = kOffBar + 8 * (kWarps + kDescSlots);
  /* resident CTAs per SM the register allocation is sized for */
  static constexpr int kMinCtas = JGPU_FUSED_MINCTAS > 0 ? JGPU_FUSED_MINCTAS : 384 / kThreads;
  static_assert(kYWarps <= kMaxYWarps && kCWarps <= kMaxCWarps, "TileDesc too small");
};

/* ---- the kernel ------------------------------------------------------------ */

/* WIDE: the instantiation for runs whose tables have entries above 255 (16-bit DQT).  Both
 * instantiations are launched; k_prep_qtabs sets *wide_flag and the one that does not apply
 * returns at once, which keeps the common kernel free of the two-step dequantisation. */
template <int HS, int VS, bool GRAY, int G, bool WIDE>
__global__ void __launch_bounds__(Cfg<HS, VS, GRAY, G>::kThreads, Cfg<HS, VS, GRAY, G>::kMinCtas)
k_fused(const __grid_constant__ CUtensorMap tm_rows,   /* (64, rows)            */
        const __grid_constant__ CUtensorMap tm_pairs,  /* (64, parity, pairs)   */
        const TileDesc *__restrict__ descs, int n_tiles, const uint32_t *__restrict__ qint,
        const uint32_t *__restrict__ wide_flag, uint8_t *__restrict__ rgb, int rgb_aligned) {
  using C = Cfg<HS, VS, GRAY, G>;
  if ((*wide_flag != 0) != WIDE) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  /* the 128-byte swizzle needs the boxes 1 KB aligned; the declaration above asks for it */
  const uint32_t smem0 = smem_u32(smem_raw);
  uint8_t *const smem_gen = smem_raw;
  if (smem0 & 1023u) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_data = smem0 + C::kOffBar + 8 * warp;           /* this warp's coefficients */
  const uint32_t bar_desc0 = smem0 + C::kOffBar + 8 * C::kWarps;     /* + 8*slot */

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kWarps + kDescSlots; i++) mbar_init(smem0 + C::kOffBar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  /* CTA-local tile n is global tile blockIdx.x + n*gridDim.x; its descriptor lives in ring
   * slot n % kDescSlots */
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto fetch_desc = [&](int n) {   /* thread 0 only */
    const uint32_t slot = (uint32_t)n % kDescSlots;
    mbar_expect_tx(bar_desc0 + 8 * slot, (uint32_t)sizeof(TileDesc));
    bulk_load(smem0 + C::kOffDesc + slot * (uint32_t)sizeof(TileDesc),
              descs + (blockIdx.x + (size_t)n * gridDim.x), (uint32_t)sizeof(TileDesc),
              bar_desc0 + 8 * slot);
  };
  auto wait_desc = [&](int n) -> uint32_t {   /* returns the slot's shared address */
    const uint32_t slot = (uint32_t)n % kDescSlots;
    mbar_wait(bar_desc0 + 8 * slot, ((uint32_t)n / kDescSlots) & 1u);
    return smem0 + C::kOffDesc + slot * (uint32_t)sizeof(TileDesc);
  };

  /* ---- role of this warp ---------------------------------------------------- */
  const bool is_c = warp >= C::kYWarps;
  const int cw = warp - C::kYWarps;   /* chroma warp index */
  const uint32_t wreg = smem0 + C::kOffWarp + warp * kWarpBytes;   /* boxes A, B */
  const uint32_t wtab = smem0 + C::kOffTab + warp * kWarpTabBytes;  /* this warp's table(s) */
  /* start this warp's loads for CTA-local tile n (lane 0 only) */
  auto fire = [&](int n) {
    const uint32_t d = wait_desc(n);
    if (is_c) {
      const int fb = (int)lds32(d + offsetof(TileDesc, cfirst) + 4 * cw);
      const int fr = (int)lds32(d + offsetof(TileDesc, cfirst) + 4 * (kMaxCWarps + cw));
      const int qb = (int)lds32(d + offsetof(TileDesc, qidx) + 4);
      const int qr = (int)lds32(d + offsetof(TileDesc, qidx) + 8);
      mbar_expect_tx(bar_data, 2 * kBoxBytes + 2 * kQtabBytes);
      tma_load_2d(wreg, &tm_rows, 0, fb, bar_data);
      tma_load_2d(wreg + kBoxBytes, &tm_rows, 0, fr, bar_data);
      bulk_load(wtab, qint + (size_t)qb * 64, kQtabBytes, bar_data);
      bulk_load(wtab + kQtabBytes, qint + (size_t)qr * 64, kQtabBytes, bar_data);
    } else {
      /* the 32 even-position and the 32 odd-position blocks of this warp's 64-block run */
      const int first = (int)lds32(d + offsetof(TileDesc, yfirst) + 4 * warp);
      const int qy = (int)lds32(d + offsetof(TileDesc, qidx));
      mbar_expect_tx(bar_data, 2 * kBoxBytes + kQtabBytes);
      tma_load_3d(wreg, &tm_pairs, 0, first & 1, first >> 1, bar_data);
      tma_load_3d(wreg + kBoxBytes, &tm_pairs, 0, (first + 1) & 1, (first + 1) >> 1, bar_data);
      bulk_load(wtab, qint + (size_t)qy * 64, kQtabBytes, bar_data);
    }
  };

  if (threadIdx.x == 0) {
    fetch_desc(0);
    if (my_tiles > 1) fetch_desc(1);
  }
  if (lane == 0) fire(0);

  /* Per-thread geometry, derived from the thread index.  It is re-derived (from a fresh,
   * un-CSE-able read of %tid.x) at each point of use instead of being kept in registers
   * across the IDCT, where every register counts. */
  struct Geo {
    int px_x, px_y;      /* pixel position of this thread's pair inside the tile */
    uint32_t ex_rel;     /* exchange-area task (chroma: the MCU it writes; luma: the one(s) it reads) */
    uint32_t ys_a;       /* luma staging rows */
    uint32_t park[4];    /* where row 0 of the tile waits between the passes */
  };
  auto geo = [&]() -> Geo {
    uint32_t tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int w = (int)(tid >> 5), l = (int)(tid & 31);
    const bool c = w >= C::kYWarps;
    const int cwi = w - C::kYWarps;
    Geo g;
    if (c) {
      g.px_x = (HS == 2 ? 16 : 8) * (32 * cwi + l);
      g.px_y = 0;
    } else if (GRAY) {
      g.px_x = 16 * (32 * w + l);
      g.px_y = 0;
    } else {
      g.px_x = 16 * (32 * (w % G) + l);
      g.px_y = 8 * (w / G);
    }
    g.ex_rel = 0;
    if (!GRAY) {
      if (HS == 2) {
        g.ex_rel = (uint32_t)(c ? 32 * cwi + l : 32 * (w % G) + l) * C::kExTask;
      } else if (c) {
        const int mcu = 32 * cwi + l;
        g.ex_rel = (mcu & 1) * C::kExRegion1 + (mcu >> 1) * C::kExTask;
      } else {
        g.ex_rel = (uint32_t)(32 * (w % G) + l) * C::kExTask;
      }
    }
    g.ys_a = smem0 + C::kOffYs + (uint32_t)(c ? 0 : 32 * w + l) * C::kYsTask;
    if (c) {
      const uint32_t base = smem0 + C::kOffPark + (uint32_t)(32 * cwi + l) * C::kParkTask;
#pragma unroll
      for (int j = 0; j < 4; j++) g.park[j] = base + 16 * j;
    } else {
      /* inside the staging rows, at places the column pass overwrites only after it has
       * fetched the chunk: pair j of every row k is written at the end of step j */
      g.park[0] = g.ys_a;             /* row 0, bytes  0..15: written in steps 0,1; read in step 0 */
      g.park[1] = g.ys_a + 32 + 16;   /* row 1, bytes 16..31: written in steps 2,3; read in step 1 */
      g.park[2] = g.ys_a + 64 + 16;   /* row 2, bytes 16..31: written in steps 2,3; read in step 2 */
      g.park[3] = g.ys_a + 256;       /* the padding */
    }
    return g;
  };
  const uint8_t *const wgen = smem_gen + C::kOffWarp + warp * kWarpBytes;

  for (int it = 0; it < my_tiles; it++) {
    /* grey has no chroma hand-off that would keep its warps within a tile of each other;
     * the descriptor ring needs that, so couple them here */
    if (GRAY) named_sync(1, C::kThreads);
    /* keep the descriptor ring two tiles ahead (slot of tile it-2: nobody reads it any more) */
    if (threadIdx.x == 0 && it + 2 < my_tiles) fetch_desc(it + 2);

    const uint32_t da = wait_desc(it);
    bool active;
    {
      const Geo g = geo();
      const uint4 h0 = lds128(da);
      active = g.px_x < (int)h0.z && g.px_y < (int)h0.w;
    }
    mbar_wait(bar_data, (uint32_t)it & 1u);

    {
      pair32 m[8][8];
      if (active) {
        const uint4 *const qa = reinterpret_cast<const uint4 *>(smem_gen + C::kOffTab + warp * kWarpTabBytes);
        const uint4 *const qb = is_c ? qa + 16 : qa;   /* chroma: Cb table, then Cr table */
        const Geo g = geo();
        pair_row_pass<WIDE>(m, wgen, wgen + kBoxBytes, lane, qa, qb, g.park);
      }
      /* this warp's boxes are in registers: start the loads of its next tile */
      __syncwarp();
      if (lane == 0 && it + 1 < my_tiles) fire(it + 1);
      /* chroma may not overwrite an exchange slot the luma warps still read (tile it-2) */
      if (!GRAY && is_c && it >= 2) named_sync(4 + (it & 1), C::kThreads);
      if (active) {
        const pair32 magic = p_make_bits(kMagicBits, kMagicBits);
        const Geo g = geo();
        const uint32_t ys_a = g.ys_a;
        const uint32_t ex_a = smem0 + C::kOffEx + (it & 1) * C::kExSlot + g.ex_rel;
        column_pass_by_pairs_parked(m, [&](int j, pair32 &a, pair32 &b) {
          const uint4 c = lds128(g.park[j]);
          a = p_make_bits(c.x, c.y);
          b = p_make_bits(c.z, c.w);
        }, [&](int j, pair32 (&u)[8], pair32 (&v)[8]) {
          if (!is_c) {
            /* luma: (short)floor + 128, clamp; pixels 2j, 2j+1 of row k of block A and of
             * block B as two s16x2 words, parked in this thread's staging rows */
#pragma unroll
            for (int k = 0; k < 8; k++) {
              uint32_t ulo, uhi, vlo, vhi;
              p_split_bits(p_add_rm(u[k], magic), ulo, uhi);
              p_split_bits(p_add_rm(v[k], magic), vlo, vhi);
              sts64(ys_a + 32 * k + 8 * j, make_uint2(clamp_pair_u8(ulo, vlo), clamp_pair_u8(uhi, vhi)));
            }
          } else if (!GRAY) {
            /* chroma: clamped samples of columns 2j, 2j+1 as four signed bytes
             * (Cb, Cr, Cb, Cr) -> exchange area */
#pragma unroll
            for (int k = 0; k < 8; k++) {
              sts32(ex_a + 16 * k + 4 * j, __byte_perm(chroma_clamped(u[k]), chroma_clamped(v[k]), 0x6420));
            }
          }
        });
      }
    }
    if (GRAY) {
      if (!active) continue;
    } else if (is_c) {
      named_arrive(2 + (it & 1), C::kThreads);   /* this tile's chroma is in the exchange area */
      continue;
    } else {
      named_sync(2 + (it & 1), C::kThreads);
    }

    /* ---- luma threads: add the colour offsets, pack, store ------------------ */
    if (active) {
      const Geo g = geo();
      const int px_x = g.px_x, px_y = g.px_y;
      const uint32_t ys_a = g.ys_a;
      const uint32_t ex_a = smem0 + C::kOffEx + (it & 1) * C::kExSlot + g.ex_rel;
      const uint4 h0 = lds128(da);
      const uint2 h1 = lds64(da + 16);
      const long long rgb_base = (long long)(((unsigned long long)h0.y << 32) | h0.x);
      const int pitch = (int)h1.x;
      const int vis_px = min(16, (int)h0.z - px_x);
      const int vis_rows = min(8, (int)h0.w - px_y);
      const bool fast = (h1.y & (uint32_t)rgb_aligned & 1u) != 0 && vis_px == 16;
      uint8_t *dst = rgb + rgb_base + (long long)px_y * pitch + (long long)px_x * C::kChannels;
      if (GRAY) {
#pragma unroll 1
        for (int k = 0; k < vis_rows; k++, dst += pitch) {
          /* staging row: (a0 b0 a1 b1 | a2 b2 a3 b3), a = block A pairs, b = block B pairs */
          const uint4 t0 = lds128(ys_a + 32 * k), t1 = lds128(ys_a + 32 * k + 16);
          uint4 v;
          v.x = __byte_perm(t0.x, t0.z, 0x6420);
          v.y = __byte_perm(t1.x, t1.z, 0x6420);
          v.z = __byte_perm(t0.y, t0.w, 0x6420);
          v.w = __byte_perm(t1.y, t1.w, 0x6420);
          if (fast) stg128_stream(dst, v);
          else store_row_slow(dst, v, v, v, vis_px);
        }
      } else {
        /* one iteration per chroma row = VS pixel rows */
#pragma unroll 1
        for (int cr = 0; cr < 8 / VS; cr++) {
          if (cr * VS >= vis_rows) break;
          uint32_t ca[12], cb[12];   /* offsets for block A / block B: 4 pixel pairs x (R,G,B) */
          const uint32_t a = ex_a + (px_y / VS + cr) * 16;
          if (HS == 2) {
            /* 8 chroma samples, each serving one horizontal pixel pair of VS rows: offsets,
             * replicated into both halves of an s16x2 word */
            const uint4 t = lds128(a);
            const uint32_t cs[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
              uint32_t *o = i < 2 ? ca : cb;
              const int p = 6 * (i & 1);
              uint32_t r[2], g[2], b[2];
              chroma_offsets_bits2(cs[i], r, g, b);
              o[p + 0] = __byte_perm(r[0], r[0], kSelRep);
              o[p + 1] = __byte_perm(g[0], g[0], kSelRep);
              o[p + 2] = __byte_perm(b[0], b[0], kSelRep);
              o[p + 3] = __byte_perm(r[1], r[1], kSelRep);
              o[p + 4] = __byte_perm(g[1], g[1], kSelRep);
              o[p + 5] = __byte_perm(b[1], b[1], kSelRep);
            }
          } else {
            /* block A = even MCU, block B = odd MCU: 8 samples each, paired horizontally */
#pragma unroll
            for (int blk = 0; blk < 2; blk++) {
              const uint4 t = lds128(blk == 0 ? a : a + C::kExRegion1);
              const uint32_t cs[4] = {t.x, t.y, t.z, t.w};
              uint32_t *o = blk == 0 ? ca : cb;
#pragma unroll
              for (int i = 0; i < 4; i++) {
                uint32_t r[2], g[2], b[2];
                chroma_offsets_bits2(cs[i], r, g, b);
                o[3 * i + 0] = __byte_perm(r[0], r[1], kSelPair);
                o[3 * i + 1] = __byte_perm(g[0], g[1], kSelPair);
                o[3 * i + 2] = __byte_perm(b[0], b[1], kSelPair);
              }
            }
          }
#pragma unroll
          for (int sub = 0; sub < VS; sub++) {
            const int k = cr * VS + sub;
            if (k < vis_rows) {
              const uint4 t0 = lds128(ys_a + 32 * k), t1 = lds128(ys_a + 32 * k + 16);
              uint32_t w[12];
              rgb4(t0.x, t0.z, ca[0], ca[1], ca[2], ca[3], ca[4], ca[5], w[0], w[1], w[2]);
              rgb4(t1.x, t1.z, ca[6], ca[7], ca[8], ca[9], ca[10], ca[11], w[3], w[4], w[5]);
              rgb4(t0.y, t0.w, cb[0], cb[1], cb[2], cb[3], cb[4], cb[5], w[6], w[7], w[8]);
              rgb4(t1.y, t1.w, cb[6], cb[7], cb[8], cb[9], cb[10], cb[11], w[9], w[10], w[11]);
              const uint4 q0 = make_uint4(w[0], w[1], w[2], w[3]);
              const uint4 q1 = make_uint4(w[4], w[5], w[6], w[7]);
              const uint4 q2 = make_uint4(w[8], w[9], w[10], w[11]);
              if (fast) {
                stg128_stream(dst, q0);
                stg128_stream(dst + 16, q1);
                stg128_stream(dst + 32, q2);
              } else {
                store_row_slow(dst, q0, q1, q2, 3 * vis_px);
              }
              dst += pitch;
            }
          }
        }
      }
    }
    /* luma is done with exchange slot it&1: chroma may refill it for tile it+2 */
    if (!GRAY && it + 2 < my_tiles) named_arrive(4 + (it & 1), C::kThreads);
  }
}

/* ---- grey, one block per thread ---------------------------------------------
 * The kernel above keeps TWO blocks per thread (the two lanes of the packed instructions are
 * two blocks): 128 registers of transform state, 168 in all, 12 warps per SM.  Here a thread
 * owns ONE block and the packed lanes are two ROWS of it (r and r+4) in the row pass, two
 * COLUMNS (c and c+4) in the column pass, with a 2x2 exchange of register halves in between
 * (16 pairs of MOVs; no shuffle, no shared memory).  64 registers of state, about 100 in all,
 * 20 warps per SM: the occupancy experiment of DESIGN 6 (T = 1.96 ms + 11.5 ms / warps) says
 * latency, not issue slots, is what the two-block form is short of.  Arithmetic is the same
 * bit-exact sequence: prescale (y*S[r])*S[c] with S[r] different in the two lanes, inv_pass8,
 * +0.5 on the vertical DC term, inv_pass8, floor, +128, clamp.
 * A warp = 32 consecutive blocks of one block row (one TMA box, 4 KB, 128-byte swizzle) and
 * writes 256 contiguous bytes per pixel row with one 8-byte store per thread. */
constexpr int kTpbWarps = 4;
constexpr int kTpbThreads = 32 * kTpbWarps;
constexpr int kTpbOffTab = kTpbWarps * kBoxBytes;
constexpr int kTpbOffDesc = kTpbOffTab + kTpbWarps * kQtabBytes;
constexpr int kTpbOffBar = kTpbOffDesc + kDescSlots * (int)sizeof(TileDesc);
constexpr int kTpbSmemBytes = kTpbOffBar + 8 * (kTpbWarps + kDescSlots);

#ifndef JGPU_TPB_MINCTAS
#define JGPU_TPB_MINCTAS 5
#endif
template <bool WIDE>
__global__ void __launch_bounds__(kTpbThreads, JGPU_TPB_MINCTAS)
k_gray_tpb(const __grid_constant__ CUtensorMap tm_rows, const TileDesc *__restrict__ descs, int n_tiles,
           const uint32_t *__restrict__ qint, const uint32_t *__restrict__ wide_flag,
           uint8_t *__restrict__ rgb, int rgb_aligned) {
  if ((*wide_flag != 0) != WIDE) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  if (smem0 & 1023u) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_data = smem0 + kTpbOffBar + 8 * warp;
  const uint32_t bar_desc0 = smem0 + kTpbOffBar + 8 * kTpbWarps;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTpbWarps + kDescSlots; i++) mbar_init(smem0 + kTpbOffBar + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto fetch_desc = [&](int n) {   /* thread 0 only */
    const uint32_t slot = (uint32_t)n % kDescSlots;
    mbar_expect_tx(bar_desc0 + 8 * slot, (uint32_t)sizeof(TileDesc));
    bulk_load(smem0 + kTpbOffDesc + slot * (uint32_t)sizeof(TileDesc), descs + (blockIdx.x + (size_t)n * gridDim.x),
              (uint32_t)sizeof(TileDesc), bar_desc0 + 8 * slot);
  };
  auto wait_desc = [&](int n) -> uint32_t {
    const uint32_t slot = (uint32_t)n % kDescSlots;
    mbar_wait(bar_desc0 + 8 * slot, ((uint32_t)n / kDescSlots) & 1u);
    return smem0 + kTpbOffDesc + slot * (uint32_t)sizeof(TileDesc);
  };
  const uint32_t wbox = smem0 + warp * kBoxBytes, wtab = smem0 + kTpbOffTab + warp * kQtabBytes;
  auto fire = [&](int n) {   /* lane 0 of each warp: this warp's 32 blocks and its table */
    const uint32_t d = wait_desc(n);
    const int first = (int)lds32(d + offsetof(TileDesc, yfirst) + 4 * warp);
    const int qy = (int)lds32(d + offsetof(TileDesc, qidx));
    mbar_expect_tx(bar_data, kBoxBytes + kQtabBytes);
    tma_load_2d(wbox, &tm_rows, 0, first, bar_data);
    bulk_load(wtab, qint + (size_t)qy * 64, kQtabBytes, bar_data);
  };
  if (threadIdx.x == 0) {
    fetch_desc(0);
    if (my_tiles > 1) fetch_desc(1);
  }
  if (lane == 0) fire(0);
  const uint8_t *const box = smem_raw + warp * kBoxBytes + 128 * lane;
  const uint4 *const q = reinterpret_cast<const uint4 *>(smem_raw + kTpbOffTab + warp * kQtabBytes);
  const int sw = lane & 7;
  const int px_x = 8 * (32 * warp + lane);

  for (int it = 0; it < my_tiles; it++) {
    /* the warps of a CTA stay within a tile of each other: the descriptor ring relies on it */
    named_sync(1, kTpbThreads);
    if (threadIdx.x == 0 && it + 2 < my_tiles) fetch_desc(it + 2);
    const uint32_t da = wait_desc(it);
    const uint4 h0 = lds128(da);
    const uint2 h1 = lds64(da + 16);
    const bool active = px_x < (int)h0.z;
    mbar_wait(bar_data, (uint32_t)it & 1u);
    pair32 m[4][8];
    if (active) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const uint4 a = *reinterpret_cast<const uint4 *>(box + 16 * (r ^ sw));
        const uint4 b = *reinterpret_cast<const uint4 *>(box + 16 * ((r + 4) ^ sw));
        const uint4 z = make_uint4(0, 0, 0, 0);
        load_two_rows_packed<WIDE>(m[r], a, b, q[r], q[r + 4], WIDE ? q[8 + r] : z, WIDE ? q[12 + r] : z, r);
        inv_pass8(m[r]);
      }
    }
    __syncwarp();
    if (lane == 0 && it + 1 < my_tiles) fire(it + 1);
    if (!active) continue;
    const pair32 magic = p_make_bits(kMagicBits, kMagicBits);
    uint32_t w[8][4];   /* clamped samples: w[k][c] = (x[k][c], x[k][c+4]) as two 16-bit halves */
#pragma unroll
    for (int c = 0; c < 4; c++) {
      pair32 v[8];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        uint32_t alo, ahi, blo, bhi;
        p_split_bits(m[r][c], alo, ahi);
        p_split_bits(m[r][c + 4], blo, bhi);
        v[r] = p_make_bits(alo, blo);       /* row r:     columns c, c+4 */
        v[r + 4] = p_make_bits(ahi, bhi);   /* row r + 4: columns c, c+4 */
      }
      v[0] = p_add_half(v[0]);
      inv_pass8(v);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        uint32_t lo, hi;
        p_split_bits(p_add_rm(v[k], magic), lo, hi);
        w[k][c] = clamp_pair_u8(lo, hi);
      }
    }
    const long long rgb_base = (long long)(((unsigned long long)h0.y << 32) | h0.x);
    const int pitch = (int)h1.x;
    const int vis_px = min(8, (int)h0.z - px_x), vis_rows = min(8, (int)h0.w);
    const bool fast = (h1.y & (uint32_t)rgb_aligned & 1u) != 0 && vis_px == 8;
    uint8_t *dst = rgb + rgb_base + px_x;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k < vis_rows) {
        const uint32_t p = __byte_perm(w[k][0], w[k][1], 0x6240);   /* x0 x1 x4 x5 */
        const uint32_t t = __byte_perm(w[k][2], w[k][3], 0x6240);   /* x2 x3 x6 x7 */
        const uint32_t o0 = __byte_perm(p, t, 0x5410), o1 = __byte_perm(p, t, 0x7632);
        if (fast) {
          asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(o0), "r"(o1) : "memory");
        } else {
          for (int i = 0; i < vis_px; i++) dst[i] = (uint8_t)((i < 4 ? o0 : o1) >> (8 * (i & 3)));
        }
      }
      dst += pitch;
    }
  }
}

/* u16 tables -> packed tables for IDP.2A (one tiny launch per run) and the 16-bit flag.  Per table 64 words:
 * word r*4+i        = (q[r][2i] & 255)  | (q[r][2i+1] & 255) << 24      (low bytes)
 * word 32 + r*4+i   = (q[r][2i] >> 8)   | (q[r][2i+1] >> 8)  << 24      (high bytes) */
/* One CTA walks all the words, so that the flag is WRITTEN (0 or 1) rather than OR-ed into a word
 * somebody must zero first: chunks of one batch run this on several streams at once with the same
 * tables (jgpu_decode_batch_host), and every such launch stores the same bytes -- a zeroing memset
 * from a later chunk could otherwise land between an earlier chunk's two kernel instantiations. */
__global__ void __launch_bounds__(1024) k_prep_qtabs(const uint16_t *__restrict__ q, uint32_t *__restrict__ out,
                                                     int n_words, uint32_t *__restrict__ wide_flag, int vec) {
  int wide = 0;
  if (vec) {
    /* eight entries per thread and step (one 128-bit load): four low-byte words, four high-byte words.  One CTA on
     * purpose -- the flag is the OR over all tables and is only ever written whole -- so it has to be quick: a batch
     * of JPEG files brings a table set per file and every group of files converts them (512 tables: 65 us with one
     * entry per thread and step, 6 us this way) */
    const int n_vec = n_words / 8;   /* 64 entries = 8 vectors per table */
    for (int i = threadIdx.x; i < n_vec; i += blockDim.x) {
      const int table = i >> 3, k = i & 7;
      const uint4 v = reinterpret_cast<const uint4 *>(q)[i];
      const uint32_t e[4] = {v.x, v.y, v.z, v.w};   /* e[j] = entries 8k+2j (low half), 8k+2j+1 (high half) */
      uint32_t lo[4], hi[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t q0 = e[j] & 0xffffu, q1 = e[j] >> 16;
        lo[j] = (q0 & 255u) | ((q1 & 255u) << 24);
        hi[j] = (q0 >> 8) | ((q1 >> 8) << 24);
        wide |= hi[j] != 0u;
      }
      uint4 *o = reinterpret_cast<uint4 *>(out + table * 64);
      o[k] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      o[8 + k] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    }
  } else {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) {
      const int table = i >> 6, w = i & 63, hi = w >> 5, idx = (w & 31) * 2;
      const uint32_t q0 = q[table * 64 + idx], q1 = q[table * 64 + idx + 1];
      const uint32_t v = hi ? ((q0 >> 8) | ((q1 >> 8) << 24)) : ((q0 & 255u) | ((q1 & 255u) << 24));
      out[i] = v;
      wide |= (hi && v);
    }
  }
  wide = __syncthreads_or(wide);
  if (threadIdx.x == 0) *wide_flag = wide ? 1u : 0u;
}

/* ---- host side ------------------------------------------------------------- */

/* u16 tables -> packed tables + the 16-bit flag, on `stream` (shared with jgpu_mcu.cu). */
cudaError_t launch_prep_qtabs(const uint16_t *qtabs, uint32_t *qint, int n_tables, uint32_t *wide_flag,
                              cudaStream_t stream) {
  const int vec = ((reinterpret_cast<uintptr_t>(qtabs) | reinterpret_cast<uintptr_t>(qint)) & 15) == 0 ? 1 : 0;
  k_prep_qtabs<<<1, 1024, 0, stream>>>(qtabs, qint, n_tables * 64, wide_flag, vec);
  return cudaGetLastError();
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


struct ModeInfo {
  int hs, vs, gray, groups;
  int tile_mcus, mcu_w, mcu_h, threads, ywarps, cwarps, channels;
  size_t smem;
  int ctas_per_sm;
};

ModeInfo g_modes[kNumFusedModes];
bool g_configured = false;
bool g_gray_tpb = false;   /* grey images go through k_gray_tpb (JGPU_GRAY_TPB=1, experiment) */

template <int HS, int VS, bool GRAY>
cudaError_t configure_mode(int mode) {
  using C = Cfg<HS, VS, GRAY, mode_groups(HS, VS, GRAY)>;
  auto *f = &k_fused<HS, VS, GRAY, mode_groups(HS, VS, GRAY), false>;
  cudaError_t ew = cudaFuncSetAttribute(&k_fused<HS, VS, GRAY, mode_groups(HS, VS, GRAY), true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes);
  if (ew != cudaSuccess) return ew;
  ModeInfo &mi = g_modes[mode];
  mi.hs = HS; mi.vs = VS; mi.gray = GRAY; mi.groups = mode_groups(HS, VS, GRAY);
  mi.tile_mcus = C::kTileMcus;
  mi.mcu_w = C::kMcuW;
  mi.mcu_h = C::kMcuH;
  mi.threads = C::kThreads;
  mi.ywarps = C::kYWarps;
  mi.cwarps = C::kCWarps;
  mi.channels = C::kChannels;
  mi.smem = C::kSmemBytes;
  /* experiment knob: pad the dynamic shared memory to lower the CTAs per SM (occupancy
   * sensitivity runs, profiles/r1_ab_notes.md); never set in production */
  if (const char *pad = getenv("JGPU_FUSED_PAD_SMEM")) {
    mi.smem = std::min<size_t>(mi.smem + (size_t)atoi(pad), 227 * 1024);
    cudaFuncSetAttribute(&k_fused<HS, VS, GRAY, mode_groups(HS, VS, GRAY), true>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.smem);
  }
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
                        const CUtensorMap &tm_pairs, const TileDesc *descs, int n_tiles,
                        const uint32_t *qint, const uint32_t *wide_flag, uint8_t *rgb, int rgb_aligned) {
  using C = Cfg<HS, VS, GRAY, mode_groups(HS, VS, GRAY)>;
  k_fused<HS, VS, GRAY, mode_groups(HS, VS, GRAY), false><<<grid, C::kThreads, smem, stream>>>(
      tm_rows, tm_pairs, descs, n_tiles, qint, wide_flag, rgb, rgb_aligned);
  k_fused<HS, VS, GRAY, mode_groups(HS, VS, GRAY), true><<<grid, C::kThreads, smem, stream>>>(
      tm_rows, tm_pairs, descs, n_tiles, qint, wide_flag, rgb, rgb_aligned);
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
  if (const char *tpb = getenv("JGPU_GRAY_TPB")) g_gray_tpb = atoi(tpb) != 0;
  if (g_gray_tpb) {
    ModeInfo &mi = g_modes[kModeGray];
    if ((e = cudaFuncSetAttribute(&k_gray_tpb<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTpbSmemBytes)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(&k_gray_tpb<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTpbSmemBytes)) != cudaSuccess) {
      return e;
    }
    mi.tile_mcus = 32 * kTpbWarps;
    mi.threads = kTpbThreads;
    mi.ywarps = kTpbWarps;
    mi.smem = kTpbSmemBytes;
    int n = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, &k_gray_tpb<false>, kTpbThreads, kTpbSmemBytes)) != cudaSuccess) return e;
    mi.ctas_per_sm = std::max(n, 1);
  }
  g_configured = true;
  return cudaSuccess;
}

struct FusedPlanImpl {
  int n = 0;
  int sm_count = 0;
  void *d_descs[kNumFusedModes] = {};
  int n_tiles[kNumFusedModes] = {};
  std::vector<int> first_tile[kNumFusedModes]; /* per mode, n+1 entries */
  void *d_qint = nullptr;
  int qint_cap = 0; /* tables */
  long long coef_rows = 0; /* 128-byte rows the batch touches */
  /* tensor maps, rebuilt when the coefficient pointer changes */
  const void *map_ptr = nullptr;
  CUtensorMap tm_rows, tm_pairs;
  /* mixed batches: every sampling mode is its own persistent kernel; they run on side streams
   * forked from / joined to the caller's stream so that one mode's tail (CTAs that have run
   * out of tiles) overlaps the next mode's start */
  cudaStream_t side[kNumFusedModes] = {};
  cudaEvent_t fork = nullptr, join[kNumFusedModes] = {};
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
  std::vector<TileDesc> tiles[kNumFusedModes];
  for (int m = 0; m < kNumFusedModes; m++) p->first_tile[m].assign(n + 1, 0);
  for (int i = 0; i < n; i++) {
    const jgpu_image_desc &d = descs[i];
    const jgpu_layout &lay = layouts[i];
    const ModeInfo &mi = g_modes[modes[i]];
    int block0[3] = {0, 0, 0}, hblocks[3] = {0, 0, 0};
    for (int c = 0; c < d.ncomps; c++) {
      block0[c] = (int)((d.coef_off + lay.plane[c].coef_off) / 64);
      hblocks[c] = lay.plane[c].hblocks;
    }
    p->coef_rows = std::max<long long>(p->coef_rows, (d.coef_off + lay.coef_len + 63) / 64);
    for (int m = 0; m < kNumFusedModes; m++) p->first_tile[m][i] = (int)tiles[m].size();
    const int nh = mi.gray ? lay.plane[0].hblocks : lay.nhmb;
    const int nv = mi.gray ? lay.plane[0].vblocks : lay.nvmb;
    for (int r = 0; r < nv; r++) {
      for (int x = 0; x < nh; x += mi.tile_mcus) {
        const int x0 = x * mi.mcu_w, y0 = r * mi.mcu_h;
        /* tiles wholly to the right of / below the visible image carry no pixels */
        if (x0 >= d.width || y0 >= d.height) continue;
        TileDesc t;
        memset(&t, 0, sizeof(t));
        t.pitch = d.width * mi.channels;
        t.rgb_base = d.rgb_off + ((long long)y0 * d.width + x0) * mi.channels;
        t.width_left = d.width - x0;
        t.rows_left = d.height - y0;
        t.flags = ((t.rgb_base & 15) == 0 && (t.pitch & 15) == 0) ? 1 : 0;
        for (int c = 0; c < d.ncomps; c++) t.qidx[c] = d.qtab_set * 4 + d.tq[c];
        for (int w = 0; w < mi.ywarps; w++) {
          if (mi.gray) {
            t.yfirst[w] = block0[0] + r * hblocks[0] + x + (mi.tile_mcus / mi.ywarps) * w;
          } else {
            t.yfirst[w] = block0[0] + (r * mi.vs + w / mi.groups) * hblocks[0] + x * mi.hs + 64 * (w % mi.groups);
          }
        }
        for (int cw = 0; cw < mi.cwarps; cw++) {
          t.cfirst[0][cw] = block0[1] + r * hblocks[1] + x + 32 * cw;
          t.cfirst[1][cw] = block0[2] + r * hblocks[2] + x + 32 * cw;
        }
        tiles[modes[i]].push_back(t);
      }
    }
  }
  for (int m = 0; m < kNumFusedModes; m++) p->first_tile[m][n] = (int)tiles[m].size();
  for (int m = 0; m < kNumFusedModes; m++) {
    p->n_tiles[m] = (int)tiles[m].size();
    if (tiles[m].empty()) continue;
    if (cudaMalloc(&p->d_descs[m], sizeof(TileDesc) * tiles[m].size()) != cudaSuccess ||
        cudaMemcpy(p->d_descs[m], tiles[m].data(), sizeof(TileDesc) * tiles[m].size(),
                   cudaMemcpyHostToDevice) != cudaSuccess) {
      return jgpu_fail("fused plan: tile descriptor upload failed");
    }
  }
  return 0;
}

void fused_plan_release(FusedPlan &fp) {
  FusedPlanImpl *p = static_cast<FusedPlanImpl *>(fp.impl);
  if (!p) return;
  for (int m = 0; m < kNumFusedModes; m++) {
    cudaFree(p->d_descs[m]);
    if (p->side[m]) cudaStreamDestroy(p->side[m]);
    if (p->join[m]) cudaEventDestroy(p->join[m]);
  }
  if (p->fork) cudaEventDestroy(p->fork);
  cudaFree(p->d_qint);
  delete p;
  fp.impl = nullptr;
}

int fused_plan_launches(const FusedPlan &fp) {
  const FusedPlanImpl *p = static_cast<const FusedPlanImpl *>(fp.impl);
  int n = 1; /* table conversion */
  for (int m = 0; m < kNumFusedModes; m++) n += 2 * (p->n_tiles[m] > 0); /* 8-bit and 16-bit table variants */
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
    if (cudaMalloc(&p->d_qint, (size_t)n_tables * 64 * 4 + 16) != cudaSuccess) {
      return jgpu_fail("fused path: table buffer allocation failed");
    }
    p->qint_cap = n_tables;
  }
  /* the wide flag lives right behind the tables */
  uint32_t *wide_flag = (uint32_t *)p->d_qint + (size_t)p->qint_cap * 64;
  cudaError_t e = launch_prep_qtabs(qtabs, (uint32_t *)p->d_qint, n_tables, wide_flag, stream);
  if (e != cudaSuccess) return jgpu_fail("k_prep_qtabs launch failed (%s)", cudaGetErrorString(e));
  const int rgb_aligned = (reinterpret_cast<uintptr_t>(rgb) & 15) == 0 ? 1 : 0;
  int n_modes = 0;
  for (int m = 0; m < kNumFusedModes; m++) n_modes += p->first_tile[m][i1] > p->first_tile[m][i0];
  const bool forked = n_modes > 1;
  if (forked) {
    if (!p->fork && cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming) != cudaSuccess) {
      return jgpu_fail("fused path: event creation failed");
    }
    if (cudaEventRecord(p->fork, stream) != cudaSuccess) return jgpu_fail("fused path: event record failed");
  }
  cudaStream_t caller = stream;
  bool first_mode = true;
  for (int m = 0; m < kNumFusedModes; m++) {
    const int t0 = p->first_tile[m][i0], t1 = p->first_tile[m][i1];
    if (t1 <= t0) continue;
    stream = caller;
    if (forked && !first_mode) {
      /* the first mode stays on the caller's stream, the others go to side streams */
      if ((!p->side[m] && cudaStreamCreateWithFlags(&p->side[m], cudaStreamNonBlocking) != cudaSuccess) ||
          (!p->join[m] && cudaEventCreateWithFlags(&p->join[m], cudaEventDisableTiming) != cudaSuccess) ||
          cudaStreamWaitEvent(p->side[m], p->fork, 0) != cudaSuccess) {
        return jgpu_fail("fused path: side stream setup failed");
      }
      stream = p->side[m];
    }
    first_mode = false;
    const ModeInfo &mi = g_modes[m];
    /* full grids: in practice the modes' kernels follow each other with overlapping tails; giving
     * each CTAs in proportion to its work, so that they would run side by side, measured 1.54 ms
     * against 0.87 ms (profiles/r1_ab_notes.md) */
    const int grid = std::min(t1 - t0, p->sm_count * mi.ctas_per_sm);
    const TileDesc *descs = static_cast<const TileDesc *>(p->d_descs[m]) + t0;
    const uint32_t *qint = static_cast<const uint32_t *>(p->d_qint);
    switch (m) {
      case kModeGray:
        if (g_gray_tpb) {
          k_gray_tpb<false><<<grid, kTpbThreads, mi.smem, stream>>>(p->tm_rows, descs, t1 - t0, qint, wide_flag, rgb, rgb_aligned);
          k_gray_tpb<true><<<grid, kTpbThreads, mi.smem, stream>>>(p->tm_rows, descs, t1 - t0, qint, wide_flag, rgb, rgb_aligned);
          e = cudaGetLastError();
        } else {
          e = launch_mode<1, 1, true>(grid, mi.smem, stream, p->tm_rows, p->tm_pairs, descs, t1 - t0, qint, wide_flag, rgb, rgb_aligned);
        }
        break;
      case kMode444: e = launch_mode<1, 1, false>(grid, mi.smem, stream, p->tm_rows, p->tm_pairs, descs, t1 - t0, qint, wide_flag, rgb, rgb_aligned); break;
      case kMode422: e = launch_mode<2, 1, false>(grid, mi.smem, stream, p->tm_rows, p->tm_pairs, descs, t1 - t0, qint, wide_flag, rgb, rgb_aligned); break;
      case kMode420: e = launch_mode<2, 2, false>(grid, mi.smem, stream, p->tm_rows, p->tm_pairs, descs, t1 - t0, qint, wide_flag, rgb, rgb_aligned); break;
      case kMode440: e = launch_mode<1, 2, false>(grid, mi.smem, stream, p->tm_rows, p->tm_pairs, descs, t1 - t0, qint, wide_flag, rgb, rgb_aligned); break;
    }
    if (e != cudaSuccess) return jgpu_fail("fused kernel launch failed (%s)", cudaGetErrorString(e));
    if (stream != caller) {
      if (cudaEventRecord(p->join[m], stream) != cudaSuccess || cudaStreamWaitEvent(caller, p->join[m], 0) != cudaSuccess) {
        return jgpu_fail("fused path: join failed");
      }
    }
  }
  return 0;
}

}  // namespace jgpu
