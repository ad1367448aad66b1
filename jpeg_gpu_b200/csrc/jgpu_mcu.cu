/* jgpu_mcu.cu — the fused coefficient -> pixels / planes kernels, one MCU column per thread (sm_100a).
 *
 * Same work as the reference's three render passes (res/horz_quant_*.fs.glsl ->
 * res/vert.fs.glsl -> res/unyuv.fs.glsl / ungrey.fs.glsl, driven by
 * src/jpeg_gpu.c:1341-1363) and the same arithmetic as jgpu_fused.cu (round 1's kernel, kept as
 * an A/B reference): dequantise, both IDCT passes, bias / clamp, nearest-neighbour chroma
 * upsample, colour, crop, RGB8 store -- or the clamped planes themselves
 * (src/xjpeg.c:565-584) when the plan asks for YUV.
 *
 * Two kernels over the same task descriptors and the same data movement:
 *   k_tk   (second half of the file) THE PRODUCT KERNEL: the work of a task is split between a
 *          transform warp and a colour warp with different register budgets (setmaxnreg), 8 such
 *          pairs per CTA, samples handed over in 4 KB strips of shared memory; see the comment
 *          above it.  gray, 4:4:4, 4:2:2, 4:2:0, 4:4:0, 4:1:1.
 *   k_mcu  (first half) its predecessor, JGPU_KERNEL=mcu: one warp does both halves.  What follows
 *          describes it; unit, task, steps and the TMA views are k_tk's too.
 *
 *   unit   = a 16-pixel-wide column of one MCU row: one MCU for the 2x luma modes (4:2:0,
 *            4:2:2), two MCUs for the 1x modes (4:4:4, 4:4:0), two blocks for grey.
 *   thread = one unit, start to finish.  It runs the block-PAIR transform of
 *            jgpu_idct_core.cuh once per "step": first the chroma pair(s) of its unit (Cb and Cr
 *            of one MCU ride in the two packed lanes), whose clamped samples it parks in a
 *            private strip of shared memory, then one luma pair per luma block row (two
 *            horizontally adjacent blocks = 16 pixels), whose clamped samples are staged while
 *            the colour offsets are added and the 48 bytes of each pixel row are stored.
 *   warp   = 32 consecutive units = 512 pixels of one MCU row: a "task".  Warps take tasks from a
 *            counter and never wait for one another: each has its own TMA landing zone (two
 *            4 KB boxes, 128-byte swizzle), its own tables, its own mbarrier and its own ring of
 *            task descriptors.
 * In jgpu_fused.cu four luma warps and two chroma warps of a CTA met at named barriers once
 * per tile; ncu showed the chroma warps asleep half of the time and the luma warps a tenth
 * of theirs at the barrier (profiles/r2_notes.md).
 *
 * Data movement
 *   HBM -> smem: cp.async.bulk.tensor boxes of 16 or 32 blocks x 128 B: a 2-D view (64, rows) for
 *     runs of consecutive blocks, a 3-D view (64, parity, pairs) that gathers every second block
 *     (lane L owns blocks 2L and 2L+1 of a 64-block run) and, for 4:1:1, (64, 4, quads).  A warp
 *     starts the loads of its next step as soon as the row pass has pulled the current boxes
 *     into registers.
 *   task descriptors: 128 bytes per task, built on the host, fetched ahead into a per-warp ring
 *     with cp.async.bulk.
 *   regs -> HBM: per pixel row three 128-bit streaming stores per thread (a warp covers 1536
 *     contiguous bytes); planes: 16 bytes of Y / 8 bytes of Cb and of Cr per thread and row.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "jgpu_fused_common.cuh"
#include "jgpu_internal.h"
#include "jgpu_launch.h"

namespace jgpu {

#ifndef JGPU_TMA_L2_PROMOTION
#define JGPU_TMA_L2_PROMOTION CU_TENSOR_MAP_L2_PROMOTION_L2_128B   /* (256B and NONE measured: profiles/r2_notes.md) */
#endif
#ifndef JGPU_MCU_WAIT_NS
#define JGPU_MCU_WAIT_NS 0       /* suspend-time hint of the mbarrier waits; 0: plain try_wait (see mbar_try_wait_hint) */
#endif
#ifndef JGPU_MCU_L2_PREFETCH
#define JGPU_MCU_L2_PREFETCH 0   /* 1: prefetch the step after next into L2 when a step's loads are issued */
#endif

#ifdef JGPU_MCU_TRACE
/* debugging aid (never in the product build): per warp {SM id, start, end (globaltimer ns), steps} */
__device__ unsigned long long g_mcu_trace[4 * 148 * 12];
#endif

constexpr int kZoneBytes = 2 * kBoxBytes;   /* a warp's TMA landing zone: box A, box B */
constexpr int kHalfRows = 16;               /* blocks per TMA box: one half-task */
constexpr int kHalfBoxBytes = kHalfRows * 128;
constexpr int kRingSlots = 4;
constexpr int kOutRgb = 1, kOutYuv = 2;
constexpr unsigned kClaimSlots = 1024;

/* One half-task = up to 16 consecutive units of one MCU row of one image; a task = two of them
 * (lanes 0-15 and lanes 16-31), consecutive in the image's row-major order, so that the second
 * may be the start of the next MCU row: a 3840-pixel row is 15 half-tasks, and with whole-warp
 * column groups every eighth warp would run half empty.  Block numbers are GLOBAL: the
 * coefficient buffer viewed as rows of 64 int16.  Built on the host (mcu_plan_build). */
struct McuHalf {
  long long base0;        /* pixels: byte offset in the rgb buffer of the half-task's top-left pixel;
                             planes: byte offset of its top-left Y sample */
  long long base1;        /* planes: byte offset of its top-left Cb sample (Cr: + cr_delta) */
  int32_t width_left;     /* visible pixels from its left edge to the image's right edge (0: no such half) */
  int32_t rows_left;      /* visible rows from its top row to the image's bottom */
  int32_t blocks_left;    /* luma blocks from its left edge to the padded plane's right edge (0: no such half) */
  int32_t yfirst[2];      /* first luma block of its run, luma block row 0 / 1 of the MCU row */
  int32_t cfirst[2];      /* first Cb / Cr block of its run */
  int32_t pad;
};
struct __align__(16) WarpTask {
  McuHalf half[2];
  long long cr_delta;     /* planes: Cr plane offset minus Cb plane offset */
  int32_t pitch0;         /* bytes per rgb row / per Y plane row */
  int32_t pitch1;         /* bytes per chroma plane row */
  int32_t flags;          /* bit 0: rgb rows are 16-byte aligned (given an aligned base pointer) */
  int32_t qidx[3];        /* 64-entry table indices: qtab_set*4 + tq */
};
static_assert(sizeof(McuHalf) == 48 && sizeof(WarpTask) == 128, "WarpTask is copied with cp.async.bulk");

/* Build knobs (A/B-tested on the GPU, profiles/r2_notes.md).
 *   JGPU_MCU_WARPS 12: 168 registers per thread; row 0 of the register tile is parked in shared
 *                      memory between the passes; the clamped luma is staged as BYTES.
 *   JGPU_MCU_WARPS 11: 184 registers; nothing parked; luma staged as s16x2 words (no pack/unpack). */
#ifndef JGPU_MCU_WARPS
#define JGPU_MCU_WARPS 12
#endif
/* A/B knob (profiles/r2_notes.md section 6): 1 = skip the dequantisation and the row pass of a coefficient row that
 * is all zero in every block the warp holds (SURVEY 7 lever ii).  Exact: the pass maps zeros to zeros, and the sign of
 * a zero never reaches the floored samples.  Off: it has to be warp-uniform to save anything, and the test costs 9
 * instructions per row. */
#ifndef JGPU_SKIP_ZERO_ROWS
#define JGPU_SKIP_ZERO_ROWS 0
#endif
#ifndef JGPU_MCU_STAGE_BYTES
#define JGPU_MCU_STAGE_BYTES (JGPU_MCU_WARPS >= 12)
#endif
#ifndef JGPU_MCU_PARK
#define JGPU_MCU_PARK (JGPU_MCU_WARPS >= 12)
#endif

/* Per-warp shared memory.  Everything a thread parks is laid out [row][lane] (16-byte chunks of
 * the 32 lanes side by side): 128-bit accesses of a warp are conflict-free without padding. */
template <int HS, int VS, bool GRAY, bool WIDE>
struct McuCfg {
  static constexpr int kChromaSteps = GRAY ? 0 : (HS == 2 ? 1 : 2);
  static constexpr int kLumaSteps = GRAY ? 1 : VS;
  static constexpr int kSteps = kChromaSteps + kLumaSteps;
  static constexpr int kChannels = GRAY ? 1 : 3;
  static constexpr bool kStageBytes = JGPU_MCU_STAGE_BYTES != 0;
  static constexpr bool kPark = JGPU_MCU_PARK != 0;
  /* packed tables of a step: Cb and Cr (chroma) or one (luma); the high-byte halves only when the
   * batch has 16-bit tables */
  static constexpr int kTabBytes = WIDE ? kQtabBytes : kQtabBytes / 2;
  /* chroma strip of one chroma step: 8 rows x [lane] x 8 samples x (Cb-128, Cr-128) signed bytes */
  static constexpr int kChromaStep = 8 * 32 * 16;
  /* luma staging: 8 rows x [lane] x 16 clamped samples, as bytes (A0-3 B0-3 A4-7 B4-7) or as
   * s16x2 words (two 16-byte halves per row: a0 b0 a1 b1 | a2 b2 a3 b3) */
  static constexpr int kStageRow = kStageBytes ? 32 * 16 : 2 * 32 * 16;
  static constexpr int kStageBytesTotal = 8 * kStageRow;
  static constexpr int kParkBytes = kPark ? 4 * 32 * 16 : 0;
  static constexpr int kOffTab = 0;
  static constexpr int kOffChroma = 2 * kTabBytes;
  static constexpr int kOffStage = kOffChroma + kChromaSteps * kChromaStep;
  static constexpr int kOffPark = kOffStage + kStageBytesTotal;
  static constexpr int kOffRing = kOffPark + kParkBytes;
  static constexpr int kOffBar = kOffRing + kRingSlots * (int)sizeof(WarpTask);   /* 5 mbarriers */
  static constexpr int kOffLoop = kOffBar + 48;   /* the warp's step counter */
  static constexpr int kOffIdx = kOffBar + 64;    /* task index of each ring slot */
  static constexpr int kWarpMisc = kOffBar + 96;
  /* warps per CTA; one CTA per SM */
  static constexpr int kFit = (227 * 1024) / (kZoneBytes + kWarpMisc);
  static constexpr int kWarps = kFit < JGPU_MCU_WARPS ? kFit : JGPU_MCU_WARPS;
  static constexpr int kThreads = 32 * kWarps;
  static constexpr int kSmemBytes = kWarps * (kZoneBytes + kWarpMisc);
  static_assert(kWarps >= 8, "shared memory budget");
  static_assert(kWarpMisc % 16 == 0, "alignment");
};

__device__ __forceinline__ void stg64_stream(uint8_t *p, uint2 v) {
  asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

/* pair_row_pass of jgpu_fused_common.cuh with 32-bit shared addresses (the generic pointers cost four
 * more registers where none are free) and the parked row laid out [chunk][lane], or not parked.
 * zone: the warp's two boxes; tab: table of block A; tab_b: table of block B. */
template <bool WIDE, bool PARK>
__device__ __forceinline__ void mcu_row_pass(pair32 (&m)[8][8], uint32_t zone, int row, uint32_t tab, uint32_t tab_b,
                                             uint32_t park) {
  const uint32_t ra = zone + 128 * row, rb = ra + kBoxBytes;
  const int sw = row & 7;
  constexpr int kHi = kQtabBytes / 2;   /* the high-byte rows follow the 8 low-byte rows */
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int off = 16 * (r ^ sw);
    const uint4 a = lds128(ra + off);
    const uint4 b = lds128(rb + off);
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (JGPU_SKIP_ZERO_ROWS && r > 0) {
      const uint32_t any = a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w;
      if (!__any_sync(__activemask(), any != 0u)) {
#pragma unroll
        for (int c = 0; c < 8; c++) m[r][c] = p_make_bits(0u, 0u);
        continue;
      }
    }
    load_row_pair_packed<WIDE>(m[r], a, b, lds128(tab + 16 * r), lds128(tab_b + 16 * r),
                               WIDE ? lds128(tab + kHi + 16 * r) : z, WIDE ? lds128(tab_b + kHi + 16 * r) : z, r);
    inv_pass8(m[r]);
    if (PARK && r == 0) {
      /* park row 0 in shared memory until the column pass asks for it, two pairs per chunk */
#pragma unroll
      for (int j = 0; j < 4; j++) {
        uint4 c;
        p_split_bits(m[0][2 * j], c.x, c.y);
        p_split_bits(m[0][2 * j + 1], c.z, c.w);
        sts128(park + 512 * j, c);
      }
    }
  }
}

template <int HS, int VS, bool GRAY, bool WIDE, int OUT>
__global__ void __launch_bounds__(McuCfg<HS, VS, GRAY, WIDE>::kThreads, 1)
k_mcu(const __grid_constant__ CUtensorMap tm_rows,   /* (64, rows), boxes of 16 rows            */
      const __grid_constant__ CUtensorMap tm_pairs,  /* (64, parity, pairs), boxes of 16 pairs  */
      const WarpTask *__restrict__ tasks, int n_tasks, const uint32_t *__restrict__ qint,
      const uint32_t *__restrict__ wide_flag, uint8_t *__restrict__ rgb, int rgb_aligned,
      uint8_t *__restrict__ yuv, int *__restrict__ claim) {
  using C = McuCfg<HS, VS, GRAY, WIDE>;
  if ((*wide_flag != 0) != WIDE) return;
#ifdef JGPU_MCU_TRACE
  unsigned long long trace_t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace_t0));
#endif
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  if (smem0 & 1023u) __trap();   /* the 128-byte swizzle needs the boxes 1 KB aligned */

  /* Where this thread's things live.  Re-derived from a fresh, un-CSE-able read of %tid.x at
   * every point of use instead of being kept in registers across the transform, where every
   * register counts (128 of them hold the block pair). */
  struct Geo {
    int lane;
    uint32_t zone;      /* the warp's landing zone: box A, box B */
    uint32_t misc;      /* the warp's area behind the zones */
    uint32_t mine;      /* misc + 16*lane: this lane's column of every [row][lane] array */
  };
  auto geo = [&]() -> Geo {
    uint32_t tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const uint32_t w = tid >> 5;
    Geo g;
    g.lane = (int)(tid & 31);
    g.zone = smem0 + w * kZoneBytes;
    g.misc = smem0 + C::kWarps * kZoneBytes + w * C::kWarpMisc;
    g.mine = g.misc + 16u * (uint32_t)g.lane;
    return g;
  };
  /* mbarriers of a warp: [0] its coefficient loads, [1 + slot] its descriptor ring */
  auto bar_data = [&](const Geo &g) { return g.misc + C::kOffBar; };
  auto bar_ring = [&](const Geo &g, uint32_t slot) { return g.misc + C::kOffBar + 8 + 8 * slot; };
  auto desc_addr = [&](const Geo &g, int n) { return g.misc + C::kOffRing + ((uint32_t)n % kRingSlots) * (uint32_t)sizeof(WarpTask); };
  /* this lane's half-task inside a descriptor */
  auto half_addr = [&](const Geo &g, int n) { return desc_addr(g, n) + (uint32_t)(g.lane >> 4) * (uint32_t)sizeof(McuHalf); };

  {
    const Geo g = geo();
    if (g.lane == 0) {
      for (int i = 0; i < 1 + kRingSlots; i++) mbar_init(bar_data(g) + 8 * i, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }

  /* This warp's tasks: the first two are gw and gw + nw; every later one is claimed from the
   * launch's counter while the task two before it runs, so that warps which fall behind (their SM
   * sees longer memory latencies; measured spread of a static split: 10 %) simply take fewer, and
   * batches of mixed sizes balance themselves.
   * Local task n lives in ring slot n % 4: its index (>= n_tasks: there is none) and its descriptor. */
  auto idx_addr = [&](const Geo &g, int n) { return g.misc + C::kOffIdx + 4u * ((uint32_t)n % kRingSlots); };
  auto fetch_desc = [&](const Geo &g, int n, int idx) {   /* lane 0 only */
    const uint32_t slot = (uint32_t)n % kRingSlots;
    sts32(idx_addr(g, n), (uint32_t)idx);
    if (idx < n_tasks) {
      mbar_expect_tx(bar_ring(g, slot), (uint32_t)sizeof(WarpTask));
      bulk_load(desc_addr(g, n), tasks + idx, (uint32_t)sizeof(WarpTask), bar_ring(g, slot));
    }
  };
  auto wait_desc = [&](const Geo &g, int n) {
    mbar_wait_hint<JGPU_MCU_WAIT_NS>(bar_ring(g, (uint32_t)n % kRingSlots), ((uint32_t)n / kRingSlots) & 1u);
  };
  /* Start the loads of step s of local task n (lane 0 only; its descriptor must have landed): per
   * half-task 16 rows of box A and 16 rows of box B (a half-task that does not exist repeats the
   * other one's blocks, see mcu_plan_build), and the step's table(s): the low-byte half always, the
   * high-byte half for 16-bit tables. */
  auto fire = [&](const Geo &g, int n, int s) {
    const uint32_t d = desc_addr(g, n), bar = bar_data(g), tab = g.misc + C::kOffTab;
    const bool chroma = s < C::kChromaSteps;
    mbar_expect_tx(bar, 2 * kBoxBytes + (chroma ? 2 : 1) * C::kTabBytes);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const uint32_t hd = d + h * (uint32_t)sizeof(McuHalf);
      const uint32_t dst = g.zone + h * kHalfBoxBytes;
      if (chroma) {
        int f0 = (int)lds32(hd + offsetof(McuHalf, cfirst));
        int f1 = (int)lds32(hd + offsetof(McuHalf, cfirst) + 4);
        if (HS == 2) {
          /* Cb and Cr of 16 consecutive MCUs */
          tma_load_2d(dst, &tm_rows, 0, f0, bar);
          tma_load_2d(dst + kBoxBytes, &tm_rows, 0, f1, bar);
        } else {
          /* step s takes the MCUs of parity s: every second Cb block and every second Cr block */
          f0 += s;
          f1 += s;
          tma_load_3d(dst, &tm_pairs, 0, f0 & 1, f0 >> 1, bar);
          tma_load_3d(dst + kBoxBytes, &tm_pairs, 0, f1 & 1, f1 >> 1, bar);
        }
      } else {
        /* the 16 even-position and the 16 odd-position blocks of a 32-block luma run */
        const int first = (int)lds32(hd + offsetof(McuHalf, yfirst) + 4 * (s - C::kChromaSteps));
        tma_load_3d(dst, &tm_pairs, 0, first & 1, first >> 1, bar);
        tma_load_3d(dst + kBoxBytes, &tm_pairs, 0, (first + 1) & 1, (first + 1) >> 1, bar);
      }
    }
    if (chroma) {
      bulk_load(tab, qint + (size_t)lds32(d + offsetof(WarpTask, qidx) + 4) * 64, C::kTabBytes, bar);
      bulk_load(tab + C::kTabBytes, qint + (size_t)lds32(d + offsetof(WarpTask, qidx) + 8) * 64, C::kTabBytes, bar);
    } else {
      bulk_load(tab, qint + (size_t)lds32(d + offsetof(WarpTask, qidx)) * 64, C::kTabBytes, bar);
    }
  };

  /* The warp's step counter lives in shared memory: no register is spent on it across the
   * transform.  Step = local task * steps per task + step inside the task. */
  {
    const Geo g = geo();
    const int gw = (int)blockIdx.x * C::kWarps + (int)(threadIdx.x >> 5), nw = (int)gridDim.x * C::kWarps;
    if (gw >= n_tasks) return;
    sts32(g.misc + C::kOffLoop, 0u);   /* every lane, same value */
    if (g.lane == 0) {
      fetch_desc(g, 0, gw);
      fetch_desc(g, 1, gw + nw);
      wait_desc(g, 0);
      fire(g, 0, 0);
    }
    __syncwarp();
  }

#pragma unroll 1
  for (;;) {
    bool is_c, active;   /* a chroma step?  does this lane's unit exist / show in this step? */
    {
      const Geo g = geo();
      const int step = (int)lds32(g.misc + C::kOffLoop);
      const int n = step / C::kSteps, s = step - n * C::kSteps;
      const bool done = (int)lds32(idx_addr(g, n)) >= n_tasks;
#ifdef JGPU_MCU_TRACE
      if (done && g.lane == 0) {
        unsigned long long t1;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long *o = g_mcu_trace + 4 * ((size_t)blockIdx.x * C::kWarps + (threadIdx.x >> 5));
        o[0] = smid; o[1] = trace_t0; o[2] = t1; o[3] = (unsigned long long)step;
      }
#endif
      if (done) break;
      __syncwarp();
      sts32(g.misc + C::kOffLoop, (uint32_t)step + 1u);
      is_c = s < C::kChromaSteps;
      wait_desc(g, n);
      const uint32_t ha = half_addr(g, n);
      const int u = g.lane & 15;
      if (OUT == kOutYuv) {
        active = 2 * u + ((HS == 1 && is_c) ? s : 0) < (int)lds32(ha + offsetof(McuHalf, blocks_left));
      } else {
        const uint2 wr = lds64(ha + offsetof(McuHalf, width_left));   /* width_left, rows_left */
        active = 16 * u + ((HS == 1 && is_c) ? 8 * s : 0) < (int)wr.x && (is_c || 8 * (s - C::kChromaSteps) < (int)wr.y);
      }
      mbar_wait_hint<JGPU_MCU_WAIT_NS>(bar_data(g), (uint32_t)step & 1u);
    }

    {
      pair32 m[8][8];
      if (active) {
        const Geo g = geo();
        const uint32_t qa = g.misc + C::kOffTab;
        const uint32_t qb = is_c ? qa + C::kTabBytes : qa;   /* chroma: Cb table, then Cr table */
        mcu_row_pass<WIDE, C::kPark>(m, g.zone, g.lane, qa, qb, g.mine + C::kOffPark);
      }
      /* the boxes are in registers: start the loads of this warp's next step */
      __syncwarp();
      {
        const Geo g = geo();
        if (g.lane == 0) {
          const int next = (int)lds32(g.misc + C::kOffLoop);   /* this step + 1 */
          const int n = next / C::kSteps, s = next - n * C::kSteps;
          if (s != 0) {
            fire(g, n, s);
          } else if ((int)lds32(idx_addr(g, n)) < n_tasks) {
            wait_desc(g, n);
            fire(g, n, 0);
          }
        }
      }
      if (!active) continue;

      const pair32 magic = p_make_bits(kMagicBits, kMagicBits);
      const Geo g = geo();
      const uint32_t mine = g.mine;
      const int s = ((int)lds32(g.misc + C::kOffLoop) - 1) % C::kSteps;
      auto row0 = [&](int j, pair32 &a, pair32 &b) {
        const uint4 c = lds128(mine + C::kOffPark + 512 * j);
        a = p_make_bits(c.x, c.y);
        b = p_make_bits(c.z, c.w);
      };
      uint32_t keep_a[8];   /* the even column step's words, until the odd one completes them */
      auto sink = [&](int j, pair32 (&u)[8], pair32 (&v)[8]) {
        if (is_c) {
          /* chroma: clamped samples of columns 2j, 2j+1 as four signed bytes (Cb, Cr, Cb, Cr);
           * two column steps make 8 bytes of the strip's row */
          const uint32_t strip = mine + C::kOffChroma + (uint32_t)s * C::kChromaStep;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const uint32_t w = __byte_perm(chroma_clamped(u[k]), chroma_clamped(v[k]), 0x6420);
            if ((j & 1) == 0) keep_a[k] = w;
            else sts64(strip + 512 * k + 8 * (j >> 1), make_uint2(keep_a[k], w));
          }
        } else {
          /* luma: (short)floor + 128, clamp: pixels 2j, 2j+1 of row k of block A and of block B */
#pragma unroll
          for (int k = 0; k < 8; k++) {
            uint32_t ulo, uhi, vlo, vhi;
            p_split_bits(p_add_rm(u[k], magic), ulo, uhi);
            p_split_bits(p_add_rm(v[k], magic), vlo, vhi);
            const uint32_t wa = clamp_pair_u8(ulo, vlo), wb = clamp_pair_u8(uhi, vhi);
            if (C::kStageBytes) {
              /* bytes (A2j A2j+1 B2j B2j+1); two column steps make 8 bytes of the staged row */
              const uint32_t w = __byte_perm(wa, wb, 0x6420);
              if ((j & 1) == 0) keep_a[k] = w;
              else sts64(mine + C::kOffStage + C::kStageRow * k + 8 * (j >> 1), make_uint2(keep_a[k], w));
            } else {
              sts64(mine + C::kOffStage + C::kStageRow * k + 512 * (j >> 1) + 8 * (j & 1), make_uint2(wa, wb));
            }
          }
        }
      };
      if (C::kPark) column_pass_by_pairs_parked(m, row0, sink);
      else column_pass_by_pairs(m, sink);
    }

    const Geo g = geo();
    const int u = g.lane & 15;   /* this lane's unit inside its half-task */
    int n, s;
    {
      const int step = (int)lds32(g.misc + C::kOffLoop) - 1;
      n = step / C::kSteps;
      s = step - n * C::kSteps;
    }
    const int yr = s - C::kChromaSteps;   /* luma block row inside the MCU row */
    const uint32_t da = desc_addr(g, n), ha = half_addr(g, n);
    if (is_c) {
      if (OUT == kOutYuv) {
        /* planes: 8 Cb bytes and 8 Cr bytes per row, out of the strip */
        const uint2 b1 = lds64(ha + offsetof(McuHalf, base1));
        const uint2 dl = lds64(da + offsetof(WarpTask, cr_delta));
        const int cpitch = (int)lds32(da + offsetof(WarpTask, pitch1));
        const long long col = HS == 2 ? 8 * u : 16 * u + 8 * s;
        uint8_t *pb = yuv + (long long)(((unsigned long long)b1.y << 32) | b1.x) + col;
        uint8_t *pr = pb + (long long)(((unsigned long long)dl.y << 32) | dl.x);
        const uint32_t strip = g.mine + C::kOffChroma + (uint32_t)s * C::kChromaStep;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const uint4 t = lds128(strip + 512 * k);
          uint2 vb, vr;
          vb.x = __byte_perm(t.x, t.y, 0x6420) ^ 0x80808080u;
          vb.y = __byte_perm(t.z, t.w, 0x6420) ^ 0x80808080u;
          vr.x = __byte_perm(t.x, t.y, 0x7531) ^ 0x80808080u;
          vr.y = __byte_perm(t.z, t.w, 0x7531) ^ 0x80808080u;
          stg64_stream(pb, vb);
          stg64_stream(pr, vr);
          pb += cpitch;
          pr += cpitch;
        }
      }
      continue;
    }

    /* one staged luma row as 16 bytes in pixel order */
    auto staged_row_bytes = [&](int k) -> uint4 {
      const uint32_t row = g.mine + C::kOffStage + C::kStageRow * k;
      if (C::kStageBytes) {
        const uint4 t = lds128(row);
        return make_uint4(__byte_perm(t.x, t.y, 0x5410), __byte_perm(t.z, t.w, 0x5410),
                          __byte_perm(t.x, t.y, 0x7632), __byte_perm(t.z, t.w, 0x7632));
      }
      const uint4 t0 = lds128(row), t1 = lds128(row + 512);
      return make_uint4(__byte_perm(t0.x, t0.z, 0x6420), __byte_perm(t1.x, t1.z, 0x6420),
                        __byte_perm(t0.y, t0.w, 0x6420), __byte_perm(t1.y, t1.w, 0x6420));
    };

    /* First luma step of a task (lane 0 always takes part in it: its unit is the task's first):
     * claim the task after next now, pick the answer up when this step's rows are stored -- the
     * round trip to the counter hides behind them -- and request its descriptor, which then has
     * more than a task's time to arrive. */
    const bool claims = yr == 0 && g.lane == 0;
    int claimed = 0;
    if (claims) claimed = atomicAdd(claim, 1);
    auto finish_claim = [&]() {
      if (claims) fetch_desc(g, n + 2, 2 * (int)gridDim.x * C::kWarps + claimed);
    };

    const uint2 b0 = lds64(ha + offsetof(McuHalf, base0));
    const long long base0 = (long long)(((unsigned long long)b0.y << 32) | b0.x);
    const int pitch = (int)lds32(da + offsetof(WarpTask, pitch0));
    if (OUT == kOutYuv) {
      /* planes: 16 Y bytes per row (8 when the unit's second block lies beyond the padded plane) */
      const bool whole = 2 * u + 1 < (int)lds32(ha + offsetof(McuHalf, blocks_left));
      const bool wide16 = whole && (pitch & 8) == 0;   /* an odd number of blocks per row: rows are only 8-byte aligned */
      uint8_t *py = yuv + base0 + (long long)(8 * yr) * pitch + 16 * u;
#pragma unroll 2
      for (int k = 0; k < 8; k++) {
        const uint4 v = staged_row_bytes(k);
        if (wide16) {
          stg128_stream(py, v);
        } else {
          stg64_stream(py, make_uint2(v.x, v.y));
          if (whole) stg64_stream(py + 8, make_uint2(v.z, v.w));
        }
        py += pitch;
      }
      finish_claim();
      continue;
    }

    /* ---- colour offsets, pack, store --------------------------------------------------------- */
    const uint2 wr = lds64(ha + offsetof(McuHalf, width_left));   /* width_left, rows_left */
    const int vis_px = min(16, (int)wr.x - 16 * u);
    const int vis_rows = min(8, (int)wr.y - 8 * yr);
    const bool fast = (lds32(da + offsetof(WarpTask, flags)) & (uint32_t)rgb_aligned & 1u) != 0 && vis_px == 16;
    uint8_t *dst = rgb + base0 + (long long)(8 * yr) * pitch + (long long)(16 * u) * C::kChannels;
    if (GRAY) {
#pragma unroll 1
      for (int k = 0; k < vis_rows; k++, dst += pitch) {
        const uint4 v = staged_row_bytes(k);
        if (fast) stg128_stream(dst, v);
        else store_row_slow(dst, v, v, v, vis_px);
      }
    } else {
      /* one iteration per chroma row = VS pixel rows */
      const uint32_t crow0 = g.mine + C::kOffChroma + (VS == 2 ? 4u * 512u * (uint32_t)yr : 0u);
#pragma unroll 1
      for (int cr = 0; cr < 8 / VS; cr++) {
        if (cr * VS >= vis_rows) break;
        uint32_t ca[12], cb[12];   /* offsets for block A / block B: 4 pixel pairs x (R,G,B) */
        if (HS == 2) {
          /* 8 chroma samples, each serving one horizontal pixel pair of VS rows: offsets,
           * replicated into both halves of an s16x2 word */
          const uint4 t = lds128(crow0 + 512 * cr);
          const uint32_t cs[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int i = 0; i < 4; i++) {
            uint32_t *o = i < 2 ? ca : cb;
            const int p = 6 * (i & 1);
            uint32_t r[2], gg[2], b[2];
            chroma_offsets_bits2(cs[i], r, gg, b);
            o[p + 0] = __byte_perm(r[0], r[0], kSelRep);
            o[p + 1] = __byte_perm(gg[0], gg[0], kSelRep);
            o[p + 2] = __byte_perm(b[0], b[0], kSelRep);
            o[p + 3] = __byte_perm(r[1], r[1], kSelRep);
            o[p + 4] = __byte_perm(gg[1], gg[1], kSelRep);
            o[p + 5] = __byte_perm(b[1], b[1], kSelRep);
          }
        } else {
          /* block A = even MCU (strip of chroma step 0), block B = odd MCU (step 1): 8 samples
           * each, one per pixel */
#pragma unroll
          for (int blk = 0; blk < 2; blk++) {
            const uint4 t = lds128(crow0 + 512 * cr + (blk ? C::kChromaStep : 0));
            const uint32_t cs[4] = {t.x, t.y, t.z, t.w};
            uint32_t *o = blk == 0 ? ca : cb;
#pragma unroll
            for (int i = 0; i < 4; i++) {
              uint32_t r[2], gg[2], b[2];
              chroma_offsets_bits2(cs[i], r, gg, b);
              o[3 * i + 0] = __byte_perm(r[0], r[1], kSelPair);
              o[3 * i + 1] = __byte_perm(gg[0], gg[1], kSelPair);
              o[3 * i + 2] = __byte_perm(b[0], b[1], kSelPair);
            }
          }
        }
#pragma unroll
        for (int sub = 0; sub < VS; sub++) {
          const int k = cr * VS + sub;
          if (k < vis_rows) {
            uint32_t ya[4], yb[4], w[12];
            if (C::kStageBytes) {   /* A01 B01 | A23 B23 | A45 B45 | A67 B67 */
              const uint4 t = lds128(g.mine + C::kOffStage + C::kStageRow * k);
              ya[0] = __byte_perm(t.x, 0u, 0x4140); yb[0] = __byte_perm(t.x, 0u, 0x4342);
              ya[1] = __byte_perm(t.y, 0u, 0x4140); yb[1] = __byte_perm(t.y, 0u, 0x4342);
              ya[2] = __byte_perm(t.z, 0u, 0x4140); yb[2] = __byte_perm(t.z, 0u, 0x4342);
              ya[3] = __byte_perm(t.w, 0u, 0x4140); yb[3] = __byte_perm(t.w, 0u, 0x4342);
            } else {                /* a0 b0 a1 b1 | a2 b2 a3 b3 */
              const uint4 t0 = lds128(g.mine + C::kOffStage + C::kStageRow * k), t1 = lds128(g.mine + C::kOffStage + C::kStageRow * k + 512);
              ya[0] = t0.x; ya[1] = t0.z; ya[2] = t1.x; ya[3] = t1.z;
              yb[0] = t0.y; yb[1] = t0.w; yb[2] = t1.y; yb[3] = t1.w;
            }
            rgb4(ya[0], ya[1], ca[0], ca[1], ca[2], ca[3], ca[4], ca[5], w[0], w[1], w[2]);
            rgb4(ya[2], ya[3], ca[6], ca[7], ca[8], ca[9], ca[10], ca[11], w[3], w[4], w[5]);
            rgb4(yb[0], yb[1], cb[0], cb[1], cb[2], cb[3], cb[4], cb[5], w[6], w[7], w[8]);
            rgb4(yb[2], yb[3], cb[6], cb[7], cb[8], cb[9], cb[10], cb[11], w[9], w[10], w[11]);
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
    finish_claim();
  }
}

/* ==== k_tk: the same work with the warps split by what they do ==================================
 *
 * ncu on k_mcu (profiles/r2_notes.md): issue slots 59 % busy, FMA pipe 50 %, ALU pipe 44 %, XU 27 % --
 * nothing saturated, yet "no eligible warp" in 41 % of the cycles.  A warp of k_mcu alternates between
 * the transform (dequantise, prescale, two 8-point passes: packed binary32 on the FMA pipe, which takes
 * two cycles per instruction) and the colour stage (PRMT / VIADDMNMX on the ALU pipe, two cycles each
 * as well); with three warps per scheduler the phases of the three coincide more often than not and
 * queue for one pipe while the other idles.
 *
 * k_tk fixes who does what.  A CTA is 8 PAIRS of warps; a pair works through tasks as a k_mcu warp does.
 *   T warp (warps 0-7, 192 registers after setmaxnreg.inc): loads, dequantise, both passes, clamp; the
 *     clamped samples of every step go into a 4 KB STRIP of shared memory (chroma: (Cb-128, Cr-128)
 *     bytes; luma: bytes in staging order), one of a ring of four per pair.
 *   K warp (warps 8-15, 64 registers after setmaxnreg.dec): waits for strips, derives the colour
 *     offsets, adds, packs and stores the pixel rows (or copies the strips to the planes), hands the
 *     strips back; it also claims the pair's tasks from the launch's counter and fetches their
 *     descriptors, four tasks ahead.
 * Every scheduler then holds two T warps and two K warps: the FMA pipe is fed by warps that want
 * nothing else and the ALU pipe by the others.  Hand-over is one mbarrier pair (full / empty) per strip,
 * touched by the two warps of the pair only; there is still no CTA-wide synchronisation after start-up.
 */
constexpr int kTkPairs = 8;
constexpr int kTkStrips = 4;
constexpr int kTkRing = 8;          /* task descriptors per pair */
constexpr int kTkAhead = 4;         /* K fetches the descriptor of task n + 4 while it works on task n */
constexpr int kStripBytes = 8 * 32 * 16;
/* Registers per thread after setmaxnreg (8 pairs x 32 x (T + K) <= 65536).  Measured (profiles/r2_notes.md):
 * 192/64 leaves the K warps spilling loop state, 184/72 is 3 % faster for 4:2:0 (2.48 -> 2.41 ms); the 1x luma
 * modes, whose colour stage reads two chroma strips, do best at 176/80.  JGPU_TK_REGS_T / _K override both. */
#ifndef JGPU_TK_REGS_T
#define JGPU_TK_REGS_T 0
#endif
#ifndef JGPU_TK_ARRIVE_ALL
#define JGPU_TK_ARRIVE_ALL 1  /* 1: every lane arrives at a strip's full / empty barrier itself (count 32): each writer is ordered
                                 before each reader directly; 0: __syncwarp(), then lane 0 arrives for the warp (count 1) */
#endif
#ifndef JGPU_TK_EXPERIMENT
#define JGPU_TK_EXPERIMENT 0  /* timing experiments only (wrong output): 1: K warps skip their work, 2: T warps skip theirs */
#endif
#ifndef JGPU_TK_PARK
#define JGPU_TK_PARK 1        /* 1: row 0 of the register tile waits in shared memory between the passes (as k_mcu at 12 warps) */
#endif
#ifndef JGPU_TK_STS32
#define JGPU_TK_STS32 0       /* 1: T stores every finished word at once (32-bit stores) instead of pairing two column steps */
#endif
#ifndef JGPU_TK_T_HIGH
#define JGPU_TK_T_HIGH 1      /* 1: the T warps are warps 8-15 (the schedulers favour high warp ids), 0: warps 0-7 */
#endif
#ifndef JGPU_TK_SPIN_NS
#define JGPU_TK_SPIN_NS 0     /* K warps: nanosleep between two looks at a strip that is not full yet (0: none) */
#endif
#ifndef JGPU_TK_REGS_K
#define JGPU_TK_REGS_K 0
#endif

template <int HS, int VS, bool GRAY, bool WIDE>
struct TkCfg {
  /* HS == 4 (4:1:1): the unit is one MCU of 32 pixels; its chroma pair comes first, then the two halves (blocks
   * 0-1, blocks 2-3) of its luma block row.  (4:1:0 would be five strips per unit with the chroma strip held to
   * the end: more than the ring of four holds, so it stays on the generic path.) */
  static_assert(!(HS == 4 && VS != 1), "4x luma modes: one luma block row per MCU only");
  static constexpr int kChromaSteps = GRAY ? 0 : (HS == 1 ? 2 : 1);
  static constexpr int kLumaSteps = GRAY ? 1 : (HS == 4 ? 2 * VS : VS);
  static constexpr int kSteps = kChromaSteps + kLumaSteps;
  static constexpr int kChannels = GRAY ? 1 : 3;
  static constexpr int kUnitPx = HS == 4 ? 32 : 16;
  static constexpr int kRegsT = JGPU_TK_REGS_T ? JGPU_TK_REGS_T : ((!GRAY && HS == 1) ? 176 : 184);
  static constexpr int kRegsK = JGPU_TK_REGS_K ? JGPU_TK_REGS_K : ((!GRAY && HS == 1) ? 80 : 72);
  static_assert(kTkPairs * 32 * (kRegsT + kRegsK) <= 65536, "register file");
  static constexpr int kTabBytes = WIDE ? kQtabBytes : kQtabBytes / 2;
  static constexpr int kOffTab = 0;                /* three resident tables: luma, Cb, Cr */
  static constexpr int kOffStrip = 3 * kTabBytes + (WIDE ? 0 : 128);
  static constexpr bool kPark = JGPU_TK_PARK != 0;
  static constexpr int kOffPark = kOffStrip + kTkStrips * kStripBytes;
  static constexpr int kOffRing = kOffPark + (kPark ? 4 * 32 * 16 : 0);
  static constexpr int kOffBar = kOffRing + kTkRing * (int)sizeof(WarpTask);   /* data, ring[8], full[4], empty[4] */
  static constexpr int kOffLoop = (kOffBar + 8 * (1 + kTkRing + 2 * kTkStrips) + 15) / 16 * 16;  /* T's place: step, local task, step in it */
  static constexpr int kOffQ = kOffLoop + 16;                                   /* which table each of the three slots holds */
  static constexpr int kOffIdx = kOffQ + 16;                                    /* task index of each ring slot */
  static constexpr int kPairMisc = (kOffIdx + 4 * kTkRing + 15) / 16 * 16;
  static constexpr int kThreads = 2 * 32 * kTkPairs;
  static constexpr int kSmemBytes = kTkPairs * (kZoneBytes + kPairMisc);
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
  static_assert(kOffStrip % 16 == 0 && kOffBar % 8 == 0, "alignment");
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

/* ---- the colour stage of one pixel row (K warps) ---------------------------------------------- */

/* one staged luma row (bytes A01 B01 | A23 B23 | A45 B45 | A67 B67) -> s16x2 words: ya[0..3] = the pixel
 * pairs of block A, yb[0..3] = of block B */
__device__ __forceinline__ void tk_staged_row(uint32_t row, uint32_t (&ya)[4], uint32_t (&yb)[4]) {
  const uint4 t = lds128(row);
  ya[0] = __byte_perm(t.x, 0u, 0x4140); yb[0] = __byte_perm(t.x, 0u, 0x4342);
  ya[1] = __byte_perm(t.y, 0u, 0x4140); yb[1] = __byte_perm(t.y, 0u, 0x4342);
  ya[2] = __byte_perm(t.z, 0u, 0x4140); yb[2] = __byte_perm(t.z, 0u, 0x4342);
  ya[3] = __byte_perm(t.w, 0u, 0x4140); yb[3] = __byte_perm(t.w, 0u, 0x4342);
}
/* ... and as 16 bytes in pixel order */
__device__ __forceinline__ uint4 tk_staged_row_bytes(uint32_t row) {
  const uint4 t = lds128(row);
  return make_uint4(__byte_perm(t.x, t.y, 0x5410), __byte_perm(t.z, t.w, 0x5410),
                    __byte_perm(t.x, t.y, 0x7632), __byte_perm(t.z, t.w, 0x7632));
}
/* colour offsets of one chroma row for block A / block B: 4 pixel pairs x (R,G,B) as s16x2 words.
 * crow: this lane's 16 bytes of the strip row; crow_b: the same in the odd MCUs' strip (1x luma modes) */
template <int HS>
__device__ __forceinline__ void tk_row_offsets(uint32_t crow, uint32_t crow_b, uint32_t (&ca)[12], uint32_t (&cb)[12]) {
  if (HS == 4) {
    /* crow: the 8 bytes of this half of the MCU: 4 chroma samples, each serving two horizontal pixel pairs */
    const uint2 t = lds64(crow);
    const uint32_t cs[2] = {t.x, t.y};
#pragma unroll
    for (int i = 0; i < 2; i++) {
      uint32_t *o = i == 0 ? ca : cb;
      uint32_t r[2], gg[2], b[2];
      chroma_offsets_bits2(cs[i], r, gg, b);
      o[0] = o[3] = __byte_perm(r[0], r[0], kSelRep);
      o[1] = o[4] = __byte_perm(gg[0], gg[0], kSelRep);
      o[2] = o[5] = __byte_perm(b[0], b[0], kSelRep);
      o[6] = o[9] = __byte_perm(r[1], r[1], kSelRep);
      o[7] = o[10] = __byte_perm(gg[1], gg[1], kSelRep);
      o[8] = o[11] = __byte_perm(b[1], b[1], kSelRep);
    }
  } else if (HS == 2) {
    /* 8 chroma samples, each serving one horizontal pixel pair: offsets replicated into both halves */
    const uint4 t = lds128(crow);
    const uint32_t cs[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint32_t *o = i < 2 ? ca : cb;
      const int p = 6 * (i & 1);
      uint32_t r[2], gg[2], b[2];
      chroma_offsets_bits2(cs[i], r, gg, b);
      o[p + 0] = __byte_perm(r[0], r[0], kSelRep);
      o[p + 1] = __byte_perm(gg[0], gg[0], kSelRep);
      o[p + 2] = __byte_perm(b[0], b[0], kSelRep);
      o[p + 3] = __byte_perm(r[1], r[1], kSelRep);
      o[p + 4] = __byte_perm(gg[1], gg[1], kSelRep);
      o[p + 5] = __byte_perm(b[1], b[1], kSelRep);
    }
  } else {
    /* block A = even MCU, block B = odd MCU: 8 samples each, one per pixel */
#pragma unroll
    for (int blk = 0; blk < 2; blk++) {
      const uint4 t = lds128(blk ? crow_b : crow);
      const uint32_t cs[4] = {t.x, t.y, t.z, t.w};
      uint32_t *o = blk == 0 ? ca : cb;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        uint32_t r[2], gg[2], b[2];
        chroma_offsets_bits2(cs[i], r, gg, b);
        o[3 * i + 0] = __byte_perm(r[0], r[1], kSelPair);
        o[3 * i + 1] = __byte_perm(gg[0], gg[1], kSelPair);
        o[3 * i + 2] = __byte_perm(b[0], b[1], kSelPair);
      }
    }
  }
}
/* 16 staged samples + the offsets -> 48 bytes of RGB as 12 words */
__device__ __forceinline__ void tk_row_words(const uint32_t (&ya)[4], const uint32_t (&yb)[4], const uint32_t (&ca)[12],
                                             const uint32_t (&cb)[12], uint32_t (&w)[12]) {
  rgb4(ya[0], ya[1], ca[0], ca[1], ca[2], ca[3], ca[4], ca[5], w[0], w[1], w[2]);
  rgb4(ya[2], ya[3], ca[6], ca[7], ca[8], ca[9], ca[10], ca[11], w[3], w[4], w[5]);
  rgb4(yb[0], yb[1], cb[0], cb[1], cb[2], cb[3], cb[4], cb[5], w[6], w[7], w[8]);
  rgb4(yb[2], yb[3], cb[6], cb[7], cb[8], cb[9], cb[10], cb[11], w[9], w[10], w[11]);
}
/* The first nbytes (<= 48) of w[0..11] to dst, whatever its alignment: byte stores up to the first
 * 4-byte boundary, funnel-shifted 32-bit stores from there, byte stores for what is left.  (The form the
 * K warps' main loop carries inline for its cropped tiles: anything cleverer costs that loop registers.) */
__device__ __forceinline__ void tk_store_any_short(uint8_t *dst, const uint32_t (&w)[12], int nbytes) {
  const int head = (int)((4u - (uint32_t)(uintptr_t)dst) & 3u);   /* bytes before the boundary */
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (i < head && i < nbytes) dst[i] = (uint8_t)(w[0] >> (8 * i));
  }
  uint8_t *p = dst + head;
  const int left = nbytes - head;
  const uint32_t sh = 8u * (uint32_t)head;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    /* the four bytes that follow byte head + 4 i of the row */
    const uint32_t v = __funnelshift_r(w[i], i + 1 < 12 ? w[i + 1] : 0u, sh);
    if (4 * i + 4 <= left) {
      *reinterpret_cast<uint32_t *>(p + 4 * i) = v;
    } else {
#pragma unroll
      for (int b = 0; b < 3; b++) {
        if (4 * i + b < left) p[4 * i + b] = (uint8_t)(v >> (8 * b));
      }
    }
  }
}

/* The first nbytes (< 48: a unit cropped by the image's right edge) of w[0..11] to dst, whatever its
 * alignment: byte stores up to the first 4-byte boundary, funnel-shifted 32-bit stores from there, byte
 * stores for what is left of the last word. */
template <int NW>
__device__ __forceinline__ void tk_store_any(uint8_t *dst, const uint32_t (&w)[NW], int nbytes) {
  const int head = (int)((4u - (uint32_t)(uintptr_t)dst) & 3u);   /* bytes before the boundary */
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (i < head && i < nbytes) dst[i] = (uint8_t)(w[0] >> (8 * i));
  }
  uint8_t *p = dst + head;
  const int left = nbytes - head;
  const uint32_t sh = 8u * (uint32_t)head;
  uint32_t last = 0u;   /* the word the row ends in */
#pragma unroll
  for (int i = 0; i < NW; i++) {
    /* the four bytes that follow byte head + 4 i of the row */
    const uint32_t v = __funnelshift_r(w[i], i + 1 < NW ? w[i + 1] : 0u, sh);
    if (4 * i + 4 <= left) *reinterpret_cast<uint32_t *>(p + 4 * i) = v;
    if (i == (left >> 2)) last = v;
  }
  if (left > 0) {
    uint8_t *t = p + (left & ~3);
#pragma unroll
    for (int b = 0; b < 3; b++) {
      if (b < (left & 3)) t[b] = (uint8_t)(last >> (8 * b));
    }
  }
}

__device__ __forceinline__ void stg32_stream(uint8_t *p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
/* All 48 bytes of w[0..11] to dst, whatever its alignment, with the widest stores the address allows: bytes
 * up to the first 4-byte boundary (hb of them), single words up to the first 16-byte boundary (nh of them),
 * two aligned 128-bit stores, then the words and bytes that are left.  x[i] = the four bytes that follow
 * byte hb + 4 i of the row; which of them start the 128-bit stores depends on nh, hence the switch (all
 * lanes of a half-task share hb and nh: a lane's 48 bytes keep the alignment of the run's first byte). */
__device__ __forceinline__ void tk_store48(uint8_t *dst, const uint32_t (&w)[12]) {
  const uint32_t a = (uint32_t)(uintptr_t)dst & 15u;
  if (a == 0u) {
    stg128_stream(dst, make_uint4(w[0], w[1], w[2], w[3]));
    stg128_stream(dst + 16, make_uint4(w[4], w[5], w[6], w[7]));
    stg128_stream(dst + 32, make_uint4(w[8], w[9], w[10], w[11]));
    return;
  }
  const uint32_t hb = (4u - a) & 3u;
  if (hb & 1u) dst[0] = (uint8_t)w[0];
  if (hb & 2u) *reinterpret_cast<uint16_t *>(dst + (hb & 1u)) = (uint16_t)(w[0] >> (8u * (hb & 1u)));
  uint8_t *p = dst + hb;                                             /* 4-byte aligned */
  const uint32_t nh = ((16u - ((uint32_t)(uintptr_t)p & 15u)) & 15u) >> 2;   /* 0..3 */
  uint32_t x[12];
#pragma unroll
  for (int i = 0; i < 12; i++) x[i] = __funnelshift_r(w[i], i + 1 < 12 ? w[i + 1] : 0u, 8u * hb);
  /* full words in x: 12 (hb == 0) or 11, then 4 - hb bytes in x[11] */
  const bool whole = hb == 0u;
#define JGPU_TK_STORE48_CASE(NH)                                                                        \
  {                                                                                                      \
    _Pragma("unroll") for (int i = 0; i < NH; i++) stg32_stream(p + 4 * i, x[i]);                        \
    stg128_stream(p + 4 * NH, make_uint4(x[NH], x[NH + 1], x[NH + 2], x[NH + 3]));                       \
    stg128_stream(p + 4 * NH + 16, make_uint4(x[NH + 4], x[NH + 5], x[NH + 6], x[NH + 7]));              \
    _Pragma("unroll") for (int i = NH + 8; i < 11; i++) stg32_stream(p + 4 * i, x[i]);                   \
    if (NH + 8 <= 11 && whole) stg32_stream(p + 44, x[11]);                                              \
  }
  switch (nh) {
    case 0: JGPU_TK_STORE48_CASE(0) break;
    case 1: JGPU_TK_STORE48_CASE(1) break;
    case 2: JGPU_TK_STORE48_CASE(2) break;
    default: JGPU_TK_STORE48_CASE(3) break;
  }
#undef JGPU_TK_STORE48_CASE
  if (!whole) {
    /* 4 - hb bytes of x[11] at p + 44 */
    const uint32_t tb = 4u - hb;
    if (tb & 2u) *reinterpret_cast<uint16_t *>(p + 44) = (uint16_t)x[11];
    if (tb & 1u) p[44 + (tb & 2u)] = (uint8_t)(x[11] >> (8u * (tb & 2u)));
  }
}

/* The colour / store stage of a tile that is cropped by the image's right or bottom edge, or whose rows are not
 * 16-byte aligned (any width that is not a multiple of 16 pixels): the same arithmetic, stores of whatever
 * alignment each row has.  Only the EDGE instantiations of k_tk carry it (plans with such rows): in the others its
 * registers would weigh on the K warps' main loop. */
template <int HS, int VS, bool GRAY>
__device__ __forceinline__ void tk_edge_rows(uint32_t lstrip, uint32_t crow0, uint32_t crow0_b, uint8_t *dst, int pitch,
                                                 int vis_rows, int nbytes) {
  if (GRAY) {
#pragma unroll 1
    for (int k = 0; k < vis_rows; k++, dst += pitch) {
      const uint4 v = tk_staged_row_bytes(lstrip + 512 * k);
      if (nbytes == 16 && ((uintptr_t)dst & 15) == 0) {
        stg128_stream(dst, v);
      } else {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        tk_store_any<4>(dst, w, nbytes);
      }
    }
  } else {
#pragma unroll 1
    for (int cr = 0; cr * VS < vis_rows; cr++) {
      uint32_t ca[12], cb[12];
      tk_row_offsets<HS>(crow0 + 512u * (uint32_t)cr, crow0_b + 512u * (uint32_t)cr, ca, cb);
#pragma unroll 1
      for (int k = cr * VS; k < cr * VS + VS && k < vis_rows; k++, dst += pitch) {
        uint32_t ya[4], yb[4], w[12];
        tk_staged_row(lstrip + 512u * (uint32_t)k, ya, yb);
        tk_row_words(ya, yb, ca, cb, w);
        if (nbytes == 48) tk_store48(dst, w);
        else tk_store_any<12>(dst, w, nbytes);
      }
    }
  }
}

template <int HS, int VS, bool GRAY, bool WIDE, int OUT, bool EDGE>
__global__ void __launch_bounds__(TkCfg<HS, VS, GRAY, WIDE>::kThreads, 1)
k_tk(const __grid_constant__ CUtensorMap tm_rows,     /* (64, rows), boxes of 16 rows            */
     const __grid_constant__ CUtensorMap tm_pairs,    /* (64, parity, pairs), boxes of 16 pairs  */
     const __grid_constant__ CUtensorMap tm_rows32,   /* the same views with boxes of 32: a task whose two  */
     const __grid_constant__ CUtensorMap tm_pairs32,  /* half-tasks follow each other in the block order    */
     const __grid_constant__ CUtensorMap tm_quads,    /* (64, 4, quads): every fourth block (HS == 4), boxes of 16 */
     const __grid_constant__ CUtensorMap tm_quads32,  /* ... and of 32                                              */
     const WarpTask *__restrict__ tasks, int n_tasks, const uint32_t *__restrict__ qint,
     const uint32_t *__restrict__ wide_flag, uint8_t *__restrict__ rgb, int rgb_aligned,
     uint8_t *__restrict__ yuv, int *__restrict__ claim) {
  using C = TkCfg<HS, VS, GRAY, WIDE>;
  if ((*wide_flag != 0) != WIDE) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  if (smem0 & 1023u) __trap();   /* the 128-byte swizzle needs the boxes 1 KB aligned */

  /* Where this thread's things live; re-derived from %tid.x at every point of use (see k_mcu). */
  struct Geo {
    int lane;
    uint32_t pair;
    uint32_t zone;      /* the pair's landing zone: box A, box B */
    uint32_t misc;      /* the pair's area behind the zones */
    uint32_t mine;      /* misc + 16*lane: this lane's column of every [row][lane] array */
  };
  auto geo = [&]() -> Geo {
    uint32_t tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    Geo g;
    g.lane = (int)(tid & 31);
    g.pair = (tid >> 5) & (kTkPairs - 1);
    g.zone = smem0 + g.pair * kZoneBytes;
    g.misc = smem0 + kTkPairs * kZoneBytes + g.pair * C::kPairMisc;
    g.mine = g.misc + 16u * (uint32_t)g.lane;
    return g;
  };
  auto bar_data = [&](const Geo &g) { return g.misc + C::kOffBar; };
  auto bar_ring = [&](const Geo &g, uint32_t slot) { return g.misc + C::kOffBar + 8 + 8 * slot; };
  auto bar_full = [&](const Geo &g, uint32_t slot) { return g.misc + C::kOffBar + 8 * (1 + kTkRing) + 8 * slot; };
  auto bar_empty = [&](const Geo &g, uint32_t slot) { return g.misc + C::kOffBar + 8 * (1 + kTkRing + kTkStrips) + 8 * slot; };
  auto strip_addr = [&](const Geo &g, uint32_t slot) { return g.mine + C::kOffStrip + slot * kStripBytes; };
  auto desc_addr = [&](const Geo &g, int n) { return g.misc + C::kOffRing + ((uint32_t)n % kTkRing) * (uint32_t)sizeof(WarpTask); };
  auto half_addr = [&](const Geo &g, int n) { return desc_addr(g, n) + (uint32_t)(g.lane >> 4) * (uint32_t)sizeof(McuHalf); };
  auto idx_addr = [&](const Geo &g, int n) { return g.misc + C::kOffIdx + 4u * ((uint32_t)n % kTkRing); };
  /* Local task n of the pair lives in ring slot n % 8.  The K warp announces every local task, whether
   * it exists or not: its index (>= n_tasks: the pair's work ends here) and, if it exists, its
   * descriptor; the slot's mbarrier completes either way. */
  auto announce = [&](const Geo &g, int n, int idx) {   /* K lane 0 only */
    const uint32_t slot = (uint32_t)n % kTkRing;
    sts32(idx_addr(g, n), (uint32_t)idx);
    if (idx < n_tasks) {
      mbar_expect_tx(bar_ring(g, slot), (uint32_t)sizeof(WarpTask));
      bulk_load(desc_addr(g, n), tasks + idx, (uint32_t)sizeof(WarpTask), bar_ring(g, slot));
    } else {
      mbar_arrive(bar_ring(g, slot));
    }
  };
  auto wait_desc = [&](const Geo &g, int n) {
    mbar_wait_hint<JGPU_MCU_WAIT_NS>(bar_ring(g, (uint32_t)n % kTkRing), ((uint32_t)n / kTkRing) & 1u);
  };

  const bool is_t = JGPU_TK_T_HIGH ? threadIdx.x >= 32 * kTkPairs : threadIdx.x < 32 * kTkPairs;
  {
    const Geo g = geo();
    if (is_t && g.lane == 0) {
      for (int i = 0; i < 1 + kTkRing + 2 * kTkStrips; i++) mbar_init(bar_data(g) + 8 * i, (JGPU_TK_ARRIVE_ALL && i >= 1 + kTkRing) ? 32 : 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();
  const int gw = (int)blockIdx.x * kTkPairs + (int)((threadIdx.x >> 5) & (kTkPairs - 1));
  const int nw = (int)gridDim.x * kTkPairs;

  if (is_t) {
    /* ================================ T: coefficients -> strips ================================ */
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::kRegsT));
    if (gw >= n_tasks) return;
    /* Start the loads of step s of local task n (lane 0 only; its descriptor must have landed): per
     * half-task 16 rows of box A and 16 rows of box B (a half-task that does not exist repeats the
     * other one's blocks, see mcu_plan_build), and the step's table(s). */
    auto fire = [&](const Geo &g, int n, int s) {
      const uint32_t d = desc_addr(g, n), bar = bar_data(g), tab = g.misc + C::kOffTab, held = g.misc + C::kOffQ;
      const bool chroma = s < C::kChromaSteps;
      /* the tables stay: slot 0 luma, 1 Cb, 2 Cr; a slot is fetched again only when the task asks for
       * another table than it holds (never, in a batch with one table set).  The slot is free: its
       * last readers were row passes that have completed. */
      const uint32_t c0 = chroma ? 1u : 0u;
      const uint32_t q0 = lds32(d + offsetof(WarpTask, qidx) + 4 * c0), q1 = lds32(d + offsetof(WarpTask, qidx) + 8);
      const bool get0 = lds32(held + 4 * c0) != q0, get1 = chroma && lds32(held + 8) != q1;
      mbar_expect_tx(bar, 2 * kBoxBytes + ((get0 ? 1 : 0) + (get1 ? 1 : 0)) * C::kTabBytes);
      const bool joined = (lds32(d + offsetof(WarpTask, flags)) & 2u) != 0;
      if (joined) {
        /* one box of 32 rows for block A, one for block B */
        if (chroma) {
          int f0 = (int)lds32(d + offsetof(McuHalf, cfirst));
          int f1 = (int)lds32(d + offsetof(McuHalf, cfirst) + 4);
          if (HS != 1) {
            tma_load_2d(g.zone, &tm_rows32, 0, f0, bar);
            tma_load_2d(g.zone + kBoxBytes, &tm_rows32, 0, f1, bar);
          } else {
            f0 += s;
            f1 += s;
            tma_load_3d(g.zone, &tm_pairs32, 0, f0 & 1, f0 >> 1, bar);
            tma_load_3d(g.zone + kBoxBytes, &tm_pairs32, 0, f1 & 1, f1 >> 1, bar);
          }
        } else if (HS == 4) {
          /* lane L owns blocks 4 L + 2 h and 4 L + 2 h + 1 of a 128-block run: every fourth block */
          const int ls = s - C::kChromaSteps;
          const int first = (int)lds32(d + offsetof(McuHalf, yfirst) + 4 * (ls >> 1)) + 2 * (ls & 1);
          tma_load_3d(g.zone, &tm_quads32, 0, first & 3, first >> 2, bar);
          tma_load_3d(g.zone + kBoxBytes, &tm_quads32, 0, (first + 1) & 3, (first + 1) >> 2, bar);
        } else {
          const int first = (int)lds32(d + offsetof(McuHalf, yfirst) + 4 * (s - C::kChromaSteps));
          tma_load_3d(g.zone, &tm_pairs32, 0, first & 1, first >> 1, bar);
          tma_load_3d(g.zone + kBoxBytes, &tm_pairs32, 0, (first + 1) & 1, (first + 1) >> 1, bar);
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const uint32_t hd = d + h * (uint32_t)sizeof(McuHalf);
          const uint32_t dst = g.zone + h * kHalfBoxBytes;
          if (chroma) {
            int f0 = (int)lds32(hd + offsetof(McuHalf, cfirst));
            int f1 = (int)lds32(hd + offsetof(McuHalf, cfirst) + 4);
            if (HS != 1) {
              tma_load_2d(dst, &tm_rows, 0, f0, bar);
              tma_load_2d(dst + kBoxBytes, &tm_rows, 0, f1, bar);
            } else {
              f0 += s;
              f1 += s;
              tma_load_3d(dst, &tm_pairs, 0, f0 & 1, f0 >> 1, bar);
              tma_load_3d(dst + kBoxBytes, &tm_pairs, 0, f1 & 1, f1 >> 1, bar);
            }
          } else if (HS == 4) {
            const int ls = s - C::kChromaSteps;
            const int first = (int)lds32(hd + offsetof(McuHalf, yfirst) + 4 * (ls >> 1)) + 2 * (ls & 1);
            tma_load_3d(dst, &tm_quads, 0, first & 3, first >> 2, bar);
            tma_load_3d(dst + kBoxBytes, &tm_quads, 0, (first + 1) & 3, (first + 1) >> 2, bar);
          } else {
            const int first = (int)lds32(hd + offsetof(McuHalf, yfirst) + 4 * (s - C::kChromaSteps));
            tma_load_3d(dst, &tm_pairs, 0, first & 1, first >> 1, bar);
            tma_load_3d(dst + kBoxBytes, &tm_pairs, 0, (first + 1) & 1, (first + 1) >> 1, bar);
          }
        }
      }
      if (get0) {
        bulk_load(tab + c0 * C::kTabBytes, qint + (size_t)q0 * 64, C::kTabBytes, bar);
        sts32(held + 4 * c0, q0);
      }
      if (get1) {
        bulk_load(tab + 2 * C::kTabBytes, qint + (size_t)q1 * 64, C::kTabBytes, bar);
        sts32(held + 8, q1);
      }
    };
    {
      const Geo g = geo();
      /* the pair's place in its work, kept in shared memory, not in registers: (steps so far, local task,
       * step inside it); every lane writes the same values */
      sts128(g.misc + C::kOffLoop, make_uint4(0u, 0u, 0u, 0u));
      sts128(g.misc + C::kOffQ, make_uint4(~0u, ~0u, ~0u, ~0u));
      __syncwarp();
      if (g.lane == 0) {
        wait_desc(g, 0);   /* (gw < n_tasks: it exists) */
        fire(g, 0, 0);
      }
      __syncwarp();
    }
#pragma unroll 1
    for (;;) {
      bool is_c, active;   /* a chroma step?  does this lane's unit exist / show in this step? */
      {
        const Geo g = geo();
        __syncwarp();   /* (the record below was written by every lane, with the same values) */
        const uint4 at = lds128(g.misc + C::kOffLoop);
        const int step = (int)at.x, n = (int)at.y, s = (int)at.z;
        if (s == 0) {
          wait_desc(g, n);
          if ((int)lds32(idx_addr(g, n)) >= n_tasks) break;
        }
        is_c = s < C::kChromaSteps;
        const uint32_t ha = half_addr(g, n);
        const int u = g.lane & 15;
        if (HS == 4) {
          const int ls = s - C::kChromaSteps;   /* luma step: block row ls >> 1, half ls & 1 */
          if (OUT == kOutYuv) {
            active = 4 * u + (is_c ? 0 : 2 * (ls & 1)) < (int)lds32(ha + offsetof(McuHalf, blocks_left));
          } else {
            const uint2 wr = lds64(ha + offsetof(McuHalf, width_left));   /* width_left, rows_left */
            active = 32 * u + (is_c ? 0 : 16 * (ls & 1)) < (int)wr.x && (is_c || 8 * (ls >> 1) < (int)wr.y);
          }
        } else if (OUT == kOutYuv) {
          active = 2 * u + ((HS == 1 && is_c) ? s : 0) < (int)lds32(ha + offsetof(McuHalf, blocks_left));
        } else {
          const uint2 wr = lds64(ha + offsetof(McuHalf, width_left));   /* width_left, rows_left */
          active = 16 * u + ((HS == 1 && is_c) ? 8 * s : 0) < (int)wr.x && (is_c || 8 * (s - C::kChromaSteps) < (int)wr.y);
        }
        mbar_wait_hint<JGPU_MCU_WAIT_NS>(bar_data(g), (uint32_t)step & 1u);
      }
      {
        pair32 m[8][8];
        if (JGPU_TK_EXPERIMENT == 2) active = false;
        if (active) {
          const Geo g = geo();
          const uint32_t qa = g.misc + C::kOffTab + (is_c ? C::kTabBytes : 0);   /* luma / Cb */
          const uint32_t qb = is_c ? qa + C::kTabBytes : qa;                       /* luma / Cr */
          mcu_row_pass<WIDE, C::kPark>(m, g.zone, g.lane, qa, qb, g.mine + C::kOffPark);
        }
        /* the boxes are in registers: start the loads of the pair's next step */
        __syncwarp();
        uint32_t strip;
        {
          const Geo g = geo();
          const uint4 at = lds128(g.misc + C::kOffLoop);
          if (g.lane == 0) {
            if ((int)at.z + 1 < C::kSteps) {
              fire(g, (int)at.y, (int)at.z + 1);
            } else {
              const int n = (int)at.y + 1;
              wait_desc(g, n);
              if ((int)lds32(idx_addr(g, n)) < n_tasks) fire(g, n, 0);
            }
          }
          /* this step's strip: free once the K warp has handed back its previous contents */
          const uint32_t cur = at.x, slot = cur % kTkStrips;
          mbar_wait_hint<JGPU_MCU_WAIT_NS>(bar_empty(g, slot), ((cur / kTkStrips) & 1u) ^ 1u);
          strip = strip_addr(g, slot);
        }
        if (active) {
          /* One sink for both kinds of step, without a branch, so that the whole column pass is one
           * basic block and the clamping / packing of a column pair (ALU pipe) is scheduled between
           * the butterflies of the next one (FMA pipe); ptxas does interleave them when it can.
           *   luma:   (short)floor, +128, clamp to 0..255; bytes (A2j A2j+1 B2j B2j+1)
           *   chroma: (short)floor, clamp to -128..127 (= clamp(v+128, 0, 255) - 128, src/xjpeg.c:578);
           *           bytes (Cb2j Cr2j Cb2j+1 Cr2j+1): Cb rides in the low lanes, Cr in the high ones
           * Two column steps make 8 bytes of the strip's row. */
          const pair32 magic = p_make_bits(kMagicBits, kMagicBits);
          const uint32_t add2 = is_c ? 0u : 0x00800080u, lim2 = is_c ? 0xff80ff80u : 0u, sel = is_c ? 0x6240u : 0x6420u;
          uint32_t keep_a[8];   /* the even column step's words, until the odd one completes them */
          auto sink = [&](int j, pair32 (&u)[8], pair32 (&v)[8]) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
              uint32_t ulo, uhi, vlo, vhi;
              p_split_bits(p_add_rm(u[k], magic), ulo, uhi);
              p_split_bits(p_add_rm(v[k], magic), vlo, vhi);
              uint32_t wa = __byte_perm(ulo, vlo, 0x5410), wb = __byte_perm(uhi, vhi, 0x5410);   /* (short)floor, x2 */
              wa = __viaddmin_s16x2(wa, 0u, 0x007f007fu);
              wb = __viaddmin_s16x2(wb, 0u, 0x007f007fu);
              wa = __viaddmax_s16x2(wa, add2, lim2);
              wb = __viaddmax_s16x2(wb, add2, lim2);
              const uint32_t w = __byte_perm(wa, wb, sel);
              if (JGPU_TK_STS32) sts32(strip + 512 * k + 4 * j, w);
              else if ((j & 1) == 0) keep_a[k] = w;
              else sts64(strip + 512 * k + 8 * (j >> 1), make_uint2(keep_a[k], w));
            }
          };
          if (C::kPark) {
            const uint32_t park = geo().mine + C::kOffPark;   /* (this lane's column) */
            auto row0 = [&](int j, pair32 &a, pair32 &b) {
              const uint4 c = lds128(park + 512 * j);
              a = p_make_bits(c.x, c.y);
              b = p_make_bits(c.z, c.w);
            };
            column_pass_by_pairs_parked(m, row0, sink);
          } else {
            column_pass_by_pairs(m, sink);
          }
        }
      }
      __syncwarp();
      {
        const Geo g = geo();
        const uint4 at = lds128(g.misc + C::kOffLoop);
        if (JGPU_TK_ARRIVE_ALL || g.lane == 0) mbar_arrive(bar_full(g, at.x % kTkStrips));
        __syncwarp();
        const bool last = (int)at.z + 1 == C::kSteps;
        sts128(g.misc + C::kOffLoop, make_uint4(at.x + 1u, last ? at.y + 1u : at.y, last ? 0u : at.z + 1u, 0u));
      }
    }
    return;
  }

  /* ================================ K: strips -> pixels / planes ================================ */
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::kRegsK));
  if (gw >= n_tasks) return;
  const Geo g = geo();
  if (g.lane == 0) {
    for (int i = 0; i < kTkAhead; i++) announce(g, i, gw + i * nw);
  }
  const int u = g.lane & 15;   /* this lane's unit inside its half-task */
  auto wait_full = [&](uint32_t st) {
    const uint32_t bar = bar_full(g, st % kTkStrips), parity = (st / kTkStrips) & 1u;
    for (uint32_t spins = 0; !mbar_try_wait_hint<0>(bar, parity); spins++) {
      if (JGPU_TK_SPIN_NS) __nanosleep(JGPU_TK_SPIN_NS);
      if (spins > (1u << 23)) __trap();
    }
  };
  auto hand_back = [&](uint32_t st) {   /* after __syncwarp(): every lane has read what it needs of the strip */
    if (JGPU_TK_ARRIVE_ALL || g.lane == 0) mbar_arrive(bar_empty(g, st % kTkStrips));
  };
  uint32_t step = 0;
#pragma unroll 1
  for (int n = 0;; n++, step += C::kSteps) {
    wait_desc(g, n);
    if ((int)lds32(idx_addr(g, n)) >= n_tasks) break;
    /* claim the task four further on now, pick the answer up when this task's first rows are out */
    int claimed = 0;
    if (g.lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(claimed) : "l"(claim) : "memory");
    const uint32_t da = desc_addr(g, n), ha = half_addr(g, n);
    const uint2 b0 = lds64(ha + offsetof(McuHalf, base0));
    const long long base0 = (long long)(((unsigned long long)b0.y << 32) | b0.x);
    const int pitch = (int)lds32(da + offsetof(WarpTask, pitch0));

    if (OUT == kOutYuv) {
      const int blocks_left = (int)lds32(ha + offsetof(McuHalf, blocks_left));
#pragma unroll 1
      for (int s = 0; s < C::kSteps; s++) {
        const uint32_t st = step + (uint32_t)s;
        wait_full(st);
        const uint32_t strip = strip_addr(g, st % kTkStrips);
        if (s < C::kChromaSteps) {
          /* 8 Cb bytes and 8 Cr bytes per row */
          if ((HS == 4 ? 4 * u : 2 * u + (HS == 1 ? s : 0)) < blocks_left) {
            const uint2 b1 = lds64(ha + offsetof(McuHalf, base1));
            const uint2 dl = lds64(da + offsetof(WarpTask, cr_delta));
            const int cpitch = (int)lds32(da + offsetof(WarpTask, pitch1));
            const long long col = HS != 1 ? 8 * u : 16 * u + 8 * s;
            uint8_t *pb = yuv + (long long)(((unsigned long long)b1.y << 32) | b1.x) + col;
            uint8_t *pr = pb + (long long)(((unsigned long long)dl.y << 32) | dl.x);
#pragma unroll 2
            for (int k = 0; k < 8; k++) {
              const uint4 t = lds128(strip + 512 * k);
              uint2 vb, vr;
              vb.x = __byte_perm(t.x, t.y, 0x6420) ^ 0x80808080u;
              vb.y = __byte_perm(t.z, t.w, 0x6420) ^ 0x80808080u;
              vr.x = __byte_perm(t.x, t.y, 0x7531) ^ 0x80808080u;
              vr.y = __byte_perm(t.z, t.w, 0x7531) ^ 0x80808080u;
              stg64_stream(pb, vb);
              stg64_stream(pr, vr);
              pb += cpitch;
              pr += cpitch;
            }
          }
        } else if ((HS == 4 ? 4 * u + 2 * ((s - C::kChromaSteps) & 1) : 2 * u) < blocks_left) {
          /* 16 Y bytes per row (8 when the unit's second block lies beyond the padded plane) */
          const int ls = s - C::kChromaSteps;
          const int yr = HS == 4 ? ls >> 1 : ls;
          const int blk = HS == 4 ? 4 * u + 2 * (ls & 1) : 2 * u;   /* first of the two blocks, in the half-task */
          const bool whole = blk + 1 < blocks_left;
          const bool wide16 = whole && (pitch & 8) == 0;   /* an odd number of blocks per row: rows are only 8-byte aligned */
          uint8_t *py = yuv + base0 + (long long)(8 * yr) * pitch + 8 * blk;
#pragma unroll 2
          for (int k = 0; k < 8; k++) {
            const uint4 v = tk_staged_row_bytes(strip + 512 * k);
            if (wide16) {
              stg128_stream(py, v);
            } else {
              stg64_stream(py, make_uint2(v.x, v.y));
              if (whole) stg64_stream(py + 8, make_uint2(v.z, v.w));
            }
            py += pitch;
          }
        }
        __syncwarp();
        hand_back(st);
        if (s == 0 && g.lane == 0) announce(g, n + kTkAhead, kTkAhead * nw + claimed);
      }
      continue;
    }

    /* ---- pixels: colour offsets, pack, store ------------------------------------------------- */
    const uint2 wr = lds64(ha + offsetof(McuHalf, width_left));   /* width_left, rows_left */
    const bool rows_aligned = (lds32(da + offsetof(WarpTask, flags)) & (uint32_t)rgb_aligned & 1u) != 0;
    int vis_px = min(16, (int)wr.x - 16 * u);   /* (HS == 4: per luma step, below) */
    bool fast = rows_aligned && vis_px == 16;
    int nbytes = C::kChannels * vis_px;
#pragma unroll 1
    for (int s = 0; s < C::kChromaSteps; s++) wait_full(step + (uint32_t)s);
    const uint32_t cstrip = GRAY ? 0u : strip_addr(g, step % kTkStrips);                       /* HS == 2: the unit's chroma; else the even MCU's */
    const uint32_t cstrip_b = HS != 1 ? cstrip : strip_addr(g, (step + 1u) % kTkStrips);        /* the odd MCU's */
#pragma unroll 1
    for (int ls = 0; ls < C::kLumaSteps; ls++) {
      const uint32_t st = step + (uint32_t)(C::kChromaSteps + ls);
      wait_full(st);
      const uint32_t lstrip = strip_addr(g, st % kTkStrips);
      const int yr = HS == 4 ? ls >> 1 : ls;          /* luma block row of the MCU row */
      const int xoff = HS == 4 ? 16 * (ls & 1) : 0;   /* HS == 4: which half of the 32-pixel unit */
      if (HS == 4) {
        vis_px = min(16, (int)wr.x - 32 * u - xoff);
        fast = rows_aligned && vis_px == 16;
        nbytes = C::kChannels * vis_px;
      }
      const int vis_rows = min(8, (int)wr.y - 8 * yr);
      uint8_t *dst = rgb + base0 + (long long)(8 * yr) * pitch + (long long)(C::kUnitPx * u + xoff) * C::kChannels;
      /* chroma rows this luma block row uses, one per VS pixel rows (HS == 4: the half's 8 bytes of each) */
      const uint32_t crow0 = cstrip + (VS == 2 ? 4u * 512u * (uint32_t)yr : 0u) + (HS == 4 ? (uint32_t)xoff / 2u : 0u);
      const uint32_t crow0_b = cstrip_b + (VS == 2 ? 4u * 512u * (uint32_t)yr : 0u);
      if (vis_px > 0 && vis_rows > 0 && JGPU_TK_EXPERIMENT != 1) {
        if (fast && vis_rows == 8) {
          /* the whole 16 x 8 tile shows and its rows are 16-byte aligned: no per-row tests */
          if (GRAY) {
#pragma unroll 2
            for (int k = 0; k < 8; k++, dst += pitch) stg128_stream(dst, tk_staged_row_bytes(lstrip + 512 * k));
          } else {
#pragma unroll 1
            for (int cr = 0; cr < 8 / VS; cr++) {
              uint32_t ca[12], cb[12];
              tk_row_offsets<HS>(crow0 + 512u * (uint32_t)cr, crow0_b + 512u * (uint32_t)cr, ca, cb);
#pragma unroll
              for (int sub = 0; sub < VS; sub++) {
                uint32_t ya[4], yb[4], w[12];
                tk_staged_row(lstrip + 512u * (uint32_t)(cr * VS + sub), ya, yb);
                tk_row_words(ya, yb, ca, cb, w);
                stg128_stream(dst, make_uint4(w[0], w[1], w[2], w[3]));
                stg128_stream(dst + 16, make_uint4(w[4], w[5], w[6], w[7]));
                stg128_stream(dst + 32, make_uint4(w[8], w[9], w[10], w[11]));
                dst += pitch;
              }
            }
          }
        } else {
          /* edge tiles: cropped by the image's right or bottom edge, or rows that are not 16-byte aligned
           * (any width that is not a multiple of 16 pixels): the same arithmetic, stores of whatever
           * alignment the row has */
          if (EDGE) {
            /* the instantiation for plans with rows of any alignment: widest stores each row allows */
            tk_edge_rows<HS, VS, GRAY>(lstrip, crow0, crow0_b, dst, pitch, vis_rows, nbytes);
          } else if (GRAY) {
            /* (plans whose rows are all 16-byte aligned get here for cropped tiles only: the short form,
             * which leaves the K warps' main loop its registers; 2.3 % on the 4K 4:2:0 batch) */
#pragma unroll 1
            for (int k = 0; k < vis_rows; k++, dst += pitch) {
              const uint4 v = tk_staged_row_bytes(lstrip + 512 * k);
              const uint32_t w[12] = {v.x, v.y, v.z, v.w, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
              tk_store_any_short(dst, w, nbytes);
            }
          } else {
#pragma unroll 1
            for (int cr = 0; cr * VS < vis_rows; cr++) {
              uint32_t ca[12], cb[12];
              tk_row_offsets<HS>(crow0 + 512u * (uint32_t)cr, crow0_b + 512u * (uint32_t)cr, ca, cb);
#pragma unroll 1
              for (int k = cr * VS; k < cr * VS + VS && k < vis_rows; k++, dst += pitch) {
                uint32_t ya[4], yb[4], w[12];
                tk_staged_row(lstrip + 512u * (uint32_t)k, ya, yb);
                tk_row_words(ya, yb, ca, cb, w);
                tk_store_any_short(dst, w, nbytes);
              }
            }
          }
        }
      }
      __syncwarp();
      hand_back(st);
      if (ls == 0 && g.lane == 0) announce(g, n + kTkAhead, kTkAhead * nw + claimed);
    }
    /* (the __syncwarp() before the last hand_back covers the chroma strips too) */
#pragma unroll 1
    for (int s = 0; s < C::kChromaSteps; s++) hand_back(step + (uint32_t)s);
  }
}

/* ---- host side ------------------------------------------------------------- */

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn mcu_encode_fn() {
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

struct McuMode {
  int hs, vs, gray, channels;
  int threads, warps;             /* the kernel for 8-bit tables */
  size_t smem;
  int threads_wide, warps_wide;   /* the kernel for 16-bit tables (more table bytes per warp) */
  size_t smem_wide;
  size_t tk_smem, tk_smem_wide;   /* k_tk: 8 warp pairs per CTA either way */
};
/* Which kernel runs.  k_tk, except where the colour warps have next to nothing to do and k_mcu's twelve
 * do-everything warps are faster (measured, profiles/r2_notes.md section 11): grey pixels (when the rows are 16-byte
 * aligned: k_mcu stores other rows byte by byte) and the planes of grey, 4:2:0 and 4:2:2.  JGPU_KERNEL=mcu / tk
 * forces one of them (A/B runs; 4:1:1 exists in k_tk only); read when a context is created (mcu_configure). */
int g_kernel_choice = 0;   /* 0: as above, 1: k_tk, 2: k_mcu */
bool mcu_use_tk(int mode, bool planes, bool edge) {
  if (mode >= kMode411 || g_kernel_choice == 1) return true;
  if (g_kernel_choice == 2) return false;
  if (mode == kModeGray) return planes ? false : edge;
  if (planes) return !(mode == kMode420 || mode == kMode422);
  return true;
}
McuMode g_mcu[kNumFusedModes];
bool g_mcu_configured = false;

template <int HS, int VS, bool GRAY>
cudaError_t mcu_configure_mode(int mode) {
  McuMode &mi = g_mcu[mode];
  mi.hs = HS; mi.vs = VS; mi.gray = GRAY; mi.channels = GRAY ? 1 : 3;
  cudaError_t e;
  if constexpr (HS != 4) {   /* (k_mcu has no 4x luma modes: they came with k_tk) */
    using C = McuCfg<HS, VS, GRAY, false>;
    using CW = McuCfg<HS, VS, GRAY, true>;
    mi.threads = C::kThreads; mi.warps = C::kWarps; mi.smem = C::kSmemBytes;
    mi.threads_wide = CW::kThreads; mi.warps_wide = CW::kWarps; mi.smem_wide = CW::kSmemBytes;
    if ((e = cudaFuncSetAttribute(&k_mcu<HS, VS, GRAY, false, kOutRgb>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(&k_mcu<HS, VS, GRAY, true, kOutRgb>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.smem_wide)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(&k_mcu<HS, VS, GRAY, false, kOutYuv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(&k_mcu<HS, VS, GRAY, true, kOutYuv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.smem_wide)) != cudaSuccess) return e;
  }
  mi.tk_smem = TkCfg<HS, VS, GRAY, false>::kSmemBytes;
  mi.tk_smem_wide = TkCfg<HS, VS, GRAY, true>::kSmemBytes;
  if ((e = cudaFuncSetAttribute(&k_tk<HS, VS, GRAY, false, kOutRgb, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.tk_smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(&k_tk<HS, VS, GRAY, true, kOutRgb, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.tk_smem_wide)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(&k_tk<HS, VS, GRAY, false, kOutRgb, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.tk_smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(&k_tk<HS, VS, GRAY, true, kOutRgb, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.tk_smem_wide)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(&k_tk<HS, VS, GRAY, false, kOutYuv, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.tk_smem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(&k_tk<HS, VS, GRAY, true, kOutYuv, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mi.tk_smem_wide)) != cudaSuccess) return e;
  return cudaSuccess;
}

template <int HS, int VS, bool GRAY>
cudaError_t mcu_launch_mode(bool use_tk, bool planes, bool edge, int sm_count, const McuMode &mi, cudaStream_t stream, const CUtensorMap &tm_rows,
                            const CUtensorMap &tm_pairs, const CUtensorMap &tm_rows32, const CUtensorMap &tm_pairs32, const CUtensorMap &tm_quads,
                            const CUtensorMap &tm_quads32, const WarpTask *tasks, int n_tasks, const uint32_t *qint,
                            const uint32_t *wide_flag, uint8_t *rgb, int rgb_aligned, uint8_t *yuv, int *claim) {
  if (use_tk) {
    const int grid = std::min((n_tasks + kTkPairs - 1) / kTkPairs, sm_count);
    const int threads = 2 * 32 * kTkPairs;
    if (planes) {
      k_tk<HS, VS, GRAY, false, kOutYuv, false><<<grid, threads, mi.tk_smem, stream>>>(tm_rows, tm_pairs, tm_rows32, tm_pairs32, tm_quads, tm_quads32, tasks, n_tasks, qint, wide_flag, rgb, rgb_aligned, yuv, claim);
      k_tk<HS, VS, GRAY, true, kOutYuv, false><<<grid, threads, mi.tk_smem_wide, stream>>>(tm_rows, tm_pairs, tm_rows32, tm_pairs32, tm_quads, tm_quads32, tasks, n_tasks, qint, wide_flag, rgb, rgb_aligned, yuv, claim);
    } else if (edge) {
      k_tk<HS, VS, GRAY, false, kOutRgb, true><<<grid, threads, mi.tk_smem, stream>>>(tm_rows, tm_pairs, tm_rows32, tm_pairs32, tm_quads, tm_quads32, tasks, n_tasks, qint, wide_flag, rgb, rgb_aligned, yuv, claim);
      k_tk<HS, VS, GRAY, true, kOutRgb, true><<<grid, threads, mi.tk_smem_wide, stream>>>(tm_rows, tm_pairs, tm_rows32, tm_pairs32, tm_quads, tm_quads32, tasks, n_tasks, qint, wide_flag, rgb, rgb_aligned, yuv, claim);
    } else {
      k_tk<HS, VS, GRAY, false, kOutRgb, false><<<grid, threads, mi.tk_smem, stream>>>(tm_rows, tm_pairs, tm_rows32, tm_pairs32, tm_quads, tm_quads32, tasks, n_tasks, qint, wide_flag, rgb, rgb_aligned, yuv, claim);
      k_tk<HS, VS, GRAY, true, kOutRgb, false><<<grid, threads, mi.tk_smem_wide, stream>>>(tm_rows, tm_pairs, tm_rows32, tm_pairs32, tm_quads, tm_quads32, tasks, n_tasks, qint, wide_flag, rgb, rgb_aligned, yuv, claim);
    }
    return cudaGetLastError();
  }
  if constexpr (HS != 4) {
    const int grid = std::min((n_tasks + mi.warps - 1) / mi.warps, sm_count);
    const int grid_w = std::min((n_tasks + mi.warps_wide - 1) / mi.warps_wide, sm_count);
    if (planes) {
      k_mcu<HS, VS, GRAY, false, kOutYuv><<<grid, mi.threads, mi.smem, stream>>>(tm_rows, tm_pairs, tasks, n_tasks, qint, wide_flag, rgb, rgb_aligned, yuv, claim);
      k_mcu<HS, VS, GRAY, true, kOutYuv><<<grid_w, mi.threads_wide, mi.smem_wide, stream>>>(tm_rows, tm_pairs, tasks, n_tasks, qint, wide_flag, rgb, rgb_aligned, yuv, claim);
    } else {
      k_mcu<HS, VS, GRAY, false, kOutRgb><<<grid, mi.threads, mi.smem, stream>>>(tm_rows, tm_pairs, tasks, n_tasks, qint, wide_flag, rgb, rgb_aligned, yuv, claim);
      k_mcu<HS, VS, GRAY, true, kOutRgb><<<grid_w, mi.threads_wide, mi.smem_wide, stream>>>(tm_rows, tm_pairs, tasks, n_tasks, qint, wide_flag, rgb, rgb_aligned, yuv, claim);
    }
    return cudaGetLastError();
  }
  return cudaErrorInvalidValue;
}

}  // namespace

#ifdef JGPU_MCU_TRACE
extern "C" int jgpu_mcu_trace_read(unsigned long long *out, int n_words) {
  return cudaMemcpyFromSymbol(out, g_mcu_trace, sizeof(unsigned long long) * (size_t)n_words) == cudaSuccess ? 0 : 1;
}
#endif

bool mcu_has_mode(int mode) { return mode < kMode411 || g_kernel_choice != 2; }

cudaError_t mcu_configure(int device) {
  (void)device;
  {
    const char *k = getenv("JGPU_KERNEL");
    g_kernel_choice = !k ? 0 : strcmp(k, "tk") == 0 ? 1 : strcmp(k, "mcu") == 0 ? 2 : 0;
  }
  cudaError_t e;
  if ((e = mcu_configure_mode<1, 1, true>(kModeGray)) != cudaSuccess) return e;
  if ((e = mcu_configure_mode<1, 1, false>(kMode444)) != cudaSuccess) return e;
  if ((e = mcu_configure_mode<2, 1, false>(kMode422)) != cudaSuccess) return e;
  if ((e = mcu_configure_mode<2, 2, false>(kMode420)) != cudaSuccess) return e;
  if ((e = mcu_configure_mode<1, 2, false>(kMode440)) != cudaSuccess) return e;
  if ((e = mcu_configure_mode<4, 1, false>(kMode411)) != cudaSuccess) return e;
  g_mcu_configured = true;
  return cudaSuccess;
}

struct McuPlanImpl {
  int n = 0;
  int sm_count = 0;
  bool planes = false;   /* tasks cover the padded planes (YUV output) instead of the visible pixels */
  void *d_tasks[kNumFusedModes] = {};
  int n_tasks[kNumFusedModes] = {};
  bool any_unaligned[kNumFusedModes] = {};   /* the mode has images whose pixel rows are not all 16-byte aligned */
  std::vector<int> first_task[kNumFusedModes]; /* per mode, n+1 entries */
  void *d_qint = nullptr;
  int qint_cap = 0; /* tables */
  /* task counters: one int per kernel launch, taken round-robin from a ring (launches of one plan may
   * be in flight on several streams: the chunks of jgpu_decode_batch_host) and zeroed on the launch's
   * stream just before it */
  int *d_claim = nullptr;
  unsigned claim_next = 0;
  long long coef_rows = 0; /* 128-byte rows the batch touches */
  const void *map_ptr = nullptr;
  CUtensorMap tm_rows, tm_pairs;       /* boxes of 16 rows / pairs */
  CUtensorMap tm_rows32, tm_pairs32;   /* boxes of 32 */
  CUtensorMap tm_quads, tm_quads32;    /* (64, 4, quads): every fourth block (4x luma modes) */
  cudaStream_t side[kNumFusedModes] = {};
  cudaEvent_t fork = nullptr, join[kNumFusedModes] = {};
};

int mcu_plan_build(FusedPlan &fp, const jgpu_image_desc *descs, const jgpu_layout *layouts,
                   const int *modes, int n, unsigned flags, int sm_count) {
  if (!g_mcu_configured) return jgpu_fail("fused kernel not configured");
  if (!mcu_encode_fn()) return jgpu_fail("cuTensorMapEncodeTiled is not available from this driver");
  McuPlanImpl *p = new McuPlanImpl();
  fp.impl = p;
  p->n = n;
  p->sm_count = sm_count;
  p->planes = (flags & JGPU_OUT_YUV) != 0;
  std::vector<WarpTask> tasks[kNumFusedModes];
  for (int m = 0; m < kNumFusedModes; m++) p->first_task[m].assign(n + 1, 0);
  for (int i = 0; i < n; i++) {
    const jgpu_image_desc &d = descs[i];
    const jgpu_layout &lay = layouts[i];
    const McuMode &mi = g_mcu[modes[i]];
    int block0[3] = {0, 0, 0}, hblocks[3] = {0, 0, 0};
    for (int c = 0; c < d.ncomps; c++) {
      block0[c] = (int)((d.coef_off + lay.plane[c].coef_off) / 64);
      hblocks[c] = lay.plane[c].hblocks;
    }
    p->coef_rows = std::max<long long>(p->coef_rows, (d.coef_off + lay.coef_len + 63) / 64);
    for (int m = 0; m < kNumFusedModes; m++) p->first_task[m][i] = (int)tasks[m].size();
    /* the image's half-tasks in row-major order: MCU rows x runs of 16 units (32 luma blocks) */
    const int mcu_h = mi.gray ? 8 : 8 * mi.vs;
    const int nv = mi.gray ? lay.plane[0].vblocks : lay.nvmb;
    const int lrows = mi.gray ? 1 : mi.vs;                /* luma block rows per MCU row */
    const int lblocks = hblocks[0];                       /* luma blocks per block row */
    const int cper = mi.gray ? 0 : (mi.hs == 1 ? 32 : 16);  /* chroma blocks a half-task spans */
    const int lper = mi.hs == 4 ? 64 : 32;                  /* luma blocks a half-task spans: 16 units */
    std::vector<McuHalf> halves;
    for (int r = 0; r < nv; r++) {
      for (int b = 0; b < lblocks; b += lper) {
        const int x0 = b * 8, y0 = r * mcu_h;
        /* half-tasks wholly to the right of / below the visible image carry no pixels */
        if (!p->planes && (x0 >= d.width || y0 >= d.height)) continue;
        McuHalf h;
        memset(&h, 0, sizeof(h));
        h.width_left = d.width - x0;
        h.rows_left = d.height - y0;
        h.blocks_left = lblocks - b;
        for (int v = 0; v < lrows; v++) h.yfirst[v] = block0[0] + (r * lrows + v) * lblocks + b;
        if (!mi.gray) {
          const int cx = b / lper * cper;
          h.cfirst[0] = block0[1] + r * hblocks[1] + cx;
          h.cfirst[1] = block0[2] + r * hblocks[2] + cx;
        }
        if (p->planes) {
          h.base0 = d.yuv_off + lay.plane[0].data_off + (long long)y0 * lay.plane[0].width + x0;
          if (!mi.gray) {
            h.base1 = d.yuv_off + lay.plane[1].data_off + (long long)(r * 8) * lay.plane[1].width + (b / lper) * cper * 8;
          }
        } else {
          h.base0 = d.rgb_off + ((long long)y0 * d.width + x0) * mi.channels;
        }
        halves.push_back(h);
      }
    }
    for (size_t k = 0; k < halves.size(); k += 2) {
      WarpTask t;
      memset(&t, 0, sizeof(t));
      t.half[0] = halves[k];
      if (k + 1 < halves.size()) {
        t.half[1] = halves[k + 1];
      } else {
        /* no second half: its lanes stay idle; the loads repeat the first half's blocks */
        t.half[1] = halves[k];
        t.half[1].width_left = t.half[1].rows_left = t.half[1].blocks_left = 0;
      }
      if (p->planes) {
        t.pitch0 = lay.plane[0].width;
        if (!mi.gray) {
          t.pitch1 = lay.plane[1].width;
          t.cr_delta = lay.plane[2].data_off - lay.plane[1].data_off;
        }
      } else {
        t.pitch0 = d.width * mi.channels;
        /* every half-task starts at a multiple of 256 pixels of a row: aligned iff the image is */
        t.flags = ((d.rgb_off & 15) == 0 && (t.pitch0 & 15) == 0) ? 1 : 0;
        if (!t.flags) p->any_unaligned[modes[i]] = true;
      }
      /* bit 1: the second half-task follows the first in the block order of every plane (the next 16
       * units of the same MCU row): k_tk fetches one box of 32 rows instead of two of 16 */
      if (k + 1 < halves.size()) {
        const McuHalf &a = t.half[0], &b = t.half[1];
        bool joined = true;
        for (int v = 0; v < lrows; v++) joined = joined && b.yfirst[v] == a.yfirst[v] + lper;
        if (!mi.gray) joined = joined && b.cfirst[0] == a.cfirst[0] + cper && b.cfirst[1] == a.cfirst[1] + cper;
        if (joined) t.flags |= 2;
      }
      for (int c = 0; c < d.ncomps; c++) t.qidx[c] = d.qtab_set * 4 + d.tq[c];
      tasks[modes[i]].push_back(t);
    }
  }
  for (int m = 0; m < kNumFusedModes; m++) p->first_task[m][n] = (int)tasks[m].size();
  for (int m = 0; m < kNumFusedModes; m++) {
    p->n_tasks[m] = (int)tasks[m].size();
    if (tasks[m].empty()) continue;
    if (cudaMalloc(&p->d_tasks[m], sizeof(WarpTask) * tasks[m].size()) != cudaSuccess ||
        cudaMemcpy(p->d_tasks[m], tasks[m].data(), sizeof(WarpTask) * tasks[m].size(),
                   cudaMemcpyHostToDevice) != cudaSuccess) {
      return jgpu_fail("fused plan: task descriptor upload failed");
    }
  }
  return 0;
}

void mcu_plan_release(FusedPlan &fp) {
  McuPlanImpl *p = static_cast<McuPlanImpl *>(fp.impl);
  if (!p) return;
  for (int m = 0; m < kNumFusedModes; m++) {
    cudaFree(p->d_tasks[m]);
    if (p->side[m]) cudaStreamDestroy(p->side[m]);
    if (p->join[m]) cudaEventDestroy(p->join[m]);
  }
  if (p->fork) cudaEventDestroy(p->fork);
  cudaFree(p->d_qint);
  cudaFree(p->d_claim);
  delete p;
  fp.impl = nullptr;
}

int mcu_plan_launches(const FusedPlan &fp) {
  const McuPlanImpl *p = static_cast<const McuPlanImpl *>(fp.impl);
  int n = 1; /* table conversion */
  for (int m = 0; m < kNumFusedModes; m++) n += 2 * (p->n_tasks[m] > 0); /* 8-bit and 16-bit table variants */
  return n;
}

static int mcu_build_maps(McuPlanImpl *p, const int16_t *coef) {
  if (p->map_ptr == coef) return 0;
  if (reinterpret_cast<uintptr_t>(coef) & 255) {
    return jgpu_fail("fused path: the coefficient buffer must be 256-byte aligned");
  }
  EncodeTiledFn enc = mcu_encode_fn();
  const cuuint64_t rows = (cuuint64_t)((p->coef_rows + 3) & ~3ll);
  for (int big = 0; big < 2; big++) {
    const cuuint32_t box_rows = big ? 2 * kHalfRows : kHalfRows;
    {
      cuuint64_t dims[2] = {64, rows};
      cuuint64_t strides[1] = {128};
      cuuint32_t box[2] = {64, box_rows};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(big ? &p->tm_rows32 : &p->tm_rows, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<int16_t *>(coef), dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       JGPU_TMA_L2_PROMOTION, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return jgpu_fail("cuTensorMapEncodeTiled(rows) failed (%d)", (int)r);
    }
    {
      cuuint64_t dims[3] = {64, 2, rows / 2};
      cuuint64_t strides[2] = {128, 256};
      cuuint32_t box[3] = {64, 1, box_rows};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = enc(big ? &p->tm_pairs32 : &p->tm_pairs, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<int16_t *>(coef), dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       JGPU_TMA_L2_PROMOTION, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return jgpu_fail("cuTensorMapEncodeTiled(pairs) failed (%d)", (int)r);
    }
    {
      cuuint64_t dims[3] = {64, 4, rows / 4};
      cuuint64_t strides[2] = {128, 512};
      cuuint32_t box[3] = {64, 1, box_rows};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = enc(big ? &p->tm_quads32 : &p->tm_quads, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<int16_t *>(coef), dims,
                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       JGPU_TMA_L2_PROMOTION, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return jgpu_fail("cuTensorMapEncodeTiled(quads) failed (%d)", (int)r);
    }
  }
  p->map_ptr = coef;
  return 0;
}

int mcu_plan_launch(FusedPlan &fp, int i0, int i1, const int16_t *coef, const uint16_t *qtabs,
                    int n_sets, uint8_t *rgb, uint8_t *yuv, cudaStream_t stream) {
  McuPlanImpl *p = static_cast<McuPlanImpl *>(fp.impl);
  if (mcu_build_maps(p, coef)) return 1;
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
  for (int m = 0; m < kNumFusedModes; m++) n_modes += p->first_task[m][i1] > p->first_task[m][i0];
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
    const int t0 = p->first_task[m][i0], t1 = p->first_task[m][i1];
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
    const McuMode &mi = g_mcu[m];
    const int grid = p->sm_count;
    if (!p->d_claim && cudaMalloc(&p->d_claim, sizeof(int) * kClaimSlots) != cudaSuccess) {
      return jgpu_fail("fused path: counter allocation failed");
    }
    int *claim = p->d_claim + (p->claim_next++ % kClaimSlots);
    if (cudaMemsetAsync(claim, 0, sizeof(int), stream) != cudaSuccess) return jgpu_fail("fused path: memset failed");
    const WarpTask *tasks = static_cast<const WarpTask *>(p->d_tasks[m]) + t0;
    const uint32_t *qint = static_cast<const uint32_t *>(p->d_qint);
    switch (m) {
      case kModeGray: e = mcu_launch_mode<1, 1, true>(mcu_use_tk(m, p->planes, p->any_unaligned[m] || !rgb_aligned), p->planes, p->any_unaligned[m] || !rgb_aligned, grid, mi, stream, p->tm_rows, p->tm_pairs, p->tm_rows32, p->tm_pairs32, p->tm_quads, p->tm_quads32, tasks, t1 - t0, qint, wide_flag, rgb, rgb_aligned, yuv, claim); break;
      case kMode444: e = mcu_launch_mode<1, 1, false>(mcu_use_tk(m, p->planes, p->any_unaligned[m] || !rgb_aligned), p->planes, p->any_unaligned[m] || !rgb_aligned, grid, mi, stream, p->tm_rows, p->tm_pairs, p->tm_rows32, p->tm_pairs32, p->tm_quads, p->tm_quads32, tasks, t1 - t0, qint, wide_flag, rgb, rgb_aligned, yuv, claim); break;
      case kMode422: e = mcu_launch_mode<2, 1, false>(mcu_use_tk(m, p->planes, p->any_unaligned[m] || !rgb_aligned), p->planes, p->any_unaligned[m] || !rgb_aligned, grid, mi, stream, p->tm_rows, p->tm_pairs, p->tm_rows32, p->tm_pairs32, p->tm_quads, p->tm_quads32, tasks, t1 - t0, qint, wide_flag, rgb, rgb_aligned, yuv, claim); break;
      case kMode420: e = mcu_launch_mode<2, 2, false>(mcu_use_tk(m, p->planes, p->any_unaligned[m] || !rgb_aligned), p->planes, p->any_unaligned[m] || !rgb_aligned, grid, mi, stream, p->tm_rows, p->tm_pairs, p->tm_rows32, p->tm_pairs32, p->tm_quads, p->tm_quads32, tasks, t1 - t0, qint, wide_flag, rgb, rgb_aligned, yuv, claim); break;
      case kMode440: e = mcu_launch_mode<1, 2, false>(mcu_use_tk(m, p->planes, p->any_unaligned[m] || !rgb_aligned), p->planes, p->any_unaligned[m] || !rgb_aligned, grid, mi, stream, p->tm_rows, p->tm_pairs, p->tm_rows32, p->tm_pairs32, p->tm_quads, p->tm_quads32, tasks, t1 - t0, qint, wide_flag, rgb, rgb_aligned, yuv, claim); break;
      case kMode411: e = mcu_launch_mode<4, 1, false>(mcu_use_tk(m, p->planes, p->any_unaligned[m] || !rgb_aligned), p->planes, p->any_unaligned[m] || !rgb_aligned, grid, mi, stream, p->tm_rows, p->tm_pairs, p->tm_rows32, p->tm_pairs32, p->tm_quads, p->tm_quads32, tasks, t1 - t0, qint, wide_flag, rgb, rgb_aligned, yuv, claim); break;
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
