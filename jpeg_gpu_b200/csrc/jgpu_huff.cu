/* jgpu_huff.cu — Huffman decoding of baseline JPEG scans on the GPU (sm_100a): the kernels
 * around the per-thread loop of jgpu_huff_core.h.  See that header for the method; this file
 * is the mapping to the machine.
 *
 *   k_huff_sync   one CTA = 248 consecutive subsequences of one file plus the 8 before them as
 *                 a warm-up (JGPU_HUFF_WARM).  The file's six decoder
 *                 tables (14.6 KB) and the CTA's 32 KB of scan words are staged in shared
 *                 memory (words XOR-swizzled by subsequence so that 32 threads reading "their
 *                 j-th word" hit 32 banks); states pass from thread to thread through shared
 *                 memory until the CTA is stable, and to the next CTA through a carry array
 *                 that the next launch of the kernel picks up.
 *   k_huff_scan   one CTA per file: segmented exclusive scan of the slot counts (restart
 *                 intervals are the segments).
 *   k_huff_write  the same staging; stores coefficients, verifies the chain of states.
 *   k_huff_dc     one CTA per piece of a (restart interval, component) chain: prefix sum of the
 *                 DC differences, in two launches when chains are longer than a piece.
 */
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include "jgpu_huff.h"
#include "jgpu_internal.h"

namespace jgpu {

namespace {

constexpr int kCta = JGPU_HUFF_CTA;
constexpr int kGuardWords = 4;

__constant__ unsigned char c_zigzag[64] = JGPU_HUFF_ZIGZAG_NATURAL;

constexpr bool kGlobalWords = JGPU_HUFF_GW != 0;

template <int S>
struct SyncSmem {
  jgpu_huff_table tabs[JGPU_HUFF_TABLES];
  uint32_t words[kGlobalWords ? 4 : (kCta + 1) * S];
  uint32_t s_out[kCta + 1];   /* s_out[i + 1]: state subsequence i ends in */
  uint32_t s_in[kCta];        /* state subsequence i starts from (current estimate) */
  uint32_t n[kCta];           /* slots it advances when decoded from s_in */
  uint16_t list[kCta];        /* subsequences to redo this round, compacted */
  int n_list;
  jgpu_huff_file file;
  unsigned char zz[64];
  uint4 brec[JGPU_HUFF_MAX_BLOCKS + 2];   /* block_records() */
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint32_t v;
  asm volatile("{\n\t.reg .u16 t;\n\tld.shared.u16 t, [%1];\n\tcvt.u32.u16 %0, t;\n\t}" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("{\n\t.reg .u16 t;\n\tld.shared.u8 t, [%1];\n\tcvt.u32.u16 %0, t;\n\t}" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}

/* The accessor jgpu_huff_core.h asks for, over 32-bit shared-memory addresses computed once
 * per thread: explicit ld.shared, so the hot loop holds no generic pointer for the compiler to
 * convert (it re-read %ctaid / the shared window base inside the loop when it did). */
template <int S>
struct DevMem {
  uint32_t words;      /* shared address of SyncSmem::words */
  uint32_t base_word;  /* file-relative index of the CTA's first word */
  uint32_t tabs, file, zz;
  uint32_t comps;      /* huff::comp_pack of the file */
  const uint32_t *gwords;   /* kGlobalWords: the file's scan in global memory, as stored */
  uint32_t brec;            /* shared address of the block records */
  /* what the decoding loop needs when a block ends, one load: the block after block c and the
   * shared-memory addresses of its tables */
  __device__ __forceinline__ uint32_t next_block(uint32_t c, uint32_t *tdc, uint32_t *tac) const {
    const uint32_t r = lds_u32(brec + 16u * c);
    *tdc = r & 0xffffffu;
    *tac = (r & 0xffffffu) + (uint32_t)sizeof(jgpu_huff_table);
    return r >> 24;
  }
  struct Cursor {
    const uint32_t *p;
    uint32_t i;
  };
  __device__ __forceinline__ Cursor cursor(uint32_t i) const {
    Cursor cur;
    cur.p = gwords + i;
    cur.i = i;
    return cur;
  }
  __device__ __forceinline__ uint32_t next_word(Cursor &cur) const {
    if (kGlobalWords) return __byte_perm(__ldg(++cur.p), 0, 0x0123);
    return word(++cur.i);
  }
  __device__ __forceinline__ uint32_t word(uint32_t i) const {
    if (kGlobalWords) return __byte_perm(__ldg(gwords + i), 0, 0x0123);
    const uint32_t l = i - base_word, row = l / S;
    return lds_u32(words + 4u * ((l & ~(uint32_t)(S - 1)) | ((l ^ row) & (S - 1))));
  }
  __device__ __forceinline__ uint32_t lut(uint32_t t, uint32_t i) const {
    return lds_u16(tabs + t * (uint32_t)sizeof(jgpu_huff_table) + 2u * i);
  }
  __device__ __forceinline__ uint32_t table_ref(uint32_t t) const { return tabs + t * (uint32_t)sizeof(jgpu_huff_table); }
  __device__ __forceinline__ uint32_t lut_at(uint32_t ref, uint32_t i) const { return lds_u16(ref + 2u * i); }
  __device__ __forceinline__ uint32_t limit(uint32_t t, int len) const {
    return lds_u32(tabs + t * (uint32_t)sizeof(jgpu_huff_table) + (uint32_t)offsetof(jgpu_huff_table, limit) + 4u * len);
  }
  __device__ __forceinline__ int32_t delta(uint32_t t, int len) const {
    return (int32_t)lds_u32(tabs + t * (uint32_t)sizeof(jgpu_huff_table) + (uint32_t)offsetof(jgpu_huff_table, delta) + 4u * len);
  }
  __device__ __forceinline__ uint32_t symbol(uint32_t t, int i) const {
    return lds_u8(tabs + t * (uint32_t)sizeof(jgpu_huff_table) + (uint32_t)offsetof(jgpu_huff_table, symbols) + (uint32_t)i);
  }
  __device__ __forceinline__ uint32_t blk_table(uint32_t c) const { return 2u * ((comps >> (2u * c)) & 3u); }
  __device__ __forceinline__ int64_t blk_base(int c) const {
    return (int64_t)lds_u64(file + (uint32_t)offsetof(jgpu_huff_file, blk_base) + 8u * c);
  }
  __device__ __forceinline__ int32_t blk_xs(int c) const {
    return (int32_t)lds_u32(file + (uint32_t)offsetof(jgpu_huff_file, blk_xs) + 4u * c);
  }
  __device__ __forceinline__ int32_t blk_ys(int c) const {
    return (int32_t)lds_u32(file + (uint32_t)offsetof(jgpu_huff_file, blk_ys) + 4u * c);
  }
  __device__ __forceinline__ uint32_t zigzag(int k) const { return lds_u8(zz + (uint32_t)k); }
};

/* A value the compiler must keep in a register: what comes out of a volatile asm cannot be
 * recomputed, and these addresses are otherwise rebuilt from special registers (S2R, tens of
 * cycles on the short scoreboard) at every use inside the loop. */
__device__ __forceinline__ uint32_t pinned(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

/* Record c describes the block that FOLLOWS block c of the MCU: x = shared address of its DC table
 * | its index << 24; y, z, w = where it lies, in blocks: y + MCU column * z + MCU row * w
 * (jgpu_huff_file_finish, every term a whole number of blocks).  Written by the first threads of
 * the CTA once the file descriptor is in shared memory; a __syncthreads() follows. */
__device__ __forceinline__ void block_records(uint4 *brec, const jgpu_huff_file &f, const jgpu_huff_table *tabs) {
  const int c = threadIdx.x;
  if (c < f.bpm && c < JGPU_HUFF_MAX_BLOCKS) {
    const int nc = c + 1 == f.bpm ? 0 : c + 1;
    const uint32_t t = 2u * (f.blk_comp[nc] & 3u);
    uint4 r;
    r.x = ((uint32_t)__cvta_generic_to_shared(tabs) + t * (uint32_t)sizeof(jgpu_huff_table)) | ((uint32_t)nc << 24);
    r.y = (uint32_t)(f.blk_base[nc] >> 6);
    r.z = (uint32_t)(f.blk_xs[nc] >> 6);
    r.w = (uint32_t)(f.blk_ys[nc] >> 6);
    brec[c] = r;
  }
}

template <int S>
__device__ __forceinline__ DevMem<S> dev_mem(const SyncSmem<S> &sm, int first, const uint32_t *stream) {
  DevMem<S> m;
  m.gwords = stream + sm.file.word0;
  m.words = pinned((uint32_t)__cvta_generic_to_shared(sm.words));
  m.base_word = pinned((uint32_t)first * S);
  m.tabs = pinned((uint32_t)__cvta_generic_to_shared(sm.tabs));
  m.file = pinned((uint32_t)__cvta_generic_to_shared(&sm.file));
  m.zz = pinned((uint32_t)__cvta_generic_to_shared(sm.zz));
  m.comps = pinned(huff::comp_pack(sm.file));
  m.brec = pinned((uint32_t)__cvta_generic_to_shared(sm.brec));
  return m;
}

/* %ctaid.x, read once (a plain blockIdx.x is cheap to re-read, so the compiler does, in loops) */
__device__ __forceinline__ int cta_x() {
  int v;
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(v));
  return v;
}

/* Stages the file descriptor, its tables and the CTA's words.  `count` subsequences. */
template <int S>
__device__ __forceinline__ void stage(SyncSmem<S> &sm, const jgpu_huff_file *files, const uint32_t *stream,
                                      const jgpu_huff_table *tables, int first, int count) {
  const int t = threadIdx.x;
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(files + blockIdx.y);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&sm.file);
    for (int i = t; i < (int)(sizeof(jgpu_huff_file) / 4); i += kCta) dst[i] = src[i];
    if (t < 64) sm.zz[t] = c_zigzag[t];
  }
  __syncthreads();
  block_records(sm.brec, sm.file, sm.tabs);
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(tables + sm.file.table0);
    uint4 *dst = reinterpret_cast<uint4 *>(sm.tabs);
    for (int i = t; i < (int)(sizeof(jgpu_huff_table) * JGPU_HUFF_TABLES / 16); i += kCta) dst[i] = src[i];
  }
  if (!kGlobalWords) {
    const uint4 *src = reinterpret_cast<const uint4 *>(stream + sm.file.word0 + (size_t)first * S);
    const int nvec = (count * S + kGuardWords) / 4;
    for (int i = t; i < nvec; i += kCta) {
      const uint4 v = src[i];
      const uint32_t l = 4u * i, row = l / S, col = l % S, sw = row & (S - 1);
      uint32_t *r = sm.words + row * S;
      r[(col + 0) ^ sw] = __byte_perm(v.x, 0, 0x0123);
      r[(col + 1) ^ sw] = __byte_perm(v.y, 0, 0x0123);
      r[(col + 2) ^ sw] = __byte_perm(v.z, 0, 0x0123);
      r[(col + 3) ^ sw] = __byte_perm(v.w, 0, 0x0123);
    }
  }
  __syncthreads();
}

template <int S>
__global__ void __launch_bounds__(kCta, JGPU_HUFF_S >= 64 ? 2 : (JGPU_HUFF_SYNC_CTAS * 256 / kCta))
k_huff_sync(const jgpu_huff_file *__restrict__ files, const uint32_t *__restrict__ stream,
            const jgpu_huff_table *__restrict__ tables, const uint32_t *__restrict__ seg_first,
            uint32_t *__restrict__ state, uint32_t *__restrict__ nslots, uint32_t *__restrict__ segid,
            const uint32_t *__restrict__ carry_in, uint32_t *__restrict__ carry_out, int pass) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SyncSmem<S> &sm = *reinterpret_cast<SyncSmem<S> *>(smem_raw);
  const jgpu_huff_file &gf = files[blockIdx.y];
  const int bx = cta_x();
  const int own0 = bx * JGPU_HUFF_OWN;             /* first subsequence this CTA owns */
  if (own0 >= (int)gf.n_subseq) return;
  const int first = max(0, own0 - JGPU_HUFF_WARM);  /* first one it decodes */
  const int lead = own0 - first;                    /* warm-up threads: their results are dropped */
  const int count = min(lead + JGPU_HUFF_OWN, (int)gf.n_subseq - first);
  const int t = threadIdx.x;
  const uint32_t gi = gf.subseq0 + (uint32_t)first + (uint32_t)min(t, count - 1);
  const uint32_t cslot = gf.cta0 + (uint32_t)bx;
  uint32_t new0 = 0;
  if (pass > 0) {
    /* Did the state handed to this CTA change since the last launch?  If not, neither does
     * anything it computes. */
    const uint32_t g0 = gf.subseq0 + (uint32_t)own0;
    new0 = (nslots[g0] >> 31) ? 0u : carry_in[cslot];
    if (new0 == state[g0]) {
      if (t == 0) carry_out[cslot + 1] = carry_in[cslot + 1];
      return;
    }
  }
  stage<S>(sm, files, stream, tables, first, count);

  bool is_first;
  uint32_t s_in, n = 0;
  bool need;
  /* the thread whose start state is given, not derived from its left neighbour */
  const int fixed = pass == 0 ? 0 : lead;
  if (pass == 0) {
    /* restart interval of this subsequence: the last entry of seg_first[] not above it */
    const uint32_t i = (uint32_t)(first + min(t, count - 1));
    const uint32_t *sf = seg_first + gf.seg0;
    uint32_t lo = 0, hi = gf.n_seg; /* answer in [lo, hi) */
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (sf[mid] <= i) lo = mid; else hi = mid;
    }
    is_first = sf[lo] == i;
    if (t >= lead && t < count) segid[gi] = lo;
    s_in = 0;
    need = t < count;
  } else if (t >= lead) {
    const uint32_t v = nslots[gi];
    is_first = (v >> 31) != 0;
    n = v & 0x7fffffffu;
    s_in = t == lead ? new0 : state[gi];
    need = t == lead;
  } else {
    is_first = false;
    s_in = 0;
    need = false;
  }
  sm.s_out[t] = s_in;
  sm.s_in[t] = s_in;
  sm.n[t] = n;
  __syncthreads();
  /* what this CTA hands on stays what it was unless its last subsequence is redone */
  if (t == 0) {
    sm.s_out[count] = pass > 0 ? carry_in[cslot + 1] : 0u;
    sm.n_list = 0;
  }
  __syncthreads();

  const DevMem<S> mem = dev_mem<S>(sm, first, stream);
  const int bpm = sm.file.bpm;
  /* Rounds.  In the first two nearly every subsequence is decoded (from the guess, then from
   * what its left neighbour ended in); after that the ones whose input still moves thin out
   * quickly (42 %, 16 %, 6 %, ... of them on a 4K 4:2:0 file) but stay spread over all warps, so
   * from the third round on they are compacted: thread j takes the j-th subsequence of a list. */
  for (int round = 0;; round++) {
    int u = -1;   /* subsequence (CTA-local) this thread decodes in this round */
    if (round < 2) {
      if (need) u = t;
    } else if (t < sm.n_list) {
      u = sm.list[t];
    }
    if (u >= 0) {
      huff::NullSink sink;
      uint32_t err = 0, nn = 0;
      sm.s_out[u + 1] = huff::decode_subsequence(mem, bpm, (uint32_t)(first + u) * S, S, sm.s_in[u], sink, &nn, &err);
      sm.n[u] = nn;
    }
    __syncthreads();
    const uint32_t ni = (t <= fixed || is_first) ? s_in : sm.s_out[t];
    need = ni != s_in && t < count;
    s_in = ni;
    sm.s_in[t] = ni;
    if (t == 0) sm.n_list = 0;
    if (!__syncthreads_or(need ? 1 : 0)) break;
    if (round >= 1) {
      if (need) sm.list[atomicAdd(&sm.n_list, 1)] = (uint16_t)t;
      __syncthreads();
    }
  }
  if (t >= lead && t < count) {
    state[gi] = s_in;
    nslots[gi] = sm.n[t] | (is_first ? 0x80000000u : 0u);
    if (t == count - 1) carry_out[cslot + 1] = sm.s_out[count];
  }
}

/* Segmented exclusive scan of the slot counts: slots[i] = slots the restart interval has
 * advanced before subsequence i.  One CTA of 1024 threads per file; a warp takes 32 * kScanPer
 * consecutive subsequences as kScanPer stripes of 32 (coalesced loads and stores), scans stripe
 * after stripe with a running carry, and the warps' totals are combined once per sweep -- a 4K
 * file's 16 000 subsequences are one sweep.  (The first version looped over chunks of 1024 with
 * three barriers each: 37 us whatever the number of files; a thread-per-16-consecutive version
 * spent its time in uncoalesced accesses: 30 us.) */
constexpr int kScanPer = 16;

__global__ void __launch_bounds__(1024)
k_huff_scan(const jgpu_huff_file *__restrict__ files, const uint32_t *__restrict__ nslots,
            uint32_t *__restrict__ slots) {
  __shared__ uint32_t w_val[32], w_flag[32];
  __shared__ uint32_t s_carry;
  const uint32_t n_subseq = files[blockIdx.x].n_subseq, subseq0 = files[blockIdx.x].subseq0;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_subseq; base += 1024 * kScanPer) {
    const uint32_t w0 = base + (uint32_t)warp * (32 * kScanPer);
    uint32_t incl[kScanPer];   /* inclusive value since the last interval start, from the warp's first element */
    uint32_t own[kScanPer];    /* the element as stored (count | start flag << 31) */
    uint32_t seen = 0;         /* bit k: an interval starts at or before this element, inside the warp's range */
    uint32_t c_val = 0, c_flag = 0;   /* carry over the stripes: value at the end of the previous stripe */
#pragma unroll
    for (int k = 0; k < kScanPer; k++) {
      const uint32_t i = w0 + 32u * k + (uint32_t)lane;
      const uint32_t raw = i < n_subseq ? nslots[subseq0 + i] : 0u;
      uint32_t val = raw & 0x7fffffffu, flag = raw >> 31;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v2 = __shfl_up_sync(0xffffffffu, val, d), f2 = __shfl_up_sync(0xffffffffu, flag, d);
        if (lane >= d) {
          if (!flag) val += v2;
          flag |= f2;
        }
      }
      if (!flag) val += c_val;
      flag |= c_flag;
      own[k] = raw;
      incl[k] = val;
      seen |= flag << k;
      c_val = __shfl_sync(0xffffffffu, val, 31);
      c_flag = __shfl_sync(0xffffffffu, flag, 31);
    }
    if (lane == 31) {
      w_val[warp] = c_val;
      w_flag[warp] = c_flag;
    }
    __syncthreads();
    if (warp == 0) {
      uint32_t v = w_val[lane], g = w_flag[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v2 = __shfl_up_sync(0xffffffffu, v, d), g2 = __shfl_up_sync(0xffffffffu, g, d);
        if (lane >= d) {
          if (!g) v += v2;
          g |= g2;
        }
      }
      w_val[lane] = v;   /* inclusive over warps 0..lane */
      w_flag[lane] = g;
    }
    __syncthreads();
    /* what precedes this warp's range: the warps before it in the sweep, and the sweeps before */
    uint32_t pre = 0, pre_flag = 0;
    if (warp > 0) {
      pre = w_val[warp - 1];
      pre_flag = w_flag[warp - 1];
    }
    if (!pre_flag) pre += s_carry;
#pragma unroll
    for (int k = 0; k < kScanPer; k++) {
      const uint32_t i = w0 + 32u * k + (uint32_t)lane;
      const uint32_t v = incl[k] + (((seen >> k) & 1u) ? 0u : pre);
      if (i < n_subseq) slots[subseq0 + i] = (own[k] >> 31) ? 0u : v - (own[k] & 0x7fffffffu);
    }
    const uint32_t total = w_flag[31] ? w_val[31] : w_val[31] + s_carry;
    __syncthreads();
    if (t == 0) s_carry = total;
    __syncthreads();
  }
}

template <int S>
__global__ void __launch_bounds__(kCta, JGPU_HUFF_S >= 64 ? 2 : (1024 / kCta))
k_huff_write(const jgpu_huff_file *__restrict__ files, const uint32_t *__restrict__ stream,
             const jgpu_huff_table *__restrict__ tables, const uint32_t *__restrict__ seg_first,
             const uint32_t *__restrict__ state, const uint32_t *__restrict__ nslots,
             const uint32_t *__restrict__ slots, const uint32_t *__restrict__ segid,
             int16_t *__restrict__ coef, uint32_t *__restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SyncSmem<S> &sm = *reinterpret_cast<SyncSmem<S> *>(smem_raw);
  const jgpu_huff_file &gf = files[blockIdx.y];
  const int first = cta_x() * kCta;
  if (first >= (int)gf.n_subseq) return;
  const int count = min(kCta, (int)gf.n_subseq - first);
  const int t = threadIdx.x;
  stage<S>(sm, files, stream, tables, first, count);
  if (t >= count) return;
  const jgpu_huff_file &f = sm.file;
  const uint32_t i = (uint32_t)(first + t), gi = f.subseq0 + i;
  const uint32_t seg = segid[gi];
  const int seg_mcu0 = (int)seg * f.mcus_per_seg;
  const int64_t seg_blocks = (int64_t)min(f.mcus_per_seg, f.total_mcus - seg_mcu0) * f.bpm;
  const uint32_t slot0 = slots[gi], st = state[gi];
  const int64_t g0 = slot0 >> 6;
  uint32_t flags = 0;
  if ((slot0 & 63u) != JGPU_HUFF_STATE_Z(st) || (uint32_t)(g0 % f.bpm) != JGPU_HUFF_STATE_C(st)) {
    flags = JGPU_HUFF_ERR_SYNC;
  } else if (g0 < seg_blocks) {
    const DevMem<S> mem = dev_mem<S>(sm, first, stream);
    huff::StoreSink<DevMem<S>> sink;
    sink.start(&mem, f.bpm, f.nhmb, coef, seg_mcu0, g0, seg_blocks);
    uint32_t n = 0, err = 0;
    int pos = 0;
    const uint32_t out = huff::decode_subsequence(mem, f.bpm, i * S, S, st, sink, &n, &err, &pos);
    if (err) flags |= JGPU_HUFF_ERR_CODE;
    if (sink.g >= seg_blocks) {
      /* this thread decoded the interval's last block: whole bytes left over are the sequential
       * reader's to judge */
      const uint32_t *sf = seg_first + f.seg0;
      const uint32_t bits = sf[f.n_seg + 1 + seg];
      const long long used = (long long)(i - sf[seg]) * (32 * S) + pos;
      if (bits != 0xffffffffu && (long long)bits - used >= 8) flags |= JGPU_HUFF_ERR_TRAIL;
    }
    if (sink.g < seg_blocks) {
      /* the interval goes on: into the next subsequence, which must start where this one ended */
      if (i + 1 == seg_first[f.seg0 + seg + 1]) flags |= JGPU_HUFF_ERR_SHORT;
      else if (out != state[gi + 1] || n != (nslots[gi] & 0x7fffffffu)) flags |= JGPU_HUFF_ERR_SYNC;
    }
  }
  if (flags) atomicOr(status + f.status_slot, flags);
}

/* ---- write pass, staged -------------------------------------------------------------------
 * The write pass above stores every non-zero coefficient where it belongs the moment it is
 * decoded: 2-byte stores from 32 lanes into 32 different 128-byte lines, 28.6 L1 tag requests per
 * store instruction, and the table look-ups of the decoding loop queue behind them in the same
 * load/store unit (profiles/r1_ncu_summary.md: the kernel runs twice as fast without its stores).
 * Here every thread owns a 128-byte block buffer in shared memory, stores coefficients there, and
 * when a lane completes a block the WARP writes it out: lane r moves word r, one full line per
 * block.  Blocks a thread only sees a part of (the one its subsequence starts in the middle of,
 * the one it ends in the middle of) are written as their non-zero halves only, like before, so
 * two threads sharing a block never overwrite each other; the coefficient range is zero before
 * the kernel either way.  The scan words come straight from global memory through L1 (one load
 * per 32 bits consumed, fetched one refill ahead), which leaves the shared memory to the block
 * buffers: 47.6 KB per CTA, four CTAs per SM as before.
 *
 * Buffer layout: word w of lane l's block sits at word w ^ l of its 128 bytes, so that lanes
 * storing the same coefficient index hit different banks and the cooperative read of one block
 * (lane r reads physical word r) is conflict-free; the zig-zag table of this kernel holds byte
 * offsets (2 x natural index), so a coefficient's address is buffer + (offset ^ 4 * lane). */
template <int S>
struct WriteSmem {
  jgpu_huff_table tabs[JGPU_HUFF_TABLES];
  uint32_t blocks[kCta * 32];
  jgpu_huff_file file;
  unsigned char zz[64];
  uint4 brec[JGPU_HUFF_MAX_BLOCKS + 2];   /* block_records() */
};

__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
  asm volatile("{\n\t.reg .u16 t;\n\tcvt.u16.u32 t, %1;\n\tst.shared.u16 [%0], t;\n\t}" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32_v(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}

/* The warp writes out the blocks of the lanes in `m`; `pm`: lanes whose block is a partial one.
 * blk = the lane's current block, in blocks from the coefficient buffer's start; lane4 = 4 * lane.
 * Lane r moves physical word r of the block's buffer = word r ^ l of lane l's block. */
__device__ __forceinline__ void flush_blocks(uint32_t m, uint32_t pm, uint32_t blk, uint32_t warpbuf, uint32_t lane4,
                                             int16_t *__restrict__ coef) {
  unsigned char *const base = reinterpret_cast<unsigned char *>(coef);
  __syncwarp();
  uint32_t full = m & ~pm;
  while (full) {   /* whole blocks: one line each */
    const int l = __ffs((int)full) - 1;
    full &= full - 1;
    const uint32_t b = __shfl_sync(0xffffffffu, blk, l);
    const uint32_t a = warpbuf + (uint32_t)l * 128u + lane4;
    const uint32_t v = lds_u32_v(a);
    sts_u32(a, 0u);
    *reinterpret_cast<uint32_t *>(base + (size_t)b * 128 + (lane4 ^ ((uint32_t)l << 2))) = v;
  }
  uint32_t part = m & pm;
  while (part) {   /* blocks shared with a neighbouring subsequence: the non-zero halves only */
    const int l = __ffs((int)part) - 1;
    part &= part - 1;
    const uint32_t b = __shfl_sync(0xffffffffu, blk, l);
    const uint32_t a = warpbuf + (uint32_t)l * 128u + lane4;
    const uint32_t v = lds_u32_v(a);
    sts_u32(a, 0u);
    int16_t *dst = reinterpret_cast<int16_t *>(base + (size_t)b * 128 + (lane4 ^ ((uint32_t)l << 2)));
    if (v & 0xffffu) dst[0] = (int16_t)(v & 0xffffu);
    if (v >> 16) dst[1] = (int16_t)(v >> 16);
  }
  __syncwarp();
}

template <int S>
__global__ void __launch_bounds__(kCta, 1024 / kCta)
k_huff_write_staged(const jgpu_huff_file *__restrict__ files, const uint32_t *__restrict__ stream,
                    const jgpu_huff_table *__restrict__ tables, const uint32_t *__restrict__ seg_first,
                    const uint32_t *__restrict__ state, const uint32_t *__restrict__ nslots,
                    const uint32_t *__restrict__ slots, const uint32_t *__restrict__ segid,
                    int16_t *__restrict__ coef, uint32_t *__restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WriteSmem<S> &sm = *reinterpret_cast<WriteSmem<S> *>(smem_raw);
  const jgpu_huff_file &gf = files[blockIdx.y];
  const int first = cta_x() * kCta;
  if (first >= (int)gf.n_subseq) return;
  const int count = min(kCta, (int)gf.n_subseq - first);
  const int t = threadIdx.x, lane = t & 31;
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(files + blockIdx.y);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&sm.file);
    for (int i = t; i < (int)(sizeof(jgpu_huff_file) / 4); i += kCta) dst[i] = src[i];
    if (t < 64) sm.zz[t] = (unsigned char)(2 * c_zigzag[t]);   /* byte offset inside the block */
    uint4 *z = reinterpret_cast<uint4 *>(sm.blocks);
    for (int i = t; i < kCta * 8; i += kCta) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  block_records(sm.brec, sm.file, sm.tabs);
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(tables + sm.file.table0);
    uint4 *dst = reinterpret_cast<uint4 *>(sm.tabs);
    for (int i = t; i < (int)(sizeof(jgpu_huff_table) * JGPU_HUFF_TABLES / 16); i += kCta) dst[i] = src[i];
  }
  __syncthreads();
  const jgpu_huff_file &f = sm.file;
  DevMem<S> mem;
  mem.words = 0;
  mem.base_word = 0;
  mem.gwords = stream + sm.file.word0;
  mem.tabs = pinned((uint32_t)__cvta_generic_to_shared(sm.tabs));
  mem.file = pinned((uint32_t)__cvta_generic_to_shared(&sm.file));
  mem.zz = pinned((uint32_t)__cvta_generic_to_shared(sm.zz));
  mem.comps = pinned(huff::comp_pack(sm.file));
  mem.brec = pinned((uint32_t)__cvta_generic_to_shared(sm.brec));
  const uint32_t mybuf = pinned((uint32_t)__cvta_generic_to_shared(sm.blocks) + (uint32_t)t * 128u);
  const uint32_t warpbuf = mybuf - (uint32_t)lane * 128u;
  const uint32_t lane4 = pinned((uint32_t)lane * 4u);
  const int bpm = f.bpm, nhmb = f.nhmb;

  /* what the thread is about */
  bool active = false;
  uint32_t flags = 0;
  const uint32_t i = (uint32_t)(first + min(t, count - 1)), gi = f.subseq0 + i;
  const uint32_t seg = segid[gi];
  const int seg_mcu0 = (int)seg * f.mcus_per_seg;
  const int64_t seg_blocks = (int64_t)min(f.mcus_per_seg, f.total_mcus - seg_mcu0) * bpm;
  const uint32_t st = state[gi];
  int64_t g = 0;
  if (t < count) {
    const uint32_t slot0 = slots[gi];
    g = slot0 >> 6;
    if ((slot0 & 63u) != JGPU_HUFF_STATE_Z(st) || (uint32_t)(g % bpm) != JGPU_HUFF_STATE_C(st)) {
      flags = JGPU_HUFF_ERR_SYNC;
    } else if (g < seg_blocks) {
      active = true;
    }
  }
  const bool decoded = active;

  /* decoder state (jgpu_huff_core.h decode_subsequence, same arithmetic) */
  uint32_t c = JGPU_HUFF_STATE_C(st), z = JGPU_HUFF_STATE_Z(st);
  const uint32_t z0 = z;
  uint32_t nblk = 0, bad = 0;
  uint32_t bp = JGPU_HUFF_STATE_P(st) & 31u;
  int left = S - (int)(JGPU_HUFF_STATE_P(st) >> 5);
  uint32_t wa = 0, wb = 0, ahead = 0;
  const uint32_t *cur = mem.gwords + (i * S + (JGPU_HUFF_STATE_P(st) >> 5));
  int mbx = 0, mby = 0;
  uint32_t blk = 0;
  /* blocks of the restart interval from this thread's first one on (g counts up to it) */
  const uint32_t blocks_left0 = active ? (uint32_t)(seg_blocks - g) : 0u;
  uint32_t blocks_left = blocks_left0;
  if (active) {
    wa = __byte_perm(__ldg(cur), 0, 0x0123);
    wb = __byte_perm(__ldg(cur + 1), 0, 0x0123);
    ahead = __byte_perm(__ldg(cur + 2), 0, 0x0123);
    cur += 2;
    const int mcu = seg_mcu0 + (int)(g / bpm);
    mbx = mcu % nhmb;
    mby = mcu / nhmb;
    blk = (uint32_t)((mem.blk_base((int)c) + (int64_t)mbx * mem.blk_xs((int)c) + (int64_t)mby * mem.blk_ys((int)c)) >> 6);
  }
  uint32_t tdc = mem.table_ref(mem.blk_table(c)), tac = mem.table_ref(mem.blk_table(c) + 1);
  uint32_t tab = z ? tac : tdc;
  bool partial = z != 0;   /* the block this subsequence starts inside belongs to two threads */

  for (;;) {
    if (active) {
      const uint32_t look = huff::window32(wa, wb, bp);
      uint32_t e = mem.lut_at(tab, look >> (32 - JGPU_HUFF_LUT_BITS));
      if (e == 0) {
        const uint32_t l = huff::lookup_long(mem, mem.blk_table(c) + (z != 0), look >> 16);
        e = l ? JGPU_HUFF_ENTRY(l >> 8, l & 0xffu, z != 0) : JGPU_HUFF_ENTRY(16u, 0u, z != 0);
        bad |= (uint32_t)(l == 0);
      }
      const uint32_t total = JGPU_HUFF_ENTRY_T(e);
      const uint32_t za = z + JGPU_HUFF_ENTRY_A(e);
      if (za <= 64u) {
        /* T.81 F.2.2.1 EXTEND: the s bits after the code, less 2^s - 1 when their first bit is 0 */
        const uint32_t s = JGPU_HUFF_ENTRY_S(e);
        const uint32_t x = look << (total - s);
        const uint32_t bits = (x >> 1) >> (31u - s);
        const int v = (int)bits - (int)(~(uint32_t)((int)x >> 31) & ((1u << s) - 1u));
        if (v != 0) sts_u16(mybuf + (mem.zigzag((int)za - 1) ^ lane4), (uint32_t)v);
      }
      bp += total;
      if (bp >= 32u) {
        bp -= 32u;
        left--;
        wa = wb;
        wb = ahead;
        ahead = __byte_perm(__ldg(++cur), 0, 0x0123);
      }
      z = za;
      tab = tac;
    }
    const bool fin = z >= 64u;   /* (z is below 64 between symbols for a lane that has stopped, too) */
    const uint32_t m = __ballot_sync(0xffffffffu, fin);
    if (m) flush_blocks(m, __ballot_sync(0xffffffffu, partial), blk, warpbuf, lane4, coef);
    if (fin) {
      bad |= (uint32_t)(z > 64u && z <= JGPU_HUFF_EOB);
      z = 0;
      nblk++;
      partial = false;
      /* the next block: its tables and where it lies, one load (block_records) */
      const uint4 r = lds_u128(mem.brec + 16u * c);
      c = r.x >> 24;
      tdc = r.x & 0xffffffu;
      tac = tdc + (uint32_t)sizeof(jgpu_huff_table);
      tab = tdc;
      if (c == 0 && ++mbx == nhmb) {
        mbx = 0;
        mby++;
      }
      blk = r.y + (uint32_t)mbx * r.z + (uint32_t)mby * r.w;
      if (--blocks_left == 0) active = false;
    }
    if (left <= 0) active = false;
    if (!__any_sync(0xffffffffu, active)) break;
  }
  const int end = 32 * S;
  const int pos = 32 * (S - left) + (int)bp;
  const uint32_t n = 64u * nblk + z - z0;
  g += nblk;
  /* blocks left unfinished: the next subsequence carries on with them */
  {
    const uint32_t m = __ballot_sync(0xffffffffu, decoded && z != 0);
    if (m) flush_blocks(m, 0xffffffffu, blk, warpbuf, lane4, coef);
  }
  if (decoded) {
    const uint32_t out = JGPU_HUFF_STATE(pos > end ? pos - end : 0, c, z);
    if (bad) flags |= JGPU_HUFF_ERR_CODE;
    if (g >= seg_blocks) {
      const uint32_t *sf = seg_first + f.seg0;
      const uint32_t bits = sf[f.n_seg + 1 + seg];
      const long long used = (long long)(i - sf[seg]) * (32 * S) + pos;
      if (bits != 0xffffffffu && (long long)bits - used >= 8) flags |= JGPU_HUFF_ERR_TRAIL;
    } else {
      if (i + 1 == seg_first[f.seg0 + seg + 1]) flags |= JGPU_HUFF_ERR_SHORT;
      else if (out != state[gi + 1] || n != (nslots[gi] & 0x7fffffffu)) flags |= JGPU_HUFF_ERR_SYNC;
    }
  }
  if (flags) atomicOr(status + f.status_slot, flags);
}

/* DC differences -> DC values: the reference's `pred += diff` in a 16-bit accumulator
 * (src/xjpeg.c:430,479), reset at every restart interval (src/xjpeg.c:593-629).  A chain (one
 * component of one restart interval, in scan order) is cut into pieces of kDcPiece elements so
 * that a file without restart markers (one chain of 129 600 luma blocks at 4K) still spreads over
 * the GPU: k_huff_dc<0> leaves each piece's sum in `partial`, k_huff_dc<1> adds up the sums of the
 * pieces before its own and scans.  blockIdx.x = (interval * ncomps + component) * pieces + piece. */
constexpr int kDcPiece = 2048;

template <int APPLY>
__global__ void __launch_bounds__(256)
k_huff_dc(const jgpu_huff_file *__restrict__ files, int16_t *__restrict__ coef, int *__restrict__ partial,
          int pieces, size_t partial_stride) {
  __shared__ int w_sum[8];
  __shared__ int s_carry;
  const jgpu_huff_file &f = files[blockIdx.y];
  const uint32_t job = blockIdx.x / (uint32_t)pieces, piece = blockIdx.x % (uint32_t)pieces;
  const uint32_t seg = job / (uint32_t)f.ncomps;
  const int comp = (int)(job % (uint32_t)f.ncomps);
  if (seg >= f.n_seg || f.n_subseq == 0) return;
  const int seg_mcu0 = (int)seg * f.mcus_per_seg;
  const int cnt = min(f.mcus_per_seg, f.total_mcus - seg_mcu0) * f.hs[comp] * f.vs[comp];
  const int e0 = (int)piece * kDcPiece, e1 = min(cnt, e0 + kDcPiece);
  if (e0 >= cnt) return;
  int *mine = partial + partial_stride * blockIdx.y + (size_t)job * pieces;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) {
    int c = 0;
    if (APPLY) {
      for (uint32_t p = 0; p < piece; p++) c += mine[p];
    }
    s_carry = c;
  }
  __syncthreads();
  for (int base = e0; base < e1; base += 256) {
    const int e = base + t;
    int16_t *p = nullptr;
    int v = 0;
    if (e < e1) {
      p = coef + huff::dc_element_offset(f, comp, seg_mcu0, e);
      v = *p;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v2 = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += v2;
    }
    if (lane == 31) w_sum[warp] = v;
    __syncthreads();
    int pre = s_carry;
    for (int w = 0; w < warp; w++) pre += w_sum[w];
    v += pre;
    if (APPLY && p) *p = (int16_t)v;
    __syncthreads();
    if (t == 255) s_carry = v;
    __syncthreads();
  }
  if (!APPLY && t == 0) mine[piece] = s_carry;
}

template <int S>
cudaError_t configure_kernels() {
  cudaError_t e = cudaFuncSetAttribute(k_huff_sync<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(SyncSmem<S>));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_huff_write<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SyncSmem<S>));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_huff_write_staged<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)sizeof(WriteSmem<S>));
}

}  // namespace

/* JGPU_HUFF_WRITE=scatter: the write pass that stores coefficient by coefficient (A/B, profiles/r2_notes.md) */
static bool g_staged_write = true;

cudaError_t huff_configure() {
  const char *w = getenv("JGPU_HUFF_WRITE");
  g_staged_write = !(w && strcmp(w, "scatter") == 0);
  return configure_kernels<kHuffSubseqWords>();
}

int huff_launches(const HuffLaunch &l) { return l.sync_passes + 3 + (l.max_dc_chain > kDcPiece ? 1 : 0); }

size_t huff_dc_partial_ints(int max_dc_jobs, int max_dc_chain) {
  return (size_t)std::max(1, max_dc_jobs) * (size_t)std::max(1, (max_dc_chain + kDcPiece - 1) / kDcPiece);
}

/* Enqueues the whole entropy decode of a group of files.  The coefficient range the files
 * cover must have been zeroed on the same stream. */
int huff_launch(const HuffLaunch &l, cudaStream_t st) {
  constexpr int S = kHuffSubseqWords;
  if (l.n_files <= 0 || l.max_subseq <= 0) return 0;
  const dim3 grid((unsigned)((l.max_subseq + kCta - 1) / kCta), (unsigned)l.n_files);
  const dim3 sync_grid((unsigned)((l.max_subseq + JGPU_HUFF_OWN - 1) / JGPU_HUFF_OWN), (unsigned)l.n_files);
  const size_t smem = sizeof(SyncSmem<S>);
  cudaError_t e;
  if ((e = cudaMemsetAsync(l.d_carry[0] + l.carry0, 0, sizeof(uint32_t) * l.n_carry, st)) != cudaSuccess ||
      (e = cudaMemsetAsync(l.d_carry[1] + l.carry0, 0, sizeof(uint32_t) * l.n_carry, st)) != cudaSuccess ||
      (e = cudaMemsetAsync(l.d_status + l.status0, 0, sizeof(uint32_t) * l.n_files, st)) != cudaSuccess) {
    return jgpu_fail("entropy decoder: memset failed (%s)", cudaGetErrorString(e));
  }
  for (int pass = 0; pass < l.sync_passes; pass++) {
    k_huff_sync<S><<<sync_grid, kCta, smem, st>>>(l.d_files, l.d_stream, l.d_tables, l.d_seg_first, l.d_state,
                                             l.d_nslots, l.d_segid, l.d_carry[(pass + 1) & 1],
                                             l.d_carry[pass & 1], pass);
  }
  k_huff_scan<<<l.n_files, 1024, 0, st>>>(l.d_files, l.d_nslots, l.d_slots);
  if (g_staged_write) {
    k_huff_write_staged<S><<<grid, kCta, sizeof(WriteSmem<S>), st>>>(l.d_files, l.d_stream, l.d_tables, l.d_seg_first,
                                                                    l.d_state, l.d_nslots, l.d_slots, l.d_segid,
                                                                    l.d_coef, l.d_status);
  } else {
    k_huff_write<S><<<grid, kCta, smem, st>>>(l.d_files, l.d_stream, l.d_tables, l.d_seg_first, l.d_state,
                                              l.d_nslots, l.d_slots, l.d_segid, l.d_coef, l.d_status);
  }
  const int pieces = std::max(1, (l.max_dc_chain + kDcPiece - 1) / kDcPiece);
  const dim3 dc_grid((unsigned)(std::max(1, l.max_dc_jobs) * pieces), (unsigned)l.n_files);
  if (pieces > 1) {
    k_huff_dc<0><<<dc_grid, 256, 0, st>>>(l.d_files, l.d_coef, l.d_dc_partial, pieces, l.dc_partial_stride);
  }
  k_huff_dc<1><<<dc_grid, 256, 0, st>>>(l.d_files, l.d_coef, l.d_dc_partial, pieces, l.dc_partial_stride);
  e = cudaGetLastError();
  if (e != cudaSuccess) return jgpu_fail("entropy decoder: launch failed (%s)", cudaGetErrorString(e));
  return 0;
}

}  // namespace jgpu
