/* jgpu_huff.h — internal: launch interface of the GPU entropy decoder (jgpu_huff.cu). */
#ifndef JGPU_HUFF_H
#define JGPU_HUFF_H

#include <cuda_runtime.h>
#include "jgpu_huff_core.h"

namespace jgpu {

#ifndef JGPU_HUFF_S
#define JGPU_HUFF_S 32
#endif
constexpr int kHuffSubseqWords = JGPU_HUFF_S;   /* 1024 bits per subsequence (16 and 64 measured, profiles/r1_ab_notes.md) */
/* JGPU_HUFF_GW=1: the sync kernel reads the scan words from global memory through L1 instead of
 * staging them in shared memory: 19 KB instead of 51 KB per CTA, so more CTAs are resident while
 * the ones in their late rounds (one or two warps busy, the rest at the barrier) hold their place. */
#ifndef JGPU_HUFF_GW
#define JGPU_HUFF_GW 1
#endif
#ifndef JGPU_HUFF_SYNC_CTAS
#define JGPU_HUFF_SYNC_CTAS (JGPU_HUFF_GW ? 6 : 4)
#endif
constexpr int kHuffSyncCtasPerSm = JGPU_HUFF_SYNC_CTAS;   /* the runtime sizes its groups of files by this */
constexpr int kHuffSyncPasses = 3;     /* launches of k_huff_sync (1 + hand-overs across CTAs) */

/* One group of files, everything on the device.  Per-subsequence arrays are indexed by
 * jgpu_huff_file.subseq0 + i, the carry arrays by cta0 + CTA. */
struct HuffLaunch {
  int n_files = 0;
  int max_subseq = 0;      /* largest n_subseq of the group */
  int max_dc_jobs = 0;     /* largest n_seg * ncomps of the group */
  int max_dc_chain = 0;    /* longest DC chain: blocks of one component in one restart interval */
  int *d_dc_partial = nullptr;      /* n_files x dc_partial_stride ints of scratch */
  size_t dc_partial_stride = 0;     /* >= huff_dc_partial_ints(max_dc_jobs, max_dc_chain) */
  int sync_passes = kHuffSyncPasses;
  size_t carry0 = 0, n_carry = 0;   /* the group's part of each carry array */
  int status0 = 0;         /* first status word of the group (n_files words are cleared) */
  const jgpu_huff_file *d_files = nullptr;
  const uint32_t *d_stream = nullptr;
  const jgpu_huff_table *d_tables = nullptr;
  const uint32_t *d_seg_first = nullptr;
  uint32_t *d_state = nullptr, *d_nslots = nullptr, *d_slots = nullptr, *d_segid = nullptr;
  uint32_t *d_carry[2] = {nullptr, nullptr};
  uint32_t *d_status = nullptr;   /* indexed by jgpu_huff_file.status_slot */
  int16_t *d_coef = nullptr;      /* base the files' plane_off are relative to */
};

cudaError_t huff_configure();
int huff_launches(const HuffLaunch &l);
size_t huff_dc_partial_ints(int max_dc_jobs, int max_dc_chain);
/* Returns 0 or 1 (jgpu_fail).  The coefficient range of the files must be zero. */
int huff_launch(const HuffLaunch &l, cudaStream_t stream);

}  // namespace jgpu
#endif
