// TEST-ONLY host emulation of the GPU entropy decoder (jgpu_huff.cu): the per-thread loop, the
// sinks and the address arithmetic are the product's own (jgpu_huff_core.h), the CTA-level
// choreography (rounds through "shared memory", carries between CTAs and launches, segmented
// scan, verification, DC pass) is restated here thread by thread, so the method and the shared
// code are checked against the sequential reader without a GPU.  Built by
// tests/test_huffman_host.py together with the product's C host sources; not part of the library.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "jgpu_front.h"
#include "jgpu_huff_core.h"
#include "jgpu_internal.h"

// jgpu_host.c can ask the CUDA runtime glue for page-locked memory; this host-only build has none
extern "C" void *jgpu_host_alloc(size_t) { return nullptr; }
extern "C" void jgpu_host_free(void *) {}
extern "C" int jgpu_host_is_pinned(const void *) { return 0; }

namespace {

const unsigned char kZigzag[64] = JGPU_HUFF_ZIGZAG_NATURAL;

// Write pass to emulate: 0 = k_huff_write (a store per coefficient), 1 = k_huff_write_staged (a block buffer per
// thread; whole blocks written out zeros included, blocks shared with a neighbouring subsequence as their non-zero
// coefficients only), 2 = the same with the threads run in descending order -- a block wrongly taken for a whole
// one would wipe what its other owner stored, in one of the two orders.
int g_write_mode = 0;

// The rule of k_huff_write_staged in terms of the decoding loop's sink.
struct StagedSink {
  const jgpu::huff::HostMem *mem;
  short *base;
  int bpm, nhmb, mbx, mby, c;
  int64_t g, seg_blocks;
  short buf[64];
  bool partial;
  short *locate() const {
    return base + mem->blk_base(c) + (int64_t)mbx * mem->blk_xs(c) + (int64_t)mby * mem->blk_ys(c);
  }
  void start(const jgpu::huff::HostMem *m, int blocks_per_mcu, int mcus_per_row, short *coef_base, int seg_mcu0,
             int64_t g0, int64_t nblocks, bool starts_inside_a_block) {
    mem = m;
    bpm = blocks_per_mcu;
    nhmb = mcus_per_row;
    base = coef_base;
    g = g0;
    seg_blocks = nblocks;
    const int mcu = seg_mcu0 + (int)(g0 / bpm);
    c = (int)(g0 % bpm);
    mbx = mcu % nhmb;
    mby = mcu / nhmb;
    memset(buf, 0, sizeof(buf));
    partial = starts_inside_a_block;
  }
  void coef(int k, int v) { buf[mem->zigzag(k)] = (short)v; }
  void flush(bool part) {
    short *blk = locate();
    if (!part) {
      memcpy(blk, buf, sizeof(buf));
    } else {
      for (int i = 0; i < 64; i++) {
        if (buf[i]) blk[i] = buf[i];
      }
    }
    memset(buf, 0, sizeof(buf));
  }
  bool block_done() {
    flush(partial);
    partial = false;
    g++;
    if (g >= seg_blocks) return true;
    if (++c == bpm) {
      c = 0;
      if (++mbx == nhmb) {
        mbx = 0;
        mby++;
      }
    }
    return false;
  }
};


}  // namespace

extern "C" void huff_set_write_mode(int mode) { g_write_mode = mode; }

// JGPU_HUFF_ENTRY and its fields, for the test of the invariants the decoding loop rests on
extern "C" unsigned huff_entry(unsigned len, unsigned sym, int ac) { return JGPU_HUFF_ENTRY(len, sym, ac); }
extern "C" unsigned huff_entry_field(unsigned e, int which) {
  return which == 0 ? JGPU_HUFF_ENTRY_T(e) : which == 1 ? JGPU_HUFF_ENTRY_S(e) : JGPU_HUFF_ENTRY_A(e);
}

// k_huff_scan's arithmetic, lane by lane: `warps` warps of 32 lanes, each warp scanning `per` stripes of 32
// consecutive counts with a running carry, the warps' totals combined once per sweep, a carry from sweep to sweep
// (the kernel: 32 warps, 16 stripes).  Bit 31 of a count starts a restart interval.
static void scan_striped(const std::vector<uint32_t> &nslots, std::vector<uint32_t> &slots, int warps, int per) {
  const uint32_t n = (uint32_t)nslots.size();
  uint32_t s_carry = 0;
  for (uint32_t base = 0; base < n; base += (uint32_t)(warps * 32 * per)) {
    std::vector<uint32_t> w_val(warps), w_flag(warps);
    std::vector<std::vector<uint32_t>> incl(warps, std::vector<uint32_t>(32 * per)), own(incl), seen(incl);
    for (int w = 0; w < warps; w++) {
      uint32_t c_val = 0, c_flag = 0;
      for (int k = 0; k < per; k++) {
        uint32_t val[32], flag[32];
        for (int l = 0; l < 32; l++) {
          const uint32_t i = base + (uint32_t)(w * 32 * per + 32 * k + l);
          const uint32_t raw = i < n ? nslots[i] : 0u;
          own[w][32 * k + l] = raw;
          val[l] = raw & 0x7fffffffu;
          flag[l] = raw >> 31;
        }
        for (int d = 1; d < 32; d <<= 1) {   // the shuffle steps: every lane reads the values of the step before
          uint32_t v2[32], f2[32];
          for (int l = 0; l < 32; l++) {
            v2[l] = l >= d ? val[l - d] : 0;
            f2[l] = l >= d ? flag[l - d] : 0;
          }
          for (int l = d; l < 32; l++) {
            if (!flag[l]) val[l] += v2[l];
            flag[l] |= f2[l];
          }
        }
        for (int l = 0; l < 32; l++) {
          if (!flag[l]) val[l] += c_val;
          flag[l] |= c_flag;
          incl[w][32 * k + l] = val[l];
          seen[w][32 * k + l] = flag[l];
        }
        c_val = val[31];
        c_flag = flag[31];
      }
      w_val[w] = c_val;
      w_flag[w] = c_flag;
    }
    for (int w = 1; w < warps; w++) {   // inclusive over the warps (the kernel: one warp, shuffles)
      if (!w_flag[w]) w_val[w] += w_val[w - 1];
      w_flag[w] |= w_flag[w - 1];
    }
    for (int w = 0; w < warps; w++) {
      uint32_t pre = 0, pre_flag = 0;
      if (w > 0) {
        pre = w_val[w - 1];
        pre_flag = w_flag[w - 1];
      }
      if (!pre_flag) pre += s_carry;
      for (int e = 0; e < 32 * per; e++) {
        const uint32_t i = base + (uint32_t)(w * 32 * per + e);
        const uint32_t v = incl[w][e] + (seen[w][e] ? 0u : pre);
        if (i < n) slots[i] = (own[w][e] >> 31) ? 0u : v - (own[w][e] & 0x7fffffffu);
      }
    }
    s_carry = w_flag[warps - 1] ? w_val[warps - 1] : w_val[warps - 1] + s_carry;
  }
}

// Decodes one JPEG file's scan into QUANT planes (reference layout) the way the kernels do.
//   subseq_words, cta   the kernel's constants (32, 256), shrinkable so that small files still
//                       exercise hand-overs between subsequences, CTAs and launches
//   sync_passes         launches of the sync kernel
//   stats[0] = subsequences, [1] = restart intervals, [2] = rounds of the busiest CTA,
//   [3] = CTAs that recomputed in a later pass
// Returns the coefficient count (int16 elements), -1 when the file is not eligible, and leaves
// the kernels' status word in *status.
extern "C" long long huff_emulate(const unsigned char *jpeg, int size, short *coef, long long coef_cap,
                                  int subseq_words, int cta, int sync_passes, unsigned *status, int *stats) {
  using namespace jgpu::huff;
  const jpeg_decode_ctx_vtbl &v = JFRONT_DECODE_CTX_VTBL;
  jpeg_info info;
  memset(&info, 0, sizeof(info));
  info.buf = const_cast<unsigned char *>(jpeg);
  info.size = size;
  jpeg_decode_ctx *front = v.decode_alloc(&info);
  jpeg_header header;
  jgpu_image_desc desc;
  jgpu_layout lay;
  if (!front) return -1;
  if (v.decode_header(front, &header) != EXIT_SUCCESS || jgpu_desc_from_header(&header, &desc) ||
      jgpu_layout_query(&desc, &lay) || lay.coef_len > coef_cap) {
    v.decode_free(front);
    return -1;
  }
  int nseg_bound = 0;
  const long long cap = jfront_huff_bound(front, subseq_words, &nseg_bound);
  std::vector<unsigned char> stream((size_t)cap);
  std::vector<unsigned int> seg_first(2 * (size_t)nseg_bound + 3);
  std::vector<jgpu_huff_table> tabs(JGPU_HUFF_TABLES);
  jgpu_huff_file f;
  memset(&f, 0, sizeof(f));
  const char *why = nullptr;
  const long long bytes = jfront_huff_prepare(front, subseq_words, stream.data(), cap, seg_first.data(),
                                              2 * nseg_bound + 3, tabs.data(), &f, &why);
  v.decode_free(front);
  if (bytes < 0) return -1;
  for (int p = 0; p < desc.ncomps; p++) {
    f.hblocks[p] = lay.plane[p].hblocks;
    f.plane_off[p] = lay.plane[p].coef_off;
  }
  jgpu_huff_file_finish(&f);
  const HostMem mem = {reinterpret_cast<const uint32_t *>(stream.data()), tabs.data(), &f, kZigzag};
  const int S = subseq_words;
  const int n = (int)f.n_subseq;
  // the sync kernel's CTAs overlap: `warm` threads re-decode the end of the previous CTA's range
  // (JGPU_HUFF_WARM = 8 of 256 in the product; scaled down with the shrunken geometries)
  const int warm = cta >= 32 ? JGPU_HUFF_WARM * cta / JGPU_HUFF_CTA : cta / 4;
  const int own = cta - warm;
  const int nctas = (n + own - 1) / own;
  std::vector<uint32_t> state(n, 0), nslots(n, 0), slots(n, 0), segid(n, 0);
  std::vector<uint32_t> carry[2] = {std::vector<uint32_t>(nctas + 1, 0), std::vector<uint32_t>(nctas + 1, 0)};
  unsigned st_flags = 0;
  int max_rounds = 0, recomputed = 0;

  // ---- k_huff_sync, `sync_passes` launches ----
  for (int pass = 0; pass < sync_passes; pass++) {
    const std::vector<uint32_t> &cin = carry[(pass + 1) & 1];
    std::vector<uint32_t> &cout = carry[pass & 1];
    for (int x = 0; x < nctas; x++) {
      const int own0 = x * own;
      const int first = std::max(0, own0 - warm), lead = own0 - first;
      const int count = std::min(lead + own, n - first);
      uint32_t new0 = 0;
      if (pass > 0) {
        new0 = (nslots[own0] >> 31) ? 0u : cin[x];
        if (new0 == state[own0]) {
          cout[x + 1] = cin[x + 1];
          continue;
        }
        recomputed++;
      }
      const int fixed = pass == 0 ? 0 : lead;
      std::vector<uint32_t> s_in(count), s_out(count + 1, 0), nn(count, 0);
      std::vector<char> need(count), is_first(count);
      for (int t = 0; t < count; t++) {
        const uint32_t i = (uint32_t)(first + t);
        if (pass == 0) {
          uint32_t lo = 0, hi = f.n_seg;
          while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (seg_first[mid] <= i) lo = mid; else hi = mid;
          }
          is_first[t] = seg_first[lo] == i;
          if (t >= lead) segid[i] = lo;
          s_in[t] = 0;
          need[t] = 1;
        } else if (t >= lead) {
          is_first[t] = (nslots[i] >> 31) != 0;
          nn[t] = nslots[i] & 0x7fffffffu;
          s_in[t] = t == lead ? new0 : state[i];
          need[t] = t == lead;
        } else {
          is_first[t] = 0;
          s_in[t] = 0;
          need[t] = 0;
        }
        s_out[t] = s_in[t];
      }
      s_out[count] = pass > 0 ? cin[x + 1] : 0u;
      int rounds = 0;
      for (;;) {
        rounds++;
        if (getenv("HUFF_EMU_TRACE") && pass == 0) {
          int act = 0, warps = 0;
          for (int t = 0; t < count; t += 32) {
            int a = 0;
            for (int u = t; u < std::min(count, t + 32); u++) a += need[u] != 0;
            act += a;
            warps += a > 0;
          }
          fprintf(stderr, "cta %d round %d active %d warps %d\n", x, rounds, act, warps);
        }
        for (int t = 0; t < count; t++) {
          if (!need[t]) continue;
          NullSink sink;
          uint32_t err = 0;
          s_out[t + 1] = decode_subsequence(mem, f.bpm, (uint32_t)(first + t) * S, S, s_in[t], sink, &nn[t], &err);
        }
        bool any = false;
        for (int t = 0; t < count; t++) {
          const uint32_t ni = (t <= fixed || is_first[t]) ? s_in[t] : s_out[t];
          need[t] = ni != s_in[t];
          s_in[t] = ni;
          any |= need[t] != 0;
        }
        if (!any) break;
      }
      max_rounds = std::max(max_rounds, rounds);
      for (int t = lead; t < count; t++) {
        state[first + t] = s_in[t];
        nslots[first + t] = nn[t] | (is_first[t] ? 0x80000000u : 0u);
      }
      cout[x + 1] = s_out[count];
    }
  }
  // ---- k_huff_scan ----
  {
    uint32_t acc = 0;
    for (int i = 0; i < n; i++) {
      if (nslots[i] >> 31) acc = 0;
      slots[i] = acc;
      acc += nslots[i] & 0x7fffffffu;
    }
    // the kernel's striped arithmetic must give the same, in its own geometry and in one that makes small files
    // cross warps and sweeps
    const int geometries[2][2] = {{32, 16}, {2, 2}};
    for (const auto &gm : geometries) {
      std::vector<uint32_t> in(nslots.begin(), nslots.begin() + n), out(n, 0xdeadbeefu);
      scan_striped(in, out, gm[0], gm[1]);
      for (int i = 0; i < n; i++) {
        if (out[i] != slots[i]) st_flags |= 0x80000000u;   // not a kernel status: fails every test that checks status == 0
      }
    }
  }
  // ---- k_huff_write ----
  memset(coef, 0, sizeof(short) * (size_t)lay.coef_len);
  for (int step = 0; step < n; step++) {
    const int i = g_write_mode == 2 ? n - 1 - step : step;
    const uint32_t seg = segid[i];
    const int seg_mcu0 = (int)seg * f.mcus_per_seg;
    const int64_t seg_blocks = (int64_t)std::min(f.mcus_per_seg, f.total_mcus - seg_mcu0) * f.bpm;
    const uint32_t slot0 = slots[i], st = state[i];
    const int64_t g0 = slot0 >> 6;
    if ((slot0 & 63u) != JGPU_HUFF_STATE_Z(st) || (uint32_t)(g0 % f.bpm) != JGPU_HUFF_STATE_C(st)) {
      st_flags |= JGPU_HUFF_ERR_SYNC;
    } else if (g0 < seg_blocks) {
      StoreSink<HostMem> sink;
      StagedSink staged;
      uint32_t nn = 0, err = 0, out;
      int pos = 0;
      int64_t g_end;
      if (g_write_mode == 0) {
        sink.start(&mem, f.bpm, f.nhmb, coef, seg_mcu0, g0, seg_blocks);
        out = decode_subsequence(mem, f.bpm, (uint32_t)i * S, S, st, sink, &nn, &err, &pos);
        g_end = sink.g;
      } else {
        staged.start(&mem, f.bpm, f.nhmb, coef, seg_mcu0, g0, seg_blocks, JGPU_HUFF_STATE_Z(st) != 0);
        out = decode_subsequence(mem, f.bpm, (uint32_t)i * S, S, st, staged, &nn, &err, &pos);
        if (JGPU_HUFF_STATE_Z(out) != 0) staged.flush(true);   // the block the next subsequence carries on with
        g_end = staged.g;
      }
      if (err) st_flags |= JGPU_HUFF_ERR_CODE;
      if (g_end >= seg_blocks) {
        const uint32_t bits = seg_first[f.n_seg + 1 + seg];
        const long long used = (long long)((uint32_t)i - seg_first[seg]) * (32 * S) + pos;
        if (bits != 0xffffffffu && (long long)bits - used >= 8) st_flags |= JGPU_HUFF_ERR_TRAIL;
      }
      if (g_end < seg_blocks) {
        if ((uint32_t)i + 1 == seg_first[seg + 1]) st_flags |= JGPU_HUFF_ERR_SHORT;
        else if (out != state[i + 1] || nn != (nslots[i] & 0x7fffffffu)) st_flags |= JGPU_HUFF_ERR_SYNC;
      }
    }
  }
  // ---- k_huff_dc ----
  for (uint32_t seg = 0; seg < f.n_seg; seg++) {
    const int seg_mcu0 = (int)seg * f.mcus_per_seg;
    for (int comp = 0; comp < f.ncomps; comp++) {
      const int cnt = std::min(f.mcus_per_seg, f.total_mcus - seg_mcu0) * f.hs[comp] * f.vs[comp];
      int acc = 0;
      for (int e = 0; e < cnt; e++) {
        short *p = coef + dc_element_offset(f, comp, seg_mcu0, e);
        acc += *p;
        *p = (short)acc;
      }
    }
  }
  *status = st_flags;
  if (stats) {
    stats[0] = n;
    stats[1] = (int)f.n_seg;
    stats[2] = max_rounds;
    stats[3] = recomputed;
  }
  return lay.coef_len;
}

// jgpu_huff_build_table against a bit-by-bit canonical decoder: every 16-bit window.
extern "C" int huff_table_selfcheck(const unsigned char *counts, const unsigned char *symbols) {
  int mismatches = 0;
  for (int ac = 0; ac < 2; ac++) {
  jgpu_huff_table t;
  if (jgpu_huff_build_table(&t, counts, symbols, ac)) return -1;
  for (uint32_t look = 0; look < 65536; look++) {
    // T.81 F.2.2.3 DECODE
    int code = 0, k = 0, want = 0, first = 0;
    for (int len = 1; len <= 16; len++) {
      code = (int)(look >> (16 - len));
      const int cnt = counts[len - 1];
      if (code - first < cnt) {
        want = (int)JGPU_HUFF_ENTRY((unsigned)len, (unsigned)symbols[k + code - first], ac);
        break;
      }
      k += cnt;
      first = (first + cnt) << 1;
    }
    const jgpu::huff::HostMem mem = {nullptr, &t, nullptr, nullptr};
    if ((int)jgpu::huff::lookup(mem, 0, look, (uint32_t)ac) != want) mismatches++;
  }
  }
  return mismatches;
}
