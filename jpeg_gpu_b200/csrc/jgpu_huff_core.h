/* jgpu_huff_core.h — entropy (Huffman) decoding of a baseline JPEG scan on the GPU: the data
 * structures and the per-thread decoding loop, shared by the CUDA kernels (jgpu_huff.cu), the
 * host-side preparation (jgpu_huff_prep.c) and a host emulation used by the CPU tests
 * (tests/host_core/huff_host.cpp), so the logic is checked without a GPU.
 *
 * What it replaces: the reference's sequential reader, src/xjpeg.c:449-632 (scan loop),
 * :163-205 (XJPEG_DECODE_HUFF), :311-336 (table build), producing the JPEG_DECODE_QUANT planes
 * of src/xjpeg.c:497-499,520-523,550-563 — SURVEY 8(f) rank 4.
 *
 * Parallelisation.  A Huffman stream has no random access, but decoders that start at a wrong
 * bit tend to fall into step with the true decoding after a few symbols (self-synchronisation;
 * Klein & Wiseman 2003, applied to GPUs by Weissenberger & Schmidt, ICPP 2018).  The
 * byte-unstuffed scan of a file is cut into SUBSEQUENCES of a fixed number of 32-bit words;
 * restart intervals (T.81 E.2.4) start on a subsequence boundary, where the decoder state is
 * known.  State at a subsequence boundary = (p, c, z): p = bits by which the symbol straddling
 * the boundary overshoots it, c = block of the MCU being decoded, z = zig-zag index of the next
 * coefficient (0: a DC symbol comes next).
 *   sync pass   thread i decodes subsequence i from its current estimate of the state at its
 *               start and hands the state it ends in to thread i+1; repeated (inside a CTA
 *               through shared memory, across CTAs by relaunching) until nothing changes.
 *               It also counts n_i, the coefficient slots the subsequence advances.
 *   scan        exclusive prefix sum of n_i inside each restart interval -> the absolute
 *               coefficient slot every subsequence starts at.
 *   write pass  thread i decodes once more, now storing each non-zero coefficient (DC as the
 *               difference it is coded as) at its place in the reference's plane layout, and
 *               VERIFIES that it ends in exactly the state subsequence i+1 starts from.  With
 *               the first state of every interval known, that check proves every state by
 *               induction, so a file either decodes exactly as a sequential reader would or is
 *               flagged and handed to the CPU reader.
 *   DC pass     per restart interval and component, a prefix sum over the DC differences in
 *               scan order (the predictor is a 16-bit accumulator, src/xjpeg.c:430,479).
 */
#ifndef JGPU_HUFF_CORE_H
#define JGPU_HUFF_CORE_H

#include <stdint.h>

#ifndef JGPU_HUFF_LUT_BITS
#define JGPU_HUFF_LUT_BITS 10
#endif
#define JGPU_HUFF_MAX_BLOCKS 10   /* blocks per MCU, T.81 B.2.3 */
#define JGPU_HUFF_TABLES 6        /* per file: (DC, AC) of each of up to 3 scan components */
#ifndef JGPU_HUFF_CTA
#define JGPU_HUFF_CTA 256         /* subsequences per CTA of the sync / write kernels (128 measured: profiles/r2_notes.md) */
#endif
/* Sync kernel: the first JGPU_HUFF_WARM threads of a CTA re-decode the last subsequences of the
 * CTA before it, only to hand the first subsequence the CTA OWNS a start state that has already
 * fallen into step (the chance that eight pieces in a row do not is about 0.1 % on 4:2:0 files).
 * Without them every CTA's first state is a blind guess and nearly every CTA has to be redone
 * in the second launch. */
#ifndef JGPU_HUFF_WARM
#define JGPU_HUFF_WARM 8
#endif
#define JGPU_HUFF_OWN (JGPU_HUFF_CTA - JGPU_HUFF_WARM)   /* subsequences a sync CTA owns */

/* What the decoding loop needs to know about a symbol, packed so that the loop spends no
 * arithmetic on taking the symbol apart (the loop is bound by the integer pipe, profiles/r2_notes.md 14):
 *   bits 0-4   T = code length + number of extra bits: what the symbol consumes
 *   bits 5-8   s = number of extra bits (T.81 F.1.2.1.1 SSSS)
 *   bits 9-15  A = how far the zig-zag index moves: 1 in a DC table; run + 1 in an AC table, and
 *              JGPU_HUFF_EOB for the end-of-block symbol 0x00, which lands every index (1..63)
 *              beyond anything a run can reach (63 + 16), so that
 *                  index + A == 64  the block's last coefficient,
 *                  index + A  > 64  the block ends without one (end of block, or a run past 63:
 *                                   an error, and below JGPU_HUFF_EOB + 1 by construction).
 * 0 is no entry (T >= 1 otherwise). */
#define JGPU_HUFF_EOB 96u
#define JGPU_HUFF_ENTRY(len, sym, ac)                                                             \
  ((uint32_t)((len) + ((sym) & 15u)) | (((uint32_t)(sym) & 15u) << 5) |                           \
   ((uint32_t)((ac) ? (((sym) & 0xffu) == 0 ? JGPU_HUFF_EOB : (((uint32_t)(sym) >> 4) & 15u) + 1u) : 1u) << 9))
#define JGPU_HUFF_ENTRY_T(e) ((e) & 31u)
#define JGPU_HUFF_ENTRY_S(e) (((e) >> 5) & 15u)
#define JGPU_HUFF_ENTRY_A(e) ((e) >> 9)

/* One Huffman table prepared for the decoder. */
typedef struct jgpu_huff_table {
  /* JGPU_HUFF_ENTRY of the symbol for every JGPU_HUFF_LUT_BITS-bit window whose leading bits are
   * a code of at most that many bits; 0 otherwise */
  uint16_t lut[1 << JGPU_HUFF_LUT_BITS];
  /* canonical decoding of the longer codes (T.81 F.2.2.3) on a 16-bit window W:
   * the code has length L for the smallest L with W < limit[L]; its symbol is
   * symbols[(W >> (16 - L)) + delta[L]] */
  uint32_t limit[18];
  int32_t delta[18];
  uint8_t symbols[256];
} jgpu_huff_table;

/* One file of a batch, as the kernels see it. */
typedef struct jgpu_huff_file {
  uint32_t word0;        /* first 32-bit word of its unstuffed scan in the stream buffer */
  uint32_t n_subseq;     /* subsequences of the scan */
  uint32_t subseq0;      /* index of its first subsequence in the per-subsequence arrays */
  uint32_t seg0;         /* index of its first entry in the segment table */
  uint32_t n_seg;        /* restart intervals (1 without DRI); the table holds n_seg+1 first
                            subsequences, then n_seg data lengths in bits (0xffffffff: not checked) */
  uint32_t cta0;         /* index of its first entry in the per-CTA carry arrays */
  int32_t mcus_per_seg;  /* restart interval in MCUs, or all MCUs */
  int32_t total_mcus;
  int32_t nhmb;          /* MCUs per MCU row */
  int32_t bpm;           /* blocks per MCU */
  int32_t ncomps;
  int32_t status_slot;   /* which entry of the status array the kernels flag */
  int32_t hs[3], vs[3];  /* sampling factors */
  int32_t hblocks[3];    /* blocks per block row of each plane */
  int64_t plane_off[3];  /* first int16 of each plane in the coefficient buffer */
  uint8_t blk_comp[JGPU_HUFF_MAX_BLOCKS + 2]; /* component of each block of the MCU */
  uint8_t blk_dx[JGPU_HUFF_MAX_BLOCKS + 2];   /* its position inside the MCU, in blocks */
  uint8_t blk_dy[JGPU_HUFF_MAX_BLOCKS + 2];
  uint32_t table0;       /* index of its first jgpu_huff_table */
  uint32_t reserved;
  /* address of block c of MCU (mbx, mby), in int16 from the coefficient buffer's start:
   * blk_base[c] + mbx * blk_xs[c] + mby * blk_ys[c]  (jgpu_huff_file_finish) */
  int64_t blk_base[JGPU_HUFF_MAX_BLOCKS + 2];
  int32_t blk_xs[JGPU_HUFF_MAX_BLOCKS + 2], blk_ys[JGPU_HUFF_MAX_BLOCKS + 2];
} jgpu_huff_file;

/* status bits the kernels raise per file */
#define JGPU_HUFF_ERR_CODE 1u      /* invalid code or coefficient index past 63 in the true decoding */
#define JGPU_HUFF_ERR_SYNC 2u      /* states did not settle within the sync passes */
#define JGPU_HUFF_ERR_SHORT 4u     /* the scan ends before the interval's last MCU */
#define JGPU_HUFF_ERR_TRAIL 8u     /* a restart interval has whole bytes left after its last MCU: the
                                      sequential reader resynchronises there (jgpu_front.c restart())
                                      and may reject the file; its call */

/* boundary state, packed */
#define JGPU_HUFF_STATE(p, c, z) ((uint32_t)(p) | ((uint32_t)(c) << 8) | ((uint32_t)(z) << 16))
#define JGPU_HUFF_STATE_P(s) ((s) & 0xffu)
#define JGPU_HUFF_STATE_C(s) (((s) >> 8) & 0xffu)
#define JGPU_HUFF_STATE_Z(s) (((s) >> 16) & 0xffu)

#ifdef __cplusplus

#if defined(__CUDACC__)
#define JGPU_HUFF_HD __host__ __device__ __forceinline__
#else
#define JGPU_HUFF_HD inline
#endif

namespace jgpu {
namespace huff {

/* The decoding loop reaches its data through a small accessor object, so that the kernels can
 * hand it 32-bit shared-memory addresses (explicit ld.shared, nothing for the compiler to
 * re-derive inside the loop) while the host emulation hands it plain pointers:
 *   word(i)            big-endian 32-bit word i of the file's unstuffed scan (16 guard bytes follow it)
 *   lut(t, i)          jgpu_huff_table[t].lut[i]          t = 2 * component + (AC ? 1 : 0)
 *   table_ref(t), lut_at(ref, i)   the same in two steps: the loop keeps the reference (on the
 *                      device the table's shared-memory address) of the table the next symbol
 *                      is read with, instead of rebuilding it for every symbol
 *   next_block(c, &tdc, &tac)      the block after block c of the MCU and the references of its
 *                      two tables (one shared-memory load on the device)
 *   cursor(i), next_word(cur)      a position in the scan that steps forward one word at a time
 *                      (on the device a pointer: no address is rebuilt from an index)
 *   limit(t, L), delta(t, L), symbol(t, i)                the canonical part of table t
 *   blk_table(c)       2 * component of block c of the MCU
 *   blk_base(c), blk_xs(c), blk_ys(c), zigzag(k)          write pass only */
/* Component of every block of the MCU, two bits each: lives in a register in the kernels. */
JGPU_HUFF_HD uint32_t comp_pack(const jgpu_huff_file &f) {
  uint32_t v = 0;
  for (int c = 0; c < f.bpm && c < JGPU_HUFF_MAX_BLOCKS; c++) v |= (uint32_t)(f.blk_comp[c] & 3u) << (2 * c);
  return v;
}

struct HostMem {
  const uint32_t *words;   /* the file's scan, as stored (big-endian bytes) */
  const jgpu_huff_table *tabs;
  const jgpu_huff_file *f;
  const unsigned char *zz;
  uint32_t word(uint32_t i) const { return __builtin_bswap32(words[i]); }
  uint32_t lut(uint32_t t, uint32_t i) const { return tabs[t].lut[i]; }
  uint32_t table_ref(uint32_t t) const { return t; }
  uint32_t lut_at(uint32_t ref, uint32_t i) const { return tabs[ref].lut[i]; }
  uint32_t next_block(uint32_t c, uint32_t *tdc, uint32_t *tac) const {
    c = c + 1 == (uint32_t)f->bpm ? 0 : c + 1;
    *tdc = blk_table(c);
    *tac = blk_table(c) + 1;
    return c;
  }
  typedef uint32_t Cursor;
  Cursor cursor(uint32_t i) const { return i; }
  uint32_t next_word(Cursor &cur) const { return word(++cur); }
  uint32_t limit(uint32_t t, int len) const { return tabs[t].limit[len]; }
  int32_t delta(uint32_t t, int len) const { return tabs[t].delta[len]; }
  uint32_t symbol(uint32_t t, int i) const { return tabs[t].symbols[i]; }
  uint32_t blk_table(uint32_t c) const { return 2u * ((comp_pack(*f) >> (2 * c)) & 3u); }
  int64_t blk_base(int c) const { return f->blk_base[c]; }
  int32_t blk_xs(int c) const { return f->blk_xs[c]; }
  int32_t blk_ys(int c) const { return f->blk_ys[c]; }
  uint32_t zigzag(int k) const { return zz[k]; }
};

/* Codes longer than the look-up table covers: canonical search by length on the 16-bit window
 * `look`; (length << 8) | symbol, or 0 for a bit pattern that is no code of the table. */
template <typename Mem>
JGPU_HUFF_HD uint32_t lookup_long(const Mem &mem, uint32_t t, uint32_t look) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int len = JGPU_HUFF_LUT_BITS + 1; len <= 16; len++) {
    if (look < mem.limit(t, len)) {
      return ((uint32_t)len << 8) | mem.symbol(t, (int)(look >> (16 - len)) + mem.delta(t, len));
    }
  }
  return 0;
}

/* JGPU_HUFF_ENTRY of the symbol at the head of the 16-bit window `look`, or 0 for a bit pattern
 * that is no code of table t (an AC table if `ac`). */
template <typename Mem>
JGPU_HUFF_HD uint32_t lookup(const Mem &mem, uint32_t t, uint32_t look, uint32_t ac) {
  const uint32_t e = mem.lut(t, look >> (16 - JGPU_HUFF_LUT_BITS));
  if (e) return e;
  const uint32_t l = lookup_long(mem, t, look);
  return l ? JGPU_HUFF_ENTRY(l >> 8, l & 0xffu, ac) : 0u;
}

/* Decodes the symbols that START inside one subsequence.
 *   mem           accessor (above); word() may be asked for up to four words past the
 *                 subsequence
 *   w0, nwords    the subsequence
 *   state         packed (p, c, z) at its start
 *   sink.coef(k, v)    a coefficient: zig-zag index k of the current block, value v (DC: the
 *                      coded difference); called for non-zero v only
 *   sink.block_done()  the current block is complete; returns true to stop (the restart
 *                      interval has all its blocks)
 * Returns the state at the end; *n_out = coefficient slots advanced; *pos_out = bits consumed
 * from the start of the subsequence when decoding stopped; *err is raised on a bit
 * pattern that is no code and on a run past coefficient 63 (the reference's reader does not
 * check either, src/xjpeg.c:67-78; ours fails on both, jgpu_front.c decode_symbol/decode_mcu). */
/* 32 bits of the stream starting `bp` bits (0..31) into word a, which word b follows */
JGPU_HUFF_HD uint32_t window32(uint32_t a, uint32_t b, uint32_t bp) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(b, a, bp);
#else
  return bp ? (a << bp) | (b >> (32u - bp)) : a;
#endif
}

template <typename Mem, typename Sink>
JGPU_HUFF_HD uint32_t decode_subsequence(const Mem &mem, int bpm, uint32_t w0, int nwords, uint32_t state,
                                         Sink &sink, uint32_t *n_out, uint32_t *err, int *pos_out = nullptr) {
  uint32_t c = JGPU_HUFF_STATE_C(state), z = JGPU_HUFF_STATE_Z(state);
  const uint32_t z0 = z;
  uint32_t nblk = 0, bad = 0;
  /* the window: 32 bits from bit bp of the current word; a symbol takes at most 16 + 15 of them.
   * `left`: words of the subsequence from the current one on. */
  uint32_t bp = JGPU_HUFF_STATE_P(state) & 31u;
  const uint32_t wi = w0 + (JGPU_HUFF_STATE_P(state) >> 5);
  int left = nwords - (int)(JGPU_HUFF_STATE_P(state) >> 5);
  uint32_t wa = mem.word(wi), wb = mem.word(wi + 1);
  typename Mem::Cursor cur = mem.cursor(wi + 2);
  uint32_t ahead = mem.word(wi + 2);   /* the word the next step forward needs, fetched one step early */
  /* the tables of the current block, and the one the next symbol is read with */
  uint32_t tdc = mem.table_ref(mem.blk_table(c)), tac = mem.table_ref(mem.blk_table(c) + 1);
  uint32_t tab = z ? tac : tdc;
  /* The body has no branch on the kind of symbol (DC / AC / end of block) except where a block
   * ends: the threads of a warp sit at unrelated places of their blocks, and every divergent
   * branch would be paid by all of them. */
  while (left > 0) {
    const uint32_t look = window32(wa, wb, bp);
    uint32_t e = mem.lut_at(tab, look >> (32 - JGPU_HUFF_LUT_BITS));
    if (e == 0) {   /* rare per thread: everything about long and invalid codes stays off the main path */
      const uint32_t l = lookup_long(mem, mem.blk_table(c) + (z != 0), look >> 16);
      /* no code: skip the window like jgpu_front.c decode_symbol; symbol 0 */
      e = l ? JGPU_HUFF_ENTRY(l >> 8, l & 0xffu, z != 0) : JGPU_HUFF_ENTRY(16u, 0u, z != 0);
      bad |= (uint32_t)(l == 0);
    }
    const uint32_t total = JGPU_HUFF_ENTRY_T(e);
    const uint32_t za = z + JGPU_HUFF_ENTRY_A(e);   /* index of the coefficient + 1, if there is one */
    if (za <= 64u) {
      /* T.81 F.2.2.1 EXTEND on the s bits after the code (s = 0 gives 0); never taken apart by a
       * sink that stores nothing */
      const uint32_t s = JGPU_HUFF_ENTRY_S(e);
      const uint32_t bits = ((look << (total - s)) >> 1) >> (31u - s);
      const uint32_t half = (1u << s) >> 1;
      const int v = bits < half ? (int)bits - (int)(1u << s) + 1 : (int)bits;
      if (v != 0) sink.coef((int)za - 1, v);
    }
    bp += total;
    if (bp >= 32u) {
      bp -= 32u;
      left--;
      wa = wb;
      wb = ahead;
      ahead = mem.next_word(cur);
    }
    z = za;
    tab = tac;
    if (za >= 64u) {   /* the block ends: last coefficient, end-of-block symbol, or a run past it */
      bad |= (uint32_t)(za > 64u && za <= JGPU_HUFF_EOB);   /* jgpu_front.c fails on the run */
      z = 0;
      nblk++;
      c = mem.next_block(c, &tdc, &tac);
      tab = tdc;
      if (sink.block_done()) break;
    }
  }
  *n_out = 64u * nblk + z - z0;
  if (bad) *err = 1;
  const int pos = 32 * (nwords - left) + (int)bp;   /* bits from the start of the subsequence to where decoding stopped */
  if (pos_out) *pos_out = pos;
  const int end = 32 * nwords;
  return JGPU_HUFF_STATE(pos > end ? pos - end : 0, c, z);
}

/* Sink of the sync pass: nothing is stored. */
struct NullSink {
  JGPU_HUFF_HD void coef(int, int) {}
  JGPU_HUFF_HD bool block_done() { return false; }
};

/* zig-zag position -> natural (row-major) position, T.81 figure A.6 (the users define their
 * own array from this initialiser: __constant__ on the device, plain on the host) */
#define JGPU_HUFF_ZIGZAG_NATURAL                                                                  \
  {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,        \
   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,        \
   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63}

/* Sink of the write pass: stores coefficients of the blocks [g, seg_blocks) of one restart
 * interval, g counted in scan order from the interval's first block.  Moving to the next block
 * is a few adds (no division): the threads of a warp finish blocks at unrelated moments, so
 * that step runs in almost every iteration of the warp. */
template <typename Mem>
struct StoreSink {
  const Mem *mem;
  int16_t *base;     /* the batch's coefficient buffer */
  int bpm, nhmb;
  int mbx, mby;      /* current MCU */
  int c;             /* current block inside it */
  int64_t g, seg_blocks;
  int16_t *blk;      /* current block */
  JGPU_HUFF_HD void locate() {
    blk = base + mem->blk_base(c) + (int64_t)mbx * mem->blk_xs(c) + (int64_t)mby * mem->blk_ys(c);
  }
  JGPU_HUFF_HD void start(const Mem *m, int blocks_per_mcu, int mcus_per_row, int16_t *coef_base, int seg_mcu0,
                          int64_t g0, int64_t nblocks) {
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
    locate();
  }
  JGPU_HUFF_HD void coef(int k, int v) {
#ifndef JGPU_HUFF_NO_STORE   /* experiment knob: the write pass without its stores (profiles/r1_ab_notes.md) */
    blk[mem->zigzag(k)] = (int16_t)v;
#else
    (void)k; (void)v;
#endif
  }
  JGPU_HUFF_HD bool block_done() {
    g++;
    if (g >= seg_blocks) return true;
    if (++c == bpm) {
      c = 0;
      if (++mbx == nhmb) {
        mbx = 0;
        mby++;
      }
    }
    locate();
    return false;
  }
};

/* Element e of the DC chain of component `comp` in the restart interval that starts at MCU
 * seg_mcu0: scan order = MCU by MCU, the component's blocks of an MCU row-major
 * (src/xjpeg.c:462-472). */
JGPU_HUFF_HD int64_t dc_element_offset(const jgpu_huff_file &f, int comp, int seg_mcu0, int e) {
  const int per = f.hs[comp] * f.vs[comp];
  const int mcu = seg_mcu0 + e / per, j = e % per;
  const int mbx = mcu % f.nhmb, mby = mcu / f.nhmb;
  const int bx = mbx * f.hs[comp] + j % f.hs[comp];
  const int by = mby * f.vs[comp] + j / f.hs[comp];
  return f.plane_off[comp] + ((int64_t)by * f.hblocks[comp] + bx) * 64;
}

}  // namespace huff
}  // namespace jgpu

#endif /* __cplusplus */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- host-side preparation (jgpu_huff_prep.c) -------------------------------------------- */

/* Builds one decoder table from the canonical description a DHT segment gives (T.81 C.2):
 * counts[L-1] codes of length L, their symbols in code order; `ac`: an AC table (the entries of
 * the look-up table differ, JGPU_HUFF_ENTRY).  Returns 0, or 1 when the counts over-subscribe the
 * code space. */
int jgpu_huff_build_table(jgpu_huff_table *t, const unsigned char counts[16], const unsigned char *symbols, int ac);

/* Derives blk_base / blk_xs / blk_ys once hs, vs, blk_*, hblocks and plane_off are set. */
void jgpu_huff_file_finish(jgpu_huff_file *f);

#ifdef __cplusplus
}
#endif
#endif
