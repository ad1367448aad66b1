/* jgpu_front.c — CPU entropy front end: baseline sequential JPEG (SOF0) ->
 * quantised coefficient planes.
 *
 * This is the producer half the GPU back end needs when the library is used
 * outside the reference tree (inside it, cuda_decode_set_frontend() plugs in
 * the reference's own XJPEG_DECODE_CTX_VTBL).  It is written from ITU-T T.81
 * (marker syntax B.2, Huffman table generation C.2 / F.2.2.3, EXTEND F.2.2.1,
 * restart intervals F.1.1.5/E.2.4) and produces, buffer for buffer, what the
 * reference's reader produces for the same file:
 *   JPEG_DECODE_QUANT  de-zigzagged quantised blocks        src/xjpeg.c:497-499,520-523
 *   JPEG_DECODE_DCT    same, multiplied by the DQT entry    src/xjpeg.c:501-503,524-527
 *   JPEG_DECODE_PACK   run/level words + per-block index    src/xjpeg.c:484-496,513-519,531-535
 *   block placement inside image.coef                       src/xjpeg.c:550-563
 *   header fields                                           src/jpeg_wrap.c:263-319
 * Like the xjpeg backend it has no YUV/RGB output of its own
 * (src/jpeg_wrap.c:335-339): those are the GPU's job.
 *
 * Unlike the reference (whose bitstream validation is compiled out by default,
 * src/xjpeg.c:67-78) every read is bounds-checked and malformed or unsupported
 * files (progressive, 12-bit, multi-scan, 4 components) fail with
 * EXIT_FAILURE and a one-line message instead of crashing.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "jgpu_internal.h"
#include "jgpu_front.h"
#include "jgpu_huff_core.h"

/* zig-zag position -> natural (row-major) position, T.81 figure A.6 */
static const unsigned char kNatural[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

#define FAST_BITS 9

typedef struct huff_table {
  int valid;
  unsigned char counts[16]; /* codes of each length, as the DHT segment lists them */
  unsigned char symbols[256];
  /* fast[peek] = (length << 8) | symbol for codes of <= FAST_BITS bits, else 0 */
  unsigned short fast[1 << FAST_BITS];
  /* canonical decoding, T.81 F.2.2.3: for code length L (1..16) */
  int maxcode[18]; /* largest code of length L, -1 if none; [17] = sentinel */
  int valptr[17];  /* index of first symbol of length L */
  int mincode[17];
} huff_table;

typedef struct front_comp {
  int id, hsamp, vsamp, tq;
} front_comp;

typedef struct jfront_ctx {
  const unsigned char *buf;
  int size;
  int pos;          /* next unread byte while parsing markers */
  int sos_pos;      /* offset of the SOS segment length, set by decode_header */
  int have_frame;
  const char *error;
  /* tables */
  jpeg_quant quant[NQUANT_MAX];
  huff_table dc[4], ac[4];
  int restart_interval;
  /* frame */
  int bits, width, height, ncomps, hmax, vmax, nhmb, nvmb;
  front_comp comp[NCOMPS_MAX];
  /* entropy-coded segment reader */
  unsigned long long acc; /* bit accumulator, MSB first */
  int nbits;              /* valid bits in acc */
  int ecs_pos;            /* next byte of entropy-coded data */
  int hit_marker;         /* a marker (not 0xFF00) was met; zeros are fed */
} jfront_ctx;

/* ---- marker-level helpers ------------------------------------------------- */

static int fail(jfront_ctx *c, const char *msg) {
  if (!c->error) c->error = msg;
  return 1;
}

static int get_u8(jfront_ctx *c, int *v) {
  if (c->pos >= c->size) return fail(c, "Error reading past the end of file.");
  *v = c->buf[c->pos++];
  return 0;
}

static int get_u16(jfront_ctx *c, int *v) {
  if (c->pos + 2 > c->size) return fail(c, "Error reading past the end of file.");
  *v = (c->buf[c->pos] << 8) | c->buf[c->pos + 1];
  c->pos += 2;
  return 0;
}

static int parse_dqt(jfront_ctx *c) {
  int len, end;
  if (get_u16(c, &len)) return 1;
  end = c->pos + len - 2;
  if (len < 2 || end > c->size) return fail(c, "Error decoding DQT, segment overruns file.");
  while (c->pos < end) {
    int pq_tq, pq, tq, i, v;
    jpeg_quant *q;
    if (get_u8(c, &pq_tq)) return 1;
    pq = pq_tq >> 4;
    tq = pq_tq & 15;
    if (pq > 1) return fail(c, "Error DQT expected Pq value 0 or 1.");
    if (tq > 3) return fail(c, "Error DQT expected Tq value 0 to 3.");
    if (c->pos + 64 * (pq + 1) > end) return fail(c, "Error decoding DQT, unprocessed bytes.");
    q = &c->quant[tq];
    q->valid = 1;
    q->bits = pq ? 16 : 8;
    for (i = 0; i < 64; i++) {
      if (pq ? get_u16(c, &v) : get_u8(c, &v)) return 1;
      q->tbl[kNatural[i]] = (unsigned short)v; /* stored de-zigzagged */
    }
  }
  return 0;
}

static int build_huffman(huff_table *h, const unsigned char counts[16]) {
  int code = 0, k = 0, len, i;
  memset(h->fast, 0, sizeof(h->fast));
  memcpy(h->counts, counts, 16);
  for (len = 1; len <= 16; len++) {
    int n = counts[len - 1];
    h->valptr[len] = k;
    h->mincode[len] = code;
    if (n) {
      if (code + n > (1 << len)) return 1; /* over-subscribed */
      if (len <= FAST_BITS) {
        for (i = 0; i < n; i++) {
          int first = (code + i) << (FAST_BITS - len);
          int span = 1 << (FAST_BITS - len), s;
          for (s = 0; s < span; s++) {
            h->fast[first + s] = (unsigned short)((len << 8) | h->symbols[k + i]);
          }
        }
      }
      h->maxcode[len] = code + n - 1;
    } else {
      h->maxcode[len] = -1;
    }
    k += n;
    code = (code + n) << 1;
  }
  h->maxcode[17] = 0x7fffffff;
  h->valid = 1;
  return 0;
}

static int parse_dht(jfront_ctx *c) {
  int len, end;
  if (get_u16(c, &len)) return 1;
  end = c->pos + len - 2;
  if (len < 2 || end > c->size) return fail(c, "Error decoding DHT, segment overruns file.");
  while (c->pos < end) {
    int tc_th, tc, th, i, total = 0, v;
    unsigned char counts[16];
    huff_table *h;
    if (get_u8(c, &tc_th)) return 1;
    tc = tc_th >> 4;
    th = tc_th & 15;
    if (tc > 1) return fail(c, "Error DHT expected Tc value 0 or 1.");
    if (th > 3) return fail(c, "Error DHT expected Th value 0 to 3.");
    h = tc ? &c->ac[th] : &c->dc[th];
    for (i = 0; i < 16; i++) {
      if (get_u8(c, &v)) return 1;
      counts[i] = (unsigned char)v;
      total += v;
    }
    if (total > 256) return fail(c, "Error DHT has more than 256 symbols.");
    if (c->pos + total > end) return fail(c, "Error DHT needs more bytes than available.");
    for (i = 0; i < total; i++) {
      if (get_u8(c, &v)) return 1;
      h->symbols[i] = (unsigned char)v;
      if (!tc && v > 15) return fail(c, "Error invalid DC symbol.");
    }
    if (build_huffman(h, counts)) return fail(c, "Error invalid DHT.");
  }
  return 0;
}

static int parse_sof0(jfront_ctx *c) {
  int len, i, v;
  if (get_u16(c, &len)) return 1;
  if (len < 11) return fail(c, "Error SOF needs at least 9 bytes");
  if (c->have_frame) return fail(c, "Error multiple SOF not supported.");
  if (get_u8(c, &c->bits) || get_u16(c, &c->height) || get_u16(c, &c->width) ||
      get_u8(c, &c->ncomps)) {
    return 1;
  }
  if (c->bits != 8) return fail(c, "Error only 8-bit baseline JPEG is supported.");
  if (c->height == 0) return fail(c, "Error SOF has invalid height.");
  if (c->width == 0) return fail(c, "Error SOF has invalid width.");
  if (len != 8 + 3 * c->ncomps) return fail(c, "Error decoding SOF, unprocessed bytes.");
  if (c->ncomps != 1 && c->ncomps != 3) {
    /* reported by decode_header in the reference's words */
    c->have_frame = 1;
    c->pos += 3 * c->ncomps;
    return 0;
  }
  c->hmax = c->vmax = 0;
  for (i = 0; i < c->ncomps; i++) {
    front_comp *fc = &c->comp[i];
    if (get_u8(c, &fc->id) || get_u8(c, &v) || get_u8(c, &fc->tq)) return 1;
    fc->hsamp = v >> 4;
    fc->vsamp = v & 15;
    if (fc->hsamp == 0 || fc->hsamp > 4) return fail(c, "Error SOF expected Hi value 1 to 4.");
    if (fc->hsamp == 3) return fail(c, "Unsupported horizontal sampling.");
    if (fc->vsamp == 0 || fc->vsamp > 4) return fail(c, "Error SOF expected Vi value 1 to 4.");
    if (fc->vsamp == 3) return fail(c, "Unsupported vertical sampling.");
    if (fc->tq > 3) return fail(c, "Error SOF expected Tq value 0 to 3.");
    if (fc->hsamp > c->hmax) c->hmax = fc->hsamp;
    if (fc->vsamp > c->vmax) c->vmax = fc->vsamp;
  }
  c->nhmb = (c->width + 8 * c->hmax - 1) / (8 * c->hmax);
  c->nvmb = (c->height + 8 * c->vmax - 1) / (8 * c->vmax);
  c->have_frame = 1;
  return 0;
}

static int parse_dri(jfront_ctx *c) {
  int len;
  if (get_u16(c, &len)) return 1;
  if (len != 4) return fail(c, "Error decoding DRI, unprocessed bytes.");
  return get_u16(c, &c->restart_interval);
}

static int skip_segment(jfront_ctx *c) {
  int len;
  if (get_u16(c, &len)) return 1;
  if (len < 2 || c->pos + len - 2 > c->size) return fail(c, "Error skipping past the end of file.");
  c->pos += len - 2;
  return 0;
}

/* Parses marker segments up to (not including) SOS; T.81 B.2. */
static int parse_headers(jfront_ctx *c) {
  if (c->size < 4 || c->buf[0] != 0xFF || c->buf[1] != 0xD8) {
    return fail(c, "Error, not a JPEG (invalid SOI marker).");
  }
  c->pos = 2;
  for (;;) {
    int m;
    if (c->pos + 2 > c->size) return fail(c, "Error underflow reading marker.");
    if (c->buf[c->pos] != 0xFF) return fail(c, "Error, invalid JPEG syntax.");
    while (c->pos < c->size && c->buf[c->pos] == 0xFF) c->pos++; /* fill bytes */
    if (get_u8(c, &m)) return 1;
    switch (m) {
      case 0xDB: if (parse_dqt(c)) return 1; break;
      case 0xC4: if (parse_dht(c)) return 1; break;
      case 0xC0: if (parse_sof0(c)) return 1; break;
      case 0xDD: if (parse_dri(c)) return 1; break;
      case 0xDA: c->sos_pos = c->pos; return 0;
      case 0xD9: return fail(c, "Error reading jpeg headers");
      case 0xC1: case 0xC2: case 0xC3: case 0xC5: case 0xC6: case 0xC7:
      case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
        return fail(c, "Error only baseline sequential (SOF0) JPEG is supported.");
      default:
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) break; /* standalone markers */
        if (skip_segment(c)) return 1;
    }
  }
}

/* ---- entropy-coded segment reader ---------------------------------------- */

/* Tops the accumulator up to at least 32 valid bits.  0xFF00 is a stuffed
 * 0xFF; any other 0xFFxx is a marker: stop there and feed zero bits, the scan
 * loop deals with it at the next MCU boundary (T.81 F.1.2.3). */
static void refill(jfront_ctx *c) {
  while (c->nbits <= 56) {
    unsigned byte = 0;
    if (!c->hit_marker) {
      if (c->ecs_pos >= c->size) {
        c->hit_marker = 1;
      } else {
        byte = c->buf[c->ecs_pos];
        if (byte == 0xFF) {
          int next = c->ecs_pos + 1 < c->size ? c->buf[c->ecs_pos + 1] : 0xD9;
          if (next == 0x00) {
            c->ecs_pos += 2;
          } else {
            c->hit_marker = 1;
            byte = 0;
          }
        } else {
          c->ecs_pos++;
        }
      }
    }
    c->acc |= (unsigned long long)byte << (56 - c->nbits);
    c->nbits += 8;
  }
}

static unsigned peek_bits(const jfront_ctx *c, int n) {
  return (unsigned)(c->acc >> (64 - n));
}

static void drop_bits(jfront_ctx *c, int n) {
  c->acc <<= n;
  c->nbits -= n;
}

static int decode_symbol(jfront_ctx *c, const huff_table *h) {
  unsigned entry;
  int len, code;
  if (c->nbits < 32) refill(c);
  entry = h->fast[peek_bits(c, FAST_BITS)];
  if (entry) {
    drop_bits(c, (int)(entry >> 8));
    return (int)(entry & 0xff);
  }
  code = (int)peek_bits(c, FAST_BITS + 1);
  for (len = FAST_BITS + 1; len <= 16; len++) {
    if (code <= h->maxcode[len]) {
      drop_bits(c, len);
      return h->symbols[h->valptr[len] + code - h->mincode[len]];
    }
    code = (int)peek_bits(c, len + 1);
  }
  drop_bits(c, 16);
  fail(c, "Error invalid Huffman code in scan.");
  return 0;
}

/* T.81 F.2.2.1: `n` additional bits -> signed value. */
static int receive_extend(jfront_ctx *c, int n) {
  int v;
  if (n == 0) return 0;
  if (c->nbits < 32) refill(c);
  v = (int)peek_bits(c, n);
  drop_bits(c, n);
  return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
}

/* ---- scan ------------------------------------------------------------------ */

typedef struct scan_comp {
  const front_comp *fc;
  const huff_table *dc, *ac;
  const unsigned short *q;
  image_plane *plane;
  short pred; /* the reference keeps the DC predictor in a short, src/xjpeg.c:430 */
} scan_comp;

static int restart(jfront_ctx *c, scan_comp *sc, int ncomps, int *expected) {
  int i, m;
  /* discard the padding bits, then expect RSTn (T.81 E.2.4) */
  c->acc = 0;
  c->nbits = 0;
  c->hit_marker = 0;
  while (c->ecs_pos < c->size && c->buf[c->ecs_pos] != 0xFF) c->ecs_pos++; /* resync */
  while (c->ecs_pos + 1 < c->size && c->buf[c->ecs_pos + 1] == 0xFF) c->ecs_pos++;
  if (c->ecs_pos + 1 >= c->size) return 2; /* ran out: treat like EOI */
  m = c->buf[c->ecs_pos + 1];
  if (m == 0xD9) return 2;
  if (m < 0xD0 || m > 0xD7) return fail(c, "Error, unknown marker found in scan.");
  if ((m & 7) != (*expected & 7)) return fail(c, "Error invalid RST counter in marker.");
  c->ecs_pos += 2;
  (*expected)++;
  for (i = 0; i < ncomps; i++) sc[i].pred = 0;
  return 0;
}

/* Parses the SOS header (T.81 B.2.3), binds the scan components to their tables and planes
 * and leaves c->ecs_pos at the first entropy-coded byte with an empty bit accumulator. */
static int scan_setup(jfront_ctx *c, image *img, scan_comp *sc, int *ns_out) {
  int len, ns, i, j, v;
  c->pos = c->sos_pos;
  if (get_u16(c, &len) || get_u8(c, &ns)) return 1;
  if (ns != c->ncomps) {
    return fail(c, ns == 1 || ns == 3 ? "Error multiple SOS not supported."
                                      : "Error only scans with 1 or 3 components supported");
  }
  if (len != 6 + 2 * ns) return fail(c, "Error decoding SOS, unprocessed bytes.");
  for (i = 0; i < ns; i++) {
    int id, tables;
    if (get_u8(c, &id) || get_u8(c, &tables)) return 1;
    sc[i].fc = NULL;
    for (j = 0; j < c->ncomps; j++) {
      if (c->comp[j].id == id) {
        sc[i].fc = &c->comp[j];
        sc[i].plane = &img->plane[j];
        break;
      }
    }
    if (!sc[i].fc) return fail(c, "Error SOS references invalid component.");
    if ((tables >> 4) > 3 || (tables & 15) > 3) return fail(c, "Error SOS table index out of range.");
    sc[i].dc = &c->dc[tables >> 4];
    sc[i].ac = &c->ac[tables & 15];
    if (!sc[i].dc->valid) return fail(c, "Error SOS component references invalid DC entropy table.");
    if (!sc[i].ac->valid) return fail(c, "Error SOS component references invalid AC entropy table.");
    if (!c->quant[sc[i].fc->tq].valid) return fail(c, "Error SOF referenced invalid quantization table.");
    sc[i].q = c->quant[sc[i].fc->tq].tbl;
    sc[i].pred = 0;
  }
  if (get_u8(c, &v) || v != 0) return fail(c, "Error SOS expected Ss value 0.");
  if (get_u8(c, &v) || v != 63) return fail(c, "Error SOS expected Se value 63.");
  if (get_u8(c, &v) || v != 0) return fail(c, "Error SOS expected Ah/Al value 0.");
  c->ecs_pos = c->pos;
  c->acc = 0;
  c->nbits = 0;
  c->hit_marker = 0;
  *ns_out = ns;
  return 0;
}

/* One MCU: the blocks of every scan component in the order of src/xjpeg.c:462-472. */
static int decode_mcu(jfront_ctx *c, scan_comp *sc, int ns, jpeg_decode_out out, int mbx, int mby,
                      short *pack, int *pack_index, long long w0x8) {
  int i, v;
  for (i = 0; i < ns; i++) {
    scan_comp *s = &sc[i];
    image_plane *ip = s->plane;
    int sby, sbx;
    for (sby = 0; sby < s->fc->vsamp; sby++) {
      for (sbx = 0; sbx < s->fc->hsamp; sbx++) {
        short block[64];
        int by = mby * s->fc->vsamp + sby;
        int bx = mbx * s->fc->hsamp + sbx;
        int k = 0, sym;
        memset(block, 0, sizeof(block));
        sym = decode_symbol(c, s->dc);
        s->pred = (short)(s->pred + (short)receive_extend(c, sym & 15));
        if (out == JPEG_DECODE_PACK) {
          /* src/xjpeg.c:484-496 */
          ip->index[by * (ip->ystride >> 3) + bx] = *pack_index;
          ip->packed++;
          pack[(*pack_index)++] = (short)(s->pred & 0xfff);
        }
        block[0] = out == JPEG_DECODE_DCT ? (short)(s->pred * s->q[0]) : s->pred;
        do {
          sym = decode_symbol(c, s->ac);
          if (sym == 0) { /* EOB */
            if (out == JPEG_DECODE_PACK) {
              ip->packed++;
              pack[(*pack_index)++] = 0;
            }
            break;
          }
          k += (sym >> 4) + 1;
          v = receive_extend(c, sym & 15);
          if (k > 63) {
            fail(c, "Error indexing outside block.");
            break;
          }
          if (out == JPEG_DECODE_PACK) {
            ip->packed++;
            pack[(*pack_index)++] = (short)(((sym >> 4) << 12) | (v & 0xfff));
          } else if (out == JPEG_DECODE_DCT) {
            block[kNatural[k]] = (short)((short)v * s->q[kNatural[k]]);
          } else {
            block[kNatural[k]] = (short)v;
          }
        } while (k < 63);
        if (c->error) return 1;
        if (out != JPEG_DECODE_PACK) {
          /* src/xjpeg.c:550-563 */
          long long off = w0x8 * (by >> ip->xdec) +
                          (w0x8 >> ip->xdec) * (by & ((1 << ip->xdec) - 1)) +
                          ((long long)bx << 6);
          memcpy(ip->coef + off, block, sizeof(block));
        }
      }
    }
  }
  return 0;
}

static int decode_scan(jfront_ctx *c, image *img, jpeg_decode_out out) {
  scan_comp sc[NCOMPS_MAX];
  int ns, mbx, mby;
  int mcus_left, rst_expected = 0;
  int pack_index = 0;
  short *pack = img->coef;
  long long w0x8 = (long long)img->plane[0].width << 3;

  if (scan_setup(c, img, sc, &ns)) return 1;
  mcus_left = c->restart_interval;

  for (mby = 0; mby < c->nvmb; mby++) {
    for (mbx = 0; mbx < c->nhmb; mbx++) {
      if (decode_mcu(c, sc, ns, out, mbx, mby, pack, &pack_index, w0x8)) return 1;
      if (c->restart_interval && --mcus_left == 0) {
        int r = restart(c, sc, ns, &rst_expected);
        if (r == 1) return 1;
        if (r == 2) return 0;
        mcus_left = c->restart_interval;
      }
    }
  }
  return 0;
}

/* ---- restart-interval parallel decoding (jgpu_front.h) ---------------------- */

/* Restart intervals are independently decodable (T.81 F.1.1.5: predictors reset, bits byte
 * aligned), so a scan with DRI can be cut at its RSTn markers and the pieces decoded by
 * different threads.  Finds the pieces: seg_pos[k] = first entropy-coded byte of interval k.
 * Returns 0 and nseg >= 1, or 1 when the markers are not where a well-formed file has them
 * (the caller then decodes sequentially, which reports the error the reference's way). */
int jfront_find_segments(jpeg_decode_ctx *ctx, jfront_segments *out) {
  jfront_ctx *c = (jfront_ctx *)ctx;
  jfront_ctx w;
  scan_comp sc[NCOMPS_MAX];
  image dummy;
  int ns, total, nseg, k, pos;
  memset(out, 0, sizeof(*out));
  if (!c->have_frame || !c->sos_pos) return 1;
  w = *c;
  w.error = NULL;
  memset(&dummy, 0, sizeof(dummy));
  if (scan_setup(&w, &dummy, sc, &ns)) return 1;
  total = c->nhmb * c->nvmb;
  out->total_mcus = total;
  if (!c->restart_interval) {
    out->mcus_per_seg = total;
    out->nseg = 1;
    out->seg_pos = (int *)malloc(sizeof(int));
    if (!out->seg_pos) return 1;
    out->seg_pos[0] = w.ecs_pos;
    return 0;
  }
  nseg = (total + c->restart_interval - 1) / c->restart_interval;
  out->mcus_per_seg = c->restart_interval;
  out->seg_pos = (int *)malloc(sizeof(int) * (size_t)nseg);
  if (!out->seg_pos) return 1;
  out->seg_pos[0] = pos = w.ecs_pos;
  for (k = 1; k < nseg; k++) {
    /* next marker that is neither a stuffed 0xFF00 nor a fill byte */
    for (;;) {
      const unsigned char *p;
      if (pos >= c->size) goto malformed;
      p = (const unsigned char *)memchr(c->buf + pos, 0xFF, (size_t)(c->size - pos));
      if (!p || p + 1 >= c->buf + c->size) goto malformed;
      pos = (int)(p - c->buf);
      if (c->buf[pos + 1] == 0x00) { pos += 2; continue; }
      if (c->buf[pos + 1] == 0xFF) { pos += 1; continue; }
      break;
    }
    if (c->buf[pos + 1] != 0xD0 + ((k - 1) & 7)) goto malformed;
    pos += 2;
    out->seg_pos[k] = pos;
  }
  out->nseg = nseg;
  return 0;
malformed:
  free(out->seg_pos);
  memset(out, 0, sizeof(*out));
  return 1;
}

void jfront_segments_free(jfront_segments *s) {
  free(s->seg_pos);
  memset(s, 0, sizeof(*s));
}

/* Decodes restart intervals [s0, s1) into img (QUANT or DCT planes).  Works on a private copy
 * of the reader state, so several threads may run it on the same ctx and image at once: the
 * blocks of different intervals are disjoint.  Returns 0, or 1 with *error set to a static
 * message. */
int jfront_decode_segments(const jpeg_decode_ctx *ctx, image *img, jpeg_decode_out out,
                           const jfront_segments *segs, int s0, int s1, const char **error) {
  jfront_ctx w = *(const jfront_ctx *)ctx;
  scan_comp sc[NCOMPS_MAX];
  int ns, s, i, pack_index = 0;
  const long long w0x8 = (long long)img->plane[0].width << 3;
  w.error = NULL;
  if (out != JPEG_DECODE_QUANT && out != JPEG_DECODE_DCT) {
    *error = "Error, parallel decoding produces quant or dct planes only";
    return 1;
  }
  if (scan_setup(&w, img, sc, &ns)) goto failed;
  for (s = s0; s < s1; s++) {
    int m = s * segs->mcus_per_seg;
    const int m_end = m + segs->mcus_per_seg < segs->total_mcus ? m + segs->mcus_per_seg : segs->total_mcus;
    w.ecs_pos = segs->seg_pos[s];
    w.acc = 0;
    w.nbits = 0;
    w.hit_marker = 0;
    for (i = 0; i < ns; i++) sc[i].pred = 0;
    for (; m < m_end; m++) {
      if (decode_mcu(&w, sc, ns, out, m % w.nhmb, m / w.nhmb, NULL, &pack_index, w0x8)) goto failed;
    }
  }
  return 0;
failed:
  *error = w.error ? w.error : "Error decoding scan";
  return 1;
}

/* ---- preparation for the GPU entropy decoder (jgpu_huff_core.h) ------------------- */

/* Copies `n` bytes, growing *out; returns 1 when the destination is too small. */
static int put_bytes(unsigned char *dst, long long cap, long long *out, const unsigned char *src, long long n) {
  if (*out + n > cap) return 1;
  memcpy(dst + *out, src, (size_t)n);
  *out += n;
  return 0;
}

/* After decode_header: fills what the GPU decoder needs for this file.
 *   stream     the entropy-coded bytes with the stuffed zeros (T.81 F.1.2.3) and the RSTn
 *              markers removed; every restart interval starts on a subsequence boundary and is
 *              zero-padded to the next one (a reader that meets a marker feeds zeros, refill()
 *              above), 16 guard bytes follow the last one
 *   seg_first  subsequence each restart interval starts at, n_seg + 1 entries, then the
 *              entropy-coded bits of each interval, n_seg entries (0xffffffff: not to be checked)
 *   tables     JGPU_HUFF_TABLES decoder tables: (DC, AC) of plane 0, 1, 2
 *   file       geometry fields (n_subseq, n_seg, mcus_per_seg, total_mcus, nhmb, bpm, ncomps,
 *              hs, vs, blk_*); the caller places it in the batch (word0, plane_off, ...)
 * Returns the bytes written to stream, or -1 with *why set when the file is not one the GPU
 * path takes (the caller then uses the sequential reader, which reports errors the
 * reference's way). */
long long jfront_huff_prepare(jpeg_decode_ctx *ctx, int subseq_words, unsigned char *stream,
                              long long stream_cap, unsigned int *seg_first, int seg_cap,
                              jgpu_huff_table *tables, jgpu_huff_file *file, const char **why) {
  jfront_ctx *c = (jfront_ctx *)ctx;
  jfront_ctx w;
  scan_comp sc[NCOMPS_MAX];
  image dummy;
  int ns, i, k, nseg, pos, bpm = 0;
  long long out = 0;
  const long long sub = 4ll * subseq_words;
  *why = "Error decoding scan";
  if (!c->have_frame || !c->sos_pos) return -1;
  w = *c;
  w.error = NULL;
  memset(&dummy, 0, sizeof(dummy));
  if (scan_setup(&w, &dummy, sc, &ns)) {
    *why = w.error ? w.error : *why;
    return -1;
  }
  memset(file->blk_comp, 0, sizeof(file->blk_comp));
  memset(file->blk_dx, 0, sizeof(file->blk_dx));
  memset(file->blk_dy, 0, sizeof(file->blk_dy));
  for (i = 0; i < ns; i++) {
    const int plane = (int)(sc[i].fc - w.comp);
    int dx, dy;
    if (jgpu_huff_build_table(&tables[2 * plane], sc[i].dc->counts, sc[i].dc->symbols, 0) ||
        jgpu_huff_build_table(&tables[2 * plane + 1], sc[i].ac->counts, sc[i].ac->symbols, 1)) {
      *why = "Error invalid DHT.";
      return -1;
    }
    file->hs[plane] = sc[i].fc->hsamp;
    file->vs[plane] = sc[i].fc->vsamp;
    for (dy = 0; dy < sc[i].fc->vsamp; dy++) {
      for (dx = 0; dx < sc[i].fc->hsamp; dx++) {
        if (bpm >= JGPU_HUFF_MAX_BLOCKS) {
          *why = "Error, more than 10 blocks per MCU";
          return -1;
        }
        file->blk_comp[bpm] = (unsigned char)plane;
        file->blk_dx[bpm] = (unsigned char)dx;
        file->blk_dy[bpm] = (unsigned char)dy;
        bpm++;
      }
    }
  }
  file->bpm = bpm;
  file->ncomps = ns;
  file->nhmb = c->nhmb;
  file->total_mcus = c->nhmb * c->nvmb;
  file->mcus_per_seg = c->restart_interval ? c->restart_interval : file->total_mcus;
  nseg = (file->total_mcus + file->mcus_per_seg - 1) / file->mcus_per_seg;
  /* coefficient slots of an interval are counted in 32 bits on the device */
  if ((long long)file->mcus_per_seg * bpm * 64 >= 0x7fffffffll || 2 * nseg + 1 > seg_cap) {
    *why = "Error, scan too large for the GPU entropy decoder";
    return -1;
  }
  pos = w.ecs_pos;
  for (k = 0; k < nseg; k++) {
    const long long seg_start = out;
    seg_first[k] = (unsigned int)(out / sub);
    for (;;) { /* up to the next marker */
      const unsigned char *p;
      long long run;
      if (pos >= c->size) break;
      p = (const unsigned char *)memchr(c->buf + pos, 0xFF, (size_t)(c->size - pos));
      run = p ? p - (c->buf + pos) : c->size - pos;
      if (put_bytes(stream, stream_cap, &out, c->buf + pos, run)) goto too_small;
      pos += (int)run;
      if (!p) break;
      if (pos + 1 < c->size && c->buf[pos + 1] == 0x00) {
        const unsigned char ff = 0xFF;
        if (put_bytes(stream, stream_cap, &out, &ff, 1)) goto too_small;
        pos += 2;
        continue;
      }
      break;
    }
    /* entropy-coded bits of the interval, for the "nothing left over" check: only where the
     * sequential reader looks for a marker after the interval (restart() above) */
    if ((out - seg_start) * 8 >= 0xffffffffll) {
      *why = "Error, scan too large for the GPU entropy decoder";
      return -1;
    }
    seg_first[nseg + 1 + k] = (k + 1 < nseg || (c->restart_interval && file->total_mcus % c->restart_interval == 0))
                                  ? (unsigned int)((out - seg_start) * 8)
                                  : 0xffffffffu;
    { /* zero padding to the next boundary; an empty interval still gets one subsequence */
      long long pad = (sub - (out - seg_start) % sub) % sub;
      if (out == seg_start) pad = sub;
      if (out + pad > stream_cap) goto too_small;
      memset(stream + out, 0, (size_t)pad);
      out += pad;
    }
    if (k + 1 < nseg) { /* RSTn, T.81 E.2.4 */
      while (pos + 1 < c->size && c->buf[pos + 1] == 0xFF) pos++; /* fill bytes */
      if (pos + 1 >= c->size || c->buf[pos] != 0xFF || c->buf[pos + 1] != 0xD0 + (k & 7)) {
        *why = "Error, restart markers are not where the header says";
        return -1;
      }
      pos += 2;
    } else if (c->restart_interval && file->total_mcus % c->restart_interval == 0) {
      /* the sequential reader looks for a marker after a complete last interval too (restart()
       * above): the end of the file, EOI or the next RSTn pass, anything else is its error */
      while (pos + 1 < c->size && c->buf[pos + 1] == 0xFF) pos++;
      if (pos + 1 < c->size && c->buf[pos + 1] != 0xD9 && c->buf[pos + 1] != 0xD0 + (k & 7)) {
        *why = "Error, unknown marker found in scan.";
        return -1;
      }
    }
  }
  seg_first[nseg] = (unsigned int)(out / sub);
  file->n_seg = (unsigned int)nseg;
  file->n_subseq = seg_first[nseg];
  if (out + 16 > stream_cap) goto too_small;
  memset(stream + out, 0, 16);
  out += 16;
  *why = NULL;
  return out;
too_small:
  *why = "Error, stream buffer too small";
  return -1;
}

/* Upper bound of what jfront_huff_prepare writes for this file, and of its restart intervals. */
long long jfront_huff_bound(const jpeg_decode_ctx *ctx, int subseq_words, int *nseg_out) {
  const jfront_ctx *c = (const jfront_ctx *)ctx;
  const long long total = (long long)c->nhmb * c->nvmb;
  const long long per = c->restart_interval ? c->restart_interval : total;
  const long long nseg = per > 0 ? (total + per - 1) / per : 1;
  if (nseg_out) *nseg_out = (int)nseg;
  return (long long)c->size + (nseg + 1) * 4ll * subseq_words + 32;   /* the segment table needs 2*nseg + 1 entries */
}

/* ---- vtable ---------------------------------------------------------------- */

static void front_reset(jfront_ctx *c, jpeg_info *info) {
  memset(c, 0, sizeof(*c));
  c->buf = info->buf;
  c->size = info->size;
}

static jfront_ctx *front_alloc(jpeg_info *info) {
  jfront_ctx *c = (jfront_ctx *)malloc(sizeof(jfront_ctx));
  if (c != NULL) front_reset(c, info);
  return c;
}

/* Same classification as the reference's decode_subsamp (src/jpeg_wrap.c:32-52). */
static jpeg_subsamp classify(const jpeg_header *h) {
  int xdec = 0, ydec = 0, a, b;
  if (h->ncomps == 1) return JPEG_SUBSAMP_MONO;
  for (a = h->comp[0].hsamp, b = h->comp[1].hsamp; a > b; a >>= 1) xdec++;
  for (a = h->comp[0].vsamp, b = h->comp[1].vsamp; a > b; a >>= 1) ydec++;
  if (h->comp[0].hsamp < h->comp[1].hsamp || h->comp[0].vsamp < h->comp[1].vsamp) {
    return JPEG_SUBSAMP_UNKNOWN;
  }
  if (xdec == 0 && ydec == 0) return JPEG_SUBSAMP_444;
  if (xdec == 1 && ydec == 0) return JPEG_SUBSAMP_422;
  if (xdec == 1 && ydec == 1) return JPEG_SUBSAMP_420;
  if (xdec == 0 && ydec == 1) return JPEG_SUBSAMP_440;
  if (xdec == 2 && ydec == 0) return JPEG_SUBSAMP_411;
  return JPEG_SUBSAMP_UNKNOWN;
}

static int front_header(jfront_ctx *c, jpeg_header *h) {
  int i;
  if (parse_headers(c)) {
    fprintf(stderr, "%s\n", c->error);
    return EXIT_FAILURE;
  }
  if (!c->have_frame) {
    fprintf(stderr, "Error reading jpeg headers\n");
    return EXIT_FAILURE;
  }
  if (c->ncomps != 1 && c->ncomps != 3) {
    fprintf(stderr, "Unsupported number of components %i\n", c->ncomps);
    return EXIT_FAILURE;
  }
  h->width = c->width;
  h->height = c->height;
  h->bits = c->bits;
  h->ncomps = c->ncomps;
  h->restart_interval = c->restart_interval;
  for (i = 0; i < NQUANT_MAX; i++) {
    h->quant[i].valid = c->quant[i].valid;
    if (c->quant[i].valid) {
      h->quant[i].bits = c->quant[i].bits;
      memcpy(h->quant[i].tbl, c->quant[i].tbl, sizeof(h->quant[i].tbl));
    }
  }
  for (i = 0; i < c->ncomps; i++) {
    jpeg_component *comp = &h->comp[i];
    comp->hblocks = c->nhmb * c->comp[i].hsamp;
    comp->vblocks = c->nvmb * c->comp[i].vsamp;
    comp->hsamp = c->comp[i].hsamp;
    comp->vsamp = c->comp[i].vsamp;
    if (!c->quant[c->comp[i].tq].valid) {
      fprintf(stderr, "Invalid quantization table for components %i\n", i);
      return EXIT_FAILURE;
    }
    comp->quant = &h->quant[c->comp[i].tq];
  }
  h->subsamp = classify(h);
  return EXIT_SUCCESS;
}

static int front_image(jfront_ctx *c, image *img, jpeg_decode_out out) {
  int i;
  switch (out) {
    case JPEG_DECODE_PACK:
    case JPEG_DECODE_QUANT:
    case JPEG_DECODE_DCT:
      break;
    default:
      fprintf(stderr, "Unsupported output '%s' for jfront wrapper.\n",
              out == JPEG_DECODE_YUV ? "yuv" : out == JPEG_DECODE_RGB ? "rgb" : "?");
      return EXIT_FAILURE;
  }
  if (!c->have_frame || !c->sos_pos) {
    fprintf(stderr, "Error, decode_image called before decode_header\n");
    return EXIT_FAILURE;
  }
  if (img->nplanes != c->ncomps || img->coef == NULL) {
    fprintf(stderr, "Error, image surface does not match the jpeg header\n");
    return EXIT_FAILURE;
  }
  for (i = 0; i < c->ncomps; i++) {
    if (img->plane[i].width != c->nhmb * c->comp[i].hsamp * 8 ||
        img->plane[i].height != c->nvmb * c->comp[i].vsamp * 8) {
      fprintf(stderr, "Error, image surface does not match the jpeg header\n");
      return EXIT_FAILURE;
    }
  }
  if (decode_scan(c, img, out)) {
    fprintf(stderr, "%s\n", c->error ? c->error : "Error decoding scan");
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}

static void front_free(jfront_ctx *c) { free(c); }

const jpeg_decode_ctx_vtbl JFRONT_DECODE_CTX_VTBL = {
    (jpeg_decode_alloc_func)front_alloc,
    (jpeg_decode_header_func)front_header,
    (jpeg_decode_image_func)front_image,
    (jpeg_decode_reset_func)front_reset,
    (jpeg_decode_free_func)front_free};
