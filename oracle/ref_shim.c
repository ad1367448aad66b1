/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Thin C driver around the UNMODIFIED reference objects.  oracle/Makefile
 * compiles /root/reference/src/{dct,xjpeg,image,internal,logging,jpeg_info}.c
 * where they lie (no reference source is copied into this repository) and
 * links this file against them into oracle/_ref/libjgpu_ref.so.  The shim only
 * marshals between flat ctypes-friendly buffers and the reference's own
 * entry points:
 *   xjpeg_init / xjpeg_decode_header / xjpeg_decode_image  (src/xjpeg.h:143-146)
 *   image_init / image_clear                               (src/image.h:53-55)
 *   glj_real_idct8x8                                       (src/dct.h:19)
 */
#include <stdlib.h>
#include <string.h>
#include "xjpeg.h"
#include "dct.h"

/* hdr[0..6] = width,height,bits,ncomps,restart_interval,nhmb,nvmb
 * hdr[8+3*i..] = hsamp,vsamp,tq of component i ; qtabs = [4][64] natural order,
 * qvalid[4].  Mirrors what xjpeg_decode_header_ copies out
 * (src/jpeg_wrap.c:263-319) without needing jpeg_wrap.c (it needs jpeglib.h). */
int refshim_probe(const unsigned char *jpg, int size, int *hdr,
                  unsigned short *qtabs, int *qvalid) {
  xjpeg_decode_ctx ctx;
  int i;
  xjpeg_init(&ctx, jpg, size);
  xjpeg_decode_header(&ctx);
  if (ctx.error || !ctx.frame.valid) return 1;
  memset(hdr, 0, 20 * sizeof(int));
  hdr[0] = ctx.frame.width;
  hdr[1] = ctx.frame.height;
  hdr[2] = ctx.frame.bits;
  hdr[3] = ctx.frame.ncomps;
  hdr[4] = ctx.restart_interval;
  hdr[5] = ctx.frame.nhmb;
  hdr[6] = ctx.frame.nvmb;
  for (i = 0; i < ctx.frame.ncomps && i < NCOMPS_MAX; i++) {
    hdr[8 + 3 * i] = ctx.frame.comp[i].hsamp;
    hdr[9 + 3 * i] = ctx.frame.comp[i].vsamp;
    hdr[10 + 3 * i] = ctx.frame.comp[i].tq;
  }
  for (i = 0; i < NQUANT_MAX; i++) {
    qvalid[i] = ctx.quant[i].valid;
    memcpy(qtabs + 64 * i, ctx.quant[i].tbl, 64 * sizeof(unsigned short));
  }
  return 0;
}

static void fill_header(const xjpeg_decode_ctx *ctx, jpeg_header *h) {
  int i;
  memset(h, 0, sizeof(*h));
  h->width = ctx->frame.width;
  h->height = ctx->frame.height;
  h->bits = ctx->frame.bits;
  h->ncomps = ctx->frame.ncomps;
  h->restart_interval = ctx->restart_interval;
  for (i = 0; i < h->ncomps; i++) {
    h->comp[i].hblocks = ctx->frame.nhmb * ctx->frame.comp[i].hsamp;
    h->comp[i].vblocks = ctx->frame.nvmb * ctx->frame.comp[i].vsamp;
    h->comp[i].hsamp = ctx->frame.comp[i].hsamp;
    h->comp[i].vsamp = ctx->frame.comp[i].vsamp;
    h->comp[i].quant = &h->quant[ctx->frame.comp[i].tq];
  }
}

/* Runs the reference decoder on an in-memory JPEG.  out_mode uses the
 * reference's numbering (src/xjpeg.h:136-142): 1 QUANT, 2 DCT, 3 YUV.
 * QUANT/DCT: copies image.coef (coef_cap shorts at most) ; YUV: copies the
 * three padded planes back to back into planes. */
int refshim_decode(const unsigned char *jpg, int size, int out_mode,
                   short *coef, long long coef_cap, unsigned char *planes,
                   long long planes_cap) {
  xjpeg_decode_ctx ctx;
  jpeg_header h;
  image img;
  int i;
  long long blocks = 0, off = 0;
  xjpeg_init(&ctx, jpg, size);
  xjpeg_decode_header(&ctx);
  if (ctx.error || !ctx.frame.valid) return 1;
  fill_header(&ctx, &h);
  if (image_init(&img, &h) != EXIT_SUCCESS) return 2;
  image_zero(&img);
  xjpeg_decode_image(&ctx, &img, (xjpeg_decode_out)out_mode);
  if (ctx.error) { image_clear(&img); return 3; }
  for (i = 0; i < img.nplanes; i++) {
    image_plane *p = &img.plane[i];
    blocks += (long long)((p->width >> 3) << p->xdec) * p->cstride;
  }
  if (out_mode == XJPEG_DECODE_QUANT || out_mode == XJPEG_DECODE_DCT) {
    if (blocks * 64 > coef_cap) { image_clear(&img); return 4; }
    memcpy(coef, img.coef, (size_t)blocks * 64 * sizeof(short));
  } else if (out_mode == XJPEG_DECODE_YUV) {
    for (i = 0; i < img.nplanes; i++) {
      image_plane *p = &img.plane[i];
      long long n = (long long)p->ystride * p->height;
      if (off + n > planes_cap) { image_clear(&img); return 4; }
      memcpy(planes + off, p->data, (size_t)n);
      off += n;
    }
  } else {
    image_clear(&img);
    return 5;
  }
  image_clear(&img);
  return 0;
}

/* JPEG_DECODE_PACK as the reference's reader writes it (src/xjpeg.c:484-496,513-519,
 * 531-535): copies the packed words (image.coef, img.packed of them), the whole index
 * array (image.index, one int per allocated block) and the per-plane word counts
 * (image_plane.packed).  total[0] = words, total[1] = index entries. */
int refshim_decode_pack(const unsigned char *jpg, int size, short *pack,
                        long long pack_cap, int *index, long long index_cap,
                        int *packed, long long *total) {
  xjpeg_decode_ctx ctx;
  jpeg_header h;
  image img;
  int i;
  long long blocks = 0, words = 0;
  xjpeg_init(&ctx, jpg, size);
  xjpeg_decode_header(&ctx);
  if (ctx.error || !ctx.frame.valid) return 1;
  fill_header(&ctx, &h);
  if (image_init(&img, &h) != EXIT_SUCCESS) return 2;
  image_zero(&img);
  xjpeg_decode_image(&ctx, &img, XJPEG_DECODE_PACK);
  if (ctx.error) { image_clear(&img); return 3; }
  for (i = 0; i < img.nplanes; i++) {
    image_plane *p = &img.plane[i];
    blocks += (long long)((p->width >> 3) << p->xdec) * p->cstride;
    packed[i] = p->packed;
    words += p->packed;
  }
  if (words > pack_cap || blocks > index_cap) { image_clear(&img); return 4; }
  memcpy(pack, img.coef, (size_t)words * sizeof(short));
  memcpy(index, img.index, (size_t)blocks * sizeof(int));
  total[0] = words;
  total[1] = blocks;
  image_clear(&img);
  return 0;
}

/* Geometry as the reference's image_init computes it, for pinning our own
 * layout code.  out[0] = allocated blocks, then per plane (stride 8):
 * width,height,xdec,ydec,ystride,cstride,coef offset (shorts),index offset. */
int refshim_layout(int width, int height, int ncomps, const int *hsamp,
                   const int *vsamp, long long *out) {
  jpeg_header h;
  image img;
  int i, hmax = 0, vmax = 0, nhmb, nvmb;
  long long blocks = 0;
  memset(&h, 0, sizeof(h));
  h.bits = 8; h.width = width; h.height = height; h.ncomps = ncomps;
  for (i = 0; i < ncomps; i++) {
    if (hsamp[i] > hmax) hmax = hsamp[i];
    if (vsamp[i] > vmax) vmax = vsamp[i];
  }
  nhmb = (width + (hmax << 3) - 1) / (hmax << 3);   /* src/xjpeg.c:403-407 */
  nvmb = (height + (vmax << 3) - 1) / (vmax << 3);
  for (i = 0; i < ncomps; i++) {
    h.comp[i].hblocks = nhmb * hsamp[i];             /* src/jpeg_wrap.c:305-306 */
    h.comp[i].vblocks = nvmb * vsamp[i];
    h.comp[i].hsamp = hsamp[i];
    h.comp[i].vsamp = vsamp[i];
    h.comp[i].quant = &h.quant[0];
  }
  if (image_init(&img, &h) != EXIT_SUCCESS) return 1;
  for (i = 0; i < img.nplanes; i++) {
    image_plane *p = &img.plane[i];
    long long *o = out + 1 + 8 * i;
    o[0] = p->width; o[1] = p->height; o[2] = p->xdec; o[3] = p->ydec;
    o[4] = p->ystride; o[5] = p->cstride;
    o[6] = (long long)(p->coef - img.coef);
    o[7] = (long long)(p->index - img.index);
    blocks += (long long)((p->width >> 3) << p->xdec) * p->cstride;
  }
  out[0] = blocks;
  image_clear(&img);
  return 0;
}

void refshim_idct_blocks(short *blocks, long long nblocks) {
  long long b;
  for (b = 0; b < nblocks; b++) {
    glj_real_idct8x8(blocks + 64 * b, 8, blocks + 64 * b, 8);
  }
}
