/* Host-side geometry, output-surface helpers and error plumbing (plain C).
 *
 * Geometry follows the reference exactly:
 *   MCU counts            src/xjpeg.c:400-407
 *   hblocks / vblocks     src/jpeg_wrap.c:301-308
 *   plane sizes, xdec/ydec, cstride, running coef pointer   src/image.c:38-95
 */
#define _POSIX_C_SOURCE 200809L
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "jgpu_internal.h"

static __thread char g_err[512];

int jgpu_fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return EXIT_FAILURE;
}

const char *jgpu_last_error(void) { return g_err; }

/* Bits needed to represent v: 1->1, 2->2, 4->3 (src/internal.c:49-67). */
static int ilog(unsigned v) {
  int n = 0;
  for (; v; v >>= 1) n++;
  return n;
}

int jgpu_layout_query(const jgpu_image_desc *d, jgpu_layout *out) {
  int i;
  int64_t coef = 0, data = 0;
  if (d == NULL || out == NULL) return jgpu_fail("jgpu_layout_query: NULL argument");
  memset(out, 0, sizeof(*out));
  if (d->ncomps != 1 && d->ncomps != 3) {
    return jgpu_fail("Unsupported number of components %i", d->ncomps);
  }
  if (d->width < 1 || d->height < 1 || d->width > 65535 || d->height > 65535) {
    return jgpu_fail("Unsupported image size %ix%i", d->width, d->height);
  }
  for (i = 0; i < d->ncomps; i++) {
    int h = d->hsamp[i], v = d->vsamp[i];
    if ((h != 1 && h != 2 && h != 4) || (v != 1 && v != 2 && v != 4)) {
      return jgpu_fail("Unsupported sampling %ix%i for component %i", h, v, i);
    }
    if (d->tq[i] < 0 || d->tq[i] >= NQUANT_MAX) {
      return jgpu_fail("Invalid quantization table for components %i", i);
    }
    if (h > out->hmax) out->hmax = h;
    if (v > out->vmax) out->vmax = v;
  }
  if (d->hsamp[0] != out->hmax) {
    return jgpu_fail("Unsupported sampling: component 0 must have the largest "
                     "horizontal factor (coefficient layout, src/xjpeg.c:558)");
  }
  if (d->vsamp[0] != out->vmax) {
    /* the reference's colour pass reads luma row t of plane 0 for output row t
     * (res/unyuv.fs.glsl:29-31, res/yuv.fs.glsl:17-19): a vertically decimated plane 0 has no
     * defined rendering there either */
    return jgpu_fail("Unsupported sampling: component 0 must have the largest vertical factor");
  }
  out->nhmb = (d->width + 8 * out->hmax - 1) / (8 * out->hmax);
  out->nvmb = (d->height + 8 * out->vmax - 1) / (8 * out->vmax);
  for (i = 0; i < d->ncomps; i++) {
    jgpu_plane_layout *p = &out->plane[i];
    p->hblocks = out->nhmb * d->hsamp[i];
    p->vblocks = out->nvmb * d->vsamp[i];
    p->width = p->hblocks * 8;
    p->height = p->vblocks * 8;
    if (p->width > 65535 || p->height > 65535) {
      return jgpu_fail("Padded plane %i is %ix%i; the reference's image_plane "
                       "holds unsigned short sizes", i, p->width, p->height);
    }
    p->xdec = ilog((unsigned)out->hmax) - ilog((unsigned)d->hsamp[i]);
    p->ydec = ilog((unsigned)out->vmax) - ilog((unsigned)d->vsamp[i]);
    p->cstride = (p->vblocks + ((1 << p->xdec) - 1)) >> p->xdec;
    p->coef_off = coef;
    p->data_off = data;
    coef += ((int64_t)p->width << (p->xdec + 3)) * p->cstride;
    data += (int64_t)p->width * p->height;
    out->coded_blocks += (int64_t)p->hblocks * p->vblocks;
  }
  out->coef_len = coef;
  out->data_len = data;
  out->rgb_len = (int64_t)d->width * d->height * (d->ncomps == 1 ? 1 : 3);
  return EXIT_SUCCESS;
}

int jgpu_desc_from_header(const jpeg_header *h, jgpu_image_desc *d) {
  int i;
  memset(d, 0, sizeof(*d));
  d->width = h->width;
  d->height = h->height;
  d->ncomps = h->ncomps;
  if (h->ncomps != 1 && h->ncomps != 3) {
    return jgpu_fail("Unsupported number of components %i", h->ncomps);
  }
  for (i = 0; i < h->ncomps; i++) {
    const jpeg_component *c = &h->comp[i];
    ptrdiff_t slot;
    d->hsamp[i] = c->hsamp;
    d->vsamp[i] = c->vsamp;
    if (c->quant == NULL) {
      return jgpu_fail("Missing quantization table for components %i", i);
    }
    slot = c->quant - h->quant;
    if (slot < 0 || slot >= NQUANT_MAX || !h->quant[slot].valid) {
      return jgpu_fail("Invalid quantization table for components %i", i);
    }
    d->tq[i] = (int32_t)slot;
  }
  d->yuv_off = -1;
  return EXIT_SUCCESS;
}

/* ---- aligned memory ------------------------------------------------------ */

void *jgpu_aligned_malloc(size_t bytes) {
  void *p = NULL;
  if (bytes == 0) bytes = 16;
  if (posix_memalign(&p, 16, bytes) != 0) return NULL;
  return p;
}

void jgpu_aligned_free(void *p) { free(p); }

/* ---- output surface: semantics of src/image.c --------------------------- */

/* Page-locked `pixels` for surfaces of callers that decode through the CUDA backend: the
 * read-back of a 4K frame into pageable memory costs 0.95 ms more than into pinned memory
 * (2.68 vs 1.73 ms per frame).  Asked for per surface (jgpu_image_init_ex) or, for callers that
 * keep the reference's two-argument image_init shape, per THREAD (jgpu_image_set_pinned): no
 * process-wide switch.  Which kind of memory a surface got is asked of the CUDA runtime when it is
 * freed (jgpu_host_is_pinned), so nothing is remembered beside the reference's `image` struct. */
static __thread int t_pinned_surfaces = 0;

void jgpu_image_set_pinned(int on) { t_pinned_surfaces = on != 0; }

static void *surface_alloc(size_t bytes, int pinned) {
  if (pinned) {
    void *p = jgpu_host_alloc(bytes);
    if (p != NULL) return p;   /* else: ordinary memory, the copy is merely slower */
  }
  return jgpu_aligned_malloc(bytes);
}

static void surface_free(void *p) {
  if (p == NULL) return;
  if (jgpu_host_is_pinned(p)) jgpu_host_free(p); else jgpu_aligned_free(p);
}

int jgpu_image_init(image *img, jpeg_header *header) {
  return jgpu_image_init_ex(img, header, t_pinned_surfaces ? JGPU_IMAGE_PINNED : 0u);
}

int jgpu_image_init_ex(image *img, jpeg_header *header, unsigned flags) {
  jgpu_image_desc d;
  jgpu_layout lay;
  int i;
  int64_t blocks = 0, index_off = 0;
  memset(img, 0, sizeof(*img));
  /* geometry straight from the header's hblocks/vblocks, as image_init does */
  memset(&d, 0, sizeof(d));
  d.width = header->width;
  d.height = header->height;
  d.ncomps = header->ncomps;
  for (i = 0; i < header->ncomps && i < NCOMPS_MAX; i++) {
    d.hsamp[i] = header->comp[i].hsamp;
    d.vsamp[i] = header->comp[i].vsamp;
  }
  if (jgpu_layout_query(&d, &lay) != EXIT_SUCCESS) {
    fprintf(stderr, "%s\n", jgpu_last_error());
    return EXIT_FAILURE;
  }
  img->width = (unsigned short)header->width;
  img->height = (unsigned short)header->height;
  img->nplanes = header->ncomps;
  for (i = 0; i < img->nplanes; i++) {
    image_plane *p = &img->plane[i];
    const jpeg_component *c = &header->comp[i];
    if (c->hblocks != lay.plane[i].hblocks || c->vblocks != lay.plane[i].vblocks) {
      fprintf(stderr, "Inconsistent block counts for component %i\n", i);
      jgpu_image_clear(img);
      return EXIT_FAILURE;
    }
    p->bitdepth = 0; /* the reference leaves it zero (memset, never assigned) */
    p->width = (unsigned short)lay.plane[i].width;
    p->height = (unsigned short)lay.plane[i].height;
    p->xstride = 1;
    p->ystride = p->width;
    p->xdec = (unsigned char)lay.plane[i].xdec;
    p->ydec = (unsigned char)lay.plane[i].ydec;
    p->cstride = lay.plane[i].cstride;
    p->data = (unsigned char *)jgpu_aligned_malloc((size_t)p->ystride * p->height);
    if (p->data == NULL) {
      jgpu_image_clear(img);
      return EXIT_FAILURE;
    }
    blocks += ((int64_t)c->hblocks << p->xdec) * p->cstride;
  }
  img->pixels = (unsigned char *)surface_alloc((size_t)img->width * img->height * 3, (flags & JGPU_IMAGE_PINNED) != 0);
  img->coef = (short *)jgpu_aligned_malloc((size_t)blocks * 64 * sizeof(short));
  img->index = (int *)jgpu_aligned_malloc((size_t)blocks * sizeof(int));
  if (img->pixels == NULL || img->coef == NULL || img->index == NULL) {
    jgpu_image_clear(img);
    return EXIT_FAILURE;
  }
  for (i = 0; i < img->nplanes; i++) {
    image_plane *p = &img->plane[i];
    p->coef = img->coef + lay.plane[i].coef_off;
    p->index = img->index + index_off;
    index_off += ((int64_t)header->comp[i].hblocks << p->xdec) * p->cstride;
  }
  return EXIT_SUCCESS;
}

void jgpu_image_zero(image *img) {
  int i;
  int64_t blocks = 0;
  for (i = 0; i < img->nplanes; i++) {
    image_plane *p = &img->plane[i];
    memset(p->data, 0, (size_t)p->ystride * p->height);
    blocks += (int64_t)((p->width >> 3) << p->xdec) * p->cstride;
    p->packed = 0;
  }
  memset(img->pixels, 0, (size_t)img->width * img->height * 3);
  memset(img->coef, 0, (size_t)blocks * 64 * sizeof(short));
  memset(img->index, 0, (size_t)blocks * sizeof(int));
}

void jgpu_image_clear(image *img) {
  int i;
  for (i = 0; i < img->nplanes && i < NPLANES_MAX; i++) {
    jgpu_aligned_free(img->plane[i].data);
  }
  surface_free(img->pixels);
  jgpu_aligned_free(img->coef);
  jgpu_aligned_free(img->index);
  memset(img, 0, sizeof(*img));
}

/* ---- file slurp: semantics of src/jpeg_info.c:31-61 ---------------------- */

int jgpu_info_init(jpeg_info *info, const char *name) {
  FILE *fp = fopen(name, "rb");
  long size;
  if (fp == NULL) {
    fprintf(stderr, "Error, could not open jpeg file %s\n", name);
    return EXIT_FAILURE;
  }
  info->buf = NULL;
  info->size = 0;
  fseek(fp, 0, SEEK_END);
  size = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  info->buf = (unsigned char *)malloc(size > 0 ? (size_t)size : 1);
  if (info->buf == NULL) {
    fprintf(stderr, "Error, could not allocate %li bytes\n", size);
    fclose(fp);
    return EXIT_FAILURE;
  }
  info->size = (int)size;
  if ((long)fread(info->buf, 1, (size_t)size, fp) != size) {
    fprintf(stderr, "Error reading jpeg file %s\n", name);
    fclose(fp);
    jgpu_info_clear(info);
    return EXIT_FAILURE;
  }
  fclose(fp);
  return EXIT_SUCCESS;
}

void jgpu_info_clear(jpeg_info *info) {
  free(info->buf);
  info->buf = NULL;
  info->size = 0;
}
