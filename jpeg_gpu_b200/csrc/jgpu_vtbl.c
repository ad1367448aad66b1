/* jgpu_vtbl.c — CUDA_DECODE_CTX_VTBL, the third decoder backend.
 *
 * Same five slots, call protocol and error convention as the reference's two
 * backends (src/jpeg_wrap.h:35-54, src/jpeg_wrap.c:246-252,352-358):
 *   alloc -> header -> [caller: image_init] -> image -> free      first frame
 *   reset -> header -> image                                      steady state
 *                                         (src/jpeg_gpu.c:612-704,1231-1237)
 * A CPU front end (our jfront, or the reference's xjpeg via
 * cuda_decode_set_frontend) turns the file into QUANT coefficient planes in
 * img->coef; the GPU does dequantise -> IDCT -> upsample -> colour and the
 * result lands in img->plane[i].data (YUV) or img->pixels (RGB).  That is the
 * work of the reference's per-frame GL sequence src/jpeg_gpu.c:1320-1363, with
 * the difference that pixels come back to the host surface.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "jgpu_internal.h"

static const jpeg_decode_ctx_vtbl *g_frontend = NULL;
static int g_device = -1;

void cuda_decode_set_frontend(const jpeg_decode_ctx_vtbl *frontend) { g_frontend = frontend; }
void cuda_decode_set_device(int device) { g_device = device; }

typedef struct cuda_decode_ctx {
  jpeg_decode_ctx_vtbl front;
  jpeg_decode_ctx *front_ctx;
  jgpu_ctx *gpu;        /* created on the first YUV/RGB decode, kept across resets */
  int device;
  int have_header;
  jpeg_header header;   /* copy taken in decode_header; decode_image needs the tables */
} cuda_decode_ctx;

static cuda_decode_ctx *cuda_decode_alloc(jpeg_info *info) {
  cuda_decode_ctx *ctx = (cuda_decode_ctx *)calloc(1, sizeof(cuda_decode_ctx));
  if (ctx == NULL) return NULL;
  ctx->front = g_frontend ? *g_frontend : JFRONT_DECODE_CTX_VTBL;
  ctx->device = g_device;
  if (ctx->device < 0) {
    const char *env = getenv("JGPU_DEVICE");
    ctx->device = env ? atoi(env) : 0;
  }
  ctx->front_ctx = (*ctx->front.decode_alloc)(info);
  if (ctx->front_ctx == NULL) {
    free(ctx);
    return NULL;
  }
  return ctx;
}

static int cuda_decode_header(cuda_decode_ctx *ctx, jpeg_header *header) {
  int i;
  ctx->have_header = 0;
  if ((*ctx->front.decode_header)(ctx->front_ctx, header) != EXIT_SUCCESS) {
    return EXIT_FAILURE;
  }
  ctx->header = *header;
  for (i = 0; i < header->ncomps && i < NCOMPS_MAX; i++) {
    /* re-point the table references into our copy */
    if (header->comp[i].quant != NULL) {
      ctx->header.comp[i].quant = &ctx->header.quant[header->comp[i].quant - header->quant];
    }
  }
  ctx->have_header = 1;
  return EXIT_SUCCESS;
}

static int cuda_decode_image(cuda_decode_ctx *ctx, image *img, jpeg_decode_out out) {
  switch (out) {
    case JPEG_DECODE_PACK:
    case JPEG_DECODE_QUANT:
    case JPEG_DECODE_DCT:
      /* CPU-side formats: the front end's business */
      return (*ctx->front.decode_image)(ctx->front_ctx, img, out);
    case JPEG_DECODE_YUV:
    case JPEG_DECODE_RGB:
      break;
    default:
      fprintf(stderr, "Unsupported output %i for cuda wrapper.\n", (int)out);
      return EXIT_FAILURE;
  }
  if (!ctx->have_header) {
    fprintf(stderr, "Error, decode_image called before decode_header\n");
    return EXIT_FAILURE;
  }
  if ((*ctx->front.decode_image)(ctx->front_ctx, img, JPEG_DECODE_QUANT) != EXIT_SUCCESS) {
    return EXIT_FAILURE;
  }
  if (ctx->gpu == NULL) {
    ctx->gpu = jgpu_create(ctx->device);
    if (ctx->gpu == NULL) {
      fprintf(stderr, "%s\n", jgpu_last_error());
      return EXIT_FAILURE;
    }
  }
  if (jgpu_decode_image(ctx->gpu, &ctx->header, img, out) != EXIT_SUCCESS) {
    fprintf(stderr, "%s\n", jgpu_last_error());
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}

static void cuda_decode_reset(cuda_decode_ctx *ctx, jpeg_info *info) {
  /* device buffers, streams and the cached plan survive; only the parser restarts */
  (*ctx->front.decode_reset)(ctx->front_ctx, info);
  ctx->have_header = 0;
}

static void cuda_decode_free(cuda_decode_ctx *ctx) {
  if (ctx == NULL) return;
  (*ctx->front.decode_free)(ctx->front_ctx);
  jgpu_destroy(ctx->gpu);
  free(ctx);
}

const jpeg_decode_ctx_vtbl CUDA_DECODE_CTX_VTBL = {
    (jpeg_decode_alloc_func)cuda_decode_alloc,
    (jpeg_decode_header_func)cuda_decode_header,
    (jpeg_decode_image_func)cuda_decode_image,
    (jpeg_decode_reset_func)cuda_decode_reset,
    (jpeg_decode_free_func)cuda_decode_free};
