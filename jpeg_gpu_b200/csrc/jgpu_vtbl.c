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
#include <stdint.h>
#include "jgpu_internal.h"

/* The calling thread's options (include/jpeg_gpu_b200.h, cuda_decode_options): thread-local, so
 * two threads configuring and allocating contexts do not see each other's settings. */
static __thread cuda_decode_options t_opt = {NULL, -1, JPEG_DECODE_QUANT, -1};

void cuda_decode_get_options(cuda_decode_options *opt) {
  if (opt != NULL) *opt = t_opt;
}
void cuda_decode_set_frontend(const jpeg_decode_ctx_vtbl *frontend) { t_opt.frontend = frontend; }
void cuda_decode_set_device(int device) { t_opt.device = device; }

int cuda_decode_set_upload(jpeg_decode_out format) {
  if (format != JPEG_DECODE_QUANT && format != JPEG_DECODE_PACK) {
    fprintf(stderr, "Unsupported upload format %i for cuda wrapper.\n", (int)format);
    return EXIT_FAILURE;
  }
  t_opt.upload = format;
  return EXIT_SUCCESS;
}

/* Where RGB / YUV decodes do their Huffman decoding: 0 = the CPU front end (default: it is the
 * pluggable part), 1 = on the device (jgpu_huff.cu), for the built-in front end only. */
int cuda_decode_set_entropy(int on_device) {
  if (on_device != 0 && on_device != 1) {
    fprintf(stderr, "Unsupported entropy decoder %i for cuda wrapper.\n", on_device);
    return EXIT_FAILURE;
  }
  t_opt.entropy_on_device = on_device;
  return EXIT_SUCCESS;
}

typedef struct cuda_decode_ctx {
  jpeg_decode_ctx_vtbl front;
  jpeg_decode_ctx *front_ctx;
  const unsigned char *buf; /* the caller's file bytes (jpeg_info.buf), for the device entropy decoder */
  int size;
  int entropy_on_device;
  uint8_t *planes;          /* pinned staging for YUV output of the device entropy path (grow-only) */
  int64_t planes_cap;
  jgpu_ctx *gpu;        /* created on the first YUV/RGB decode, kept across resets */
  int device;
  int have_header;
  jpeg_decode_out upload; /* what crosses to the device: QUANT planes or the PACK stream */
  jpeg_header header;   /* copy taken in decode_header; decode_image needs the tables */
} cuda_decode_ctx;

jpeg_decode_ctx *cuda_decode_alloc_ex(jpeg_info *info, const cuda_decode_options *opt) {
  cuda_decode_ctx *ctx;
  if (info == NULL || opt == NULL) {
    jgpu_fail("cuda_decode_alloc_ex: NULL argument");
    return NULL;
  }
  if (opt->upload != JPEG_DECODE_QUANT && opt->upload != JPEG_DECODE_PACK) {
    jgpu_fail("cuda_decode_alloc_ex: unsupported upload format %i", (int)opt->upload);
    return NULL;
  }
  ctx = (cuda_decode_ctx *)calloc(1, sizeof(cuda_decode_ctx));
  if (ctx == NULL) return NULL;
  ctx->front = opt->frontend ? *opt->frontend : JFRONT_DECODE_CTX_VTBL;
  ctx->device = opt->device;
  ctx->upload = opt->upload;
  {
    int on = opt->entropy_on_device;
    if (on < 0) {
      const char *env = getenv("JGPU_ENTROPY");
      on = env != NULL && strcmp(env, "gpu") == 0;
    }
    ctx->entropy_on_device = on && opt->frontend == NULL;
  }
  ctx->buf = info->buf;
  ctx->size = info->size;
  if (ctx->device < 0) {
    const char *env = getenv("JGPU_DEVICE");
    ctx->device = env ? atoi(env) : 0;
  }
  ctx->front_ctx = (*ctx->front.decode_alloc)(info);
  if (ctx->front_ctx == NULL) {
    free(ctx);
    return NULL;
  }
  return (jpeg_decode_ctx *)ctx;
}

/* The table's slot: the calling thread's options. */
static cuda_decode_ctx *cuda_decode_alloc(jpeg_info *info) {
  return (cuda_decode_ctx *)cuda_decode_alloc_ex(info, &t_opt);
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
  int i, rc;
  int64_t words = 0;
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
  if (ctx->entropy_on_device && out == JPEG_DECODE_RGB && img->pixels != NULL) {
    /* the file's bytes go to the device as they are; nothing is decoded on the host */
    jgpu_jpeg file;
    jgpu_jpeg_info jinfo;
    const int64_t cap = (int64_t)img->width * img->height * 3;
    if (ctx->gpu == NULL) {
      ctx->gpu = jgpu_create(ctx->device);
      if (ctx->gpu == NULL) {
        fprintf(stderr, "%s\n", jgpu_last_error());
        return EXIT_FAILURE;
      }
    }
    file.data = ctx->buf;
    file.size = ctx->size;
    if (jgpu_decode_jpegs_ex(ctx->gpu, &file, 1, 1, JGPU_ENTROPY_GPU, img->pixels, cap, &jinfo) != EXIT_SUCCESS) {
      fprintf(stderr, "%s\n", jgpu_last_error());
      return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
  }
  if (ctx->entropy_on_device && out == JPEG_DECODE_YUV) {
    /* planes come back packed Y | Cb | Cr; the surface holds them as separate allocations */
    jgpu_jpeg file;
    jgpu_jpeg_info jinfo;
    jgpu_image_desc d;
    jgpu_layout lay;
    if (jgpu_desc_from_header(&ctx->header, &d) || jgpu_layout_query(&d, &lay)) {
      fprintf(stderr, "%s\n", jgpu_last_error());
      return EXIT_FAILURE;
    }
    for (i = 0; i < d.ncomps; i++) {
      if (img->plane[i].data == NULL || img->plane[i].width != lay.plane[i].width ||
          img->plane[i].height != lay.plane[i].height) {
        fprintf(stderr, "Error, image surface does not match the jpeg header\n");
        return EXIT_FAILURE;
      }
    }
    if (ctx->gpu == NULL) {
      ctx->gpu = jgpu_create(ctx->device);
      if (ctx->gpu == NULL) {
        fprintf(stderr, "%s\n", jgpu_last_error());
        return EXIT_FAILURE;
      }
    }
    if (ctx->planes_cap < lay.data_len) {
      jgpu_host_free(ctx->planes);
      ctx->planes = (uint8_t *)jgpu_host_alloc((size_t)lay.data_len);
      ctx->planes_cap = ctx->planes ? lay.data_len : 0;
      if (ctx->planes == NULL) {
        fprintf(stderr, "%s\n", jgpu_last_error());
        return EXIT_FAILURE;
      }
    }
    file.data = ctx->buf;
    file.size = ctx->size;
    if (jgpu_decode_jpegs_ex(ctx->gpu, &file, 1, 1, JGPU_ENTROPY_GPU | JGPU_JPEGS_OUT_YUV, ctx->planes,
                             ctx->planes_cap, &jinfo) != EXIT_SUCCESS) {
      fprintf(stderr, "%s\n", jgpu_last_error());
      return EXIT_FAILURE;
    }
    for (i = 0; i < d.ncomps; i++) {
      memcpy(img->plane[i].data, ctx->planes + lay.plane[i].data_off,
             (size_t)lay.plane[i].width * lay.plane[i].height);
    }
    return EXIT_SUCCESS;
  }
  if (ctx->upload == JPEG_DECODE_PACK) {
    /* the reader counts words into plane[i].packed (src/xjpeg.c:492,515,532); start from zero */
    for (i = 0; i < img->nplanes && i < NPLANES_MAX; i++) img->plane[i].packed = 0;
  }
  if ((*ctx->front.decode_image)(ctx->front_ctx, img, ctx->upload) != EXIT_SUCCESS) {
    return EXIT_FAILURE;
  }
  if (ctx->gpu == NULL) {
    ctx->gpu = jgpu_create(ctx->device);
    if (ctx->gpu == NULL) {
      fprintf(stderr, "%s\n", jgpu_last_error());
      return EXIT_FAILURE;
    }
  }
  if (ctx->upload == JPEG_DECODE_PACK) {
    for (i = 0; i < img->nplanes && i < NPLANES_MAX; i++) words += img->plane[i].packed;
    img->packed = (int)words; /* as the reference's caller does, src/jpeg_gpu.c:766-770 */
    rc = jgpu_decode_image_packed(ctx->gpu, &ctx->header, img, words, out);
  } else {
    rc = jgpu_decode_image(ctx->gpu, &ctx->header, img, out);
  }
  if (rc != EXIT_SUCCESS) {
    fprintf(stderr, "%s\n", jgpu_last_error());
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}

static void cuda_decode_reset(cuda_decode_ctx *ctx, jpeg_info *info) {
  /* device buffers, streams and the cached plan survive; only the parser restarts */
  (*ctx->front.decode_reset)(ctx->front_ctx, info);
  ctx->buf = info->buf;
  ctx->size = info->size;
  ctx->have_header = 0;
}

static void cuda_decode_free(cuda_decode_ctx *ctx) {
  if (ctx == NULL) return;
  (*ctx->front.decode_free)(ctx->front_ctx);
  jgpu_host_free(ctx->planes);
  jgpu_destroy(ctx->gpu);
  free(ctx);
}

const jpeg_decode_ctx_vtbl CUDA_DECODE_CTX_VTBL = {
    (jpeg_decode_alloc_func)cuda_decode_alloc,
    (jpeg_decode_header_func)cuda_decode_header,
    (jpeg_decode_image_func)cuda_decode_image,
    (jpeg_decode_reset_func)cuda_decode_reset,
    (jpeg_decode_free_func)cuda_decode_free};
