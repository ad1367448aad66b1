/* Internal declarations shared by the C host side and the CUDA runtime. */
#ifndef JGPU_INTERNAL_H
#define JGPU_INTERNAL_H

#include "jpeg_gpu_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define JGPU_PRINTF(a, b) __attribute__((format(printf, a, b)))
#else
#define JGPU_PRINTF(a, b)
#endif

/* Records the message for jgpu_last_error() and returns 1 (EXIT_FAILURE). */
int jgpu_fail(const char *fmt, ...) JGPU_PRINTF(1, 2);

/* Fills a batch descriptor from the reference's header/image pair
 * (coefficients at img->coef, tables header->quant[]). */
int jgpu_desc_from_header(const jpeg_header *header, jgpu_image_desc *desc);

/* 16-byte aligned allocation used by the image surface helpers. */
void *jgpu_aligned_malloc(size_t bytes);
void jgpu_aligned_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
