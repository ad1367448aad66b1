/* jgpu_front.h — internal: the CPU entropy front end's entry points for the batch decoder
 * (jgpu_jpegs.cu).  `ctx` is a context of JFRONT_DECODE_CTX_VTBL after decode_header. */
#ifndef JGPU_FRONT_H
#define JGPU_FRONT_H

#include "jpeg_gpu_b200.h"
#include "jgpu_huff_core.h"

#ifdef __cplusplus
extern "C" {
#endif

/* The restart intervals of a scan: independently decodable pieces (T.81 F.1.1.5, E.2.4; the
 * reference walks them in sequence, src/xjpeg.c:593-629). */
typedef struct jfront_segments {
  int nseg;          /* 1 when the file has no DRI */
  int mcus_per_seg;  /* the restart interval, or all MCUs */
  int total_mcus;
  int *seg_pos;      /* byte offset of each piece's first entropy-coded byte */
} jfront_segments;

int jfront_find_segments(jpeg_decode_ctx *ctx, jfront_segments *out);
void jfront_segments_free(jfront_segments *s);
int jfront_decode_segments(const jpeg_decode_ctx *ctx, image *img, jpeg_decode_out out,
                           const jfront_segments *segs, int s0, int s1, const char **error);

/* Preparation for the GPU entropy decoder (jgpu_huff.cu); see jgpu_front.c. */
long long jfront_huff_prepare(jpeg_decode_ctx *ctx, int subseq_words, unsigned char *stream,
                              long long stream_cap, unsigned int *seg_first, int seg_cap,
                              jgpu_huff_table *tables, jgpu_huff_file *file, const char **why);
long long jfront_huff_bound(const jpeg_decode_ctx *ctx, int subseq_words, int *nseg_out);

#ifdef __cplusplus
}
#endif
#endif
