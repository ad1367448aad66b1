/* jpeg_gpu_b200.h — C ABI of the B200 JPEG block-decode back end.
 *
 * The library replaces the GPU half of negge/jpeg_gpu — the fragment-shader
 * passes res/{horz,vert,horz_quant_yuv,horz_quant_grey,unyuv,ungrey,rgb}.fs.glsl
 * driven from src/jpeg_gpu.c:1320-1363 — with one fused sm_100a CUDA kernel:
 *
 *     quantised int16 coefficient planes (JPEG_DECODE_QUANT layout,
 *     src/xjpeg.c:550-563, src/image.c:66-95)  +  DQT tables (natural order)
 *         -> dequantise -> 8x8 inverse DCT (bit-exact with src/dct.c)
 *         -> +128 / clamp (Y, Cb, Cr planes, bit-exact with src/xjpeg.c:565-584)
 *         -> nearest-neighbour chroma upsample -> YCbCr->RGB (res/yuv.fs.glsl)
 *         -> interleaved RGB8 in the layout of image.pixels (src/image.h:50)
 *
 * Two surfaces are exported:
 *
 *  (1) the reference's own plugin boundary: CUDA_DECODE_CTX_VTBL is a third
 *      `jpeg_decode_ctx_vtbl` (src/jpeg_wrap.h:45-54) with the same five slots,
 *      call protocol, ownership and EXIT_SUCCESS/EXIT_FAILURE + stderr error
 *      convention as LIBJPEG_DECODE_CTX_VTBL / XJPEG_DECODE_CTX_VTBL
 *      (src/jpeg_wrap.c:246-252,352-358);
 *  (2) a batch API over device-resident (or host) buffers, which the vtable
 *      path is a one-image wrapper around.
 *
 * Everything is plain C: pointers, sizes, ints.  Functions returning int
 * return 0 (EXIT_SUCCESS) or 1 (EXIT_FAILURE); jgpu_last_error() holds the
 * message of the last failure on the calling thread.  No call aborts.
 * A jgpu_ctx / jgpu_plan must be used by one host thread at a time.
 */
#ifndef JPEG_GPU_B200_H
#define JPEG_GPU_B200_H

#include <stddef.h>
#include <stdint.h>
#include "jgpu_ref_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

#define JGPU_VERSION 100

/* ------------------------------------------------------------------------
 * (1) decoder-plugin boundary
 * --------------------------------------------------------------------- */

/* Third backend.  decode_image(dec, img, out):
 *   PACK / QUANT / DCT  forwarded to the CPU front end (no GPU work);
 *   YUV   front end -> QUANT planes -> GPU -> img->plane[i].data
 *         (bit-exact with XJPEG_DECODE_CTX_VTBL's YUV output);
 *   RGB   front end -> QUANT planes -> GPU -> img->pixels, width*height*3
 *         bytes R,G,B (width*height bytes for 1-component files, the
 *         convention of the libjpeg backend, src/jpeg_wrap.c:215-220).
 * Replaces: the per-frame GL sequence src/jpeg_gpu.c:1320-1363. */
extern const jpeg_decode_ctx_vtbl CUDA_DECODE_CTX_VTBL;

/* Our own CPU entropy front end (baseline sequential Huffman, DRI/RSTn).  It
 * supports PACK/QUANT/DCT and produces buffers identical to
 * XJPEG_DECODE_CTX_VTBL's (src/xjpeg.c:449-632); YUV/RGB return
 * EXIT_FAILURE exactly like the xjpeg backend does for RGB
 * (src/jpeg_wrap.c:335-339).  It is the default producer behind
 * CUDA_DECODE_CTX_VTBL when the library is used outside the reference tree. */
extern const jpeg_decode_ctx_vtbl JFRONT_DECODE_CTX_VTBL;

/* Options of a CUDA decoder context.  The reference's five-slot table has no room for them
 * (decode_alloc takes the file only, src/jpeg_wrap.h:35), so the slot reads the CALLING THREAD's
 * current options -- set with the cuda_decode_set_* functions below, each thread its own copy, no
 * process-wide state -- and the context keeps what it read for its lifetime.  Callers outside the
 * table's shape can pass them explicitly: cuda_decode_alloc_ex. */
typedef struct cuda_decode_options {
  const jpeg_decode_ctx_vtbl *frontend; /* NULL: JFRONT_DECODE_CTX_VTBL */
  int device;                           /* CUDA ordinal; < 0: $JGPU_DEVICE, else 0 */
  jpeg_decode_out upload;               /* JPEG_DECODE_QUANT or JPEG_DECODE_PACK */
  int entropy_on_device;                /* 0 / 1; < 0: $JGPU_ENTROPY=gpu turns it on */
} cuda_decode_options;
/* Fills *opt with the calling thread's current options. */
void cuda_decode_get_options(cuda_decode_options *opt);
/* decode_alloc with explicit options (the other four slots of CUDA_DECODE_CTX_VTBL apply to the
 * result).  NULL on a bad option (message in jgpu_last_error()) or out of memory. */
jpeg_decode_ctx *cuda_decode_alloc_ex(jpeg_info *info, const cuda_decode_options *opt);

/* Inside the reference tree: make CUDA_DECODE_CTX_VTBL pull its coefficient
 * planes from the reference's own reader, e.g.
 *     cuda_decode_set_frontend(&XJPEG_DECODE_CTX_VTBL);
 * NULL restores JFRONT_DECODE_CTX_VTBL.  Affects contexts this thread allocates later. */
void cuda_decode_set_frontend(const jpeg_decode_ctx_vtbl *frontend);
/* CUDA device ordinal used by contexts this thread allocates later (default 0, or
 * $JGPU_DEVICE). */
void cuda_decode_set_device(int device);

/* What the backend pulls from its front end and sends to the device for YUV/RGB output:
 * JPEG_DECODE_QUANT (default; dense planes, 128 bytes per block) or JPEG_DECODE_PACK (the
 * zero-run packed stream, ~20 bytes per block, expanded on the device) -- the choice the
 * reference offers with `-o quant` / `-o pack` (src/jpeg_gpu.c:556-566,759-830).  Affects
 * contexts allocated later.  Returns EXIT_FAILURE for any other value. */
int cuda_decode_set_upload(jpeg_decode_out format);

/* Where decode_image(..., JPEG_DECODE_RGB / JPEG_DECODE_YUV) does its Huffman decoding: 0 (default) on the host, in
 * the front end; 1 on the device (jgpu_huff.cu, see jgpu_decode_jpegs_ex): the file's bytes are
 * uploaded as they are and no coefficient ever exists on the host.  Applies to the built-in
 * front end only (a front end set with cuda_decode_set_frontend keeps decoding on the host) and
 * to contexts allocated later.  Never called: $JGPU_ENTROPY=gpu turns it on.  Well-formed files
 * give the same bytes either way; a damaged file the device decoder hands back is decoded as
 * jgpu_decode_jpegs_ex(JGPU_ENTROPY_CPU) would (restart intervals at their markers), which accepts
 * some files the strictly sequential front end rejects.  Returns EXIT_FAILURE for any other value. */
int cuda_decode_set_entropy(int on_device);

/* Output-surface helpers with the semantics of the reference's
 * image_init / image_zero / image_clear (src/image.c:24-123) and
 * jpeg_info_init / jpeg_info_clear (src/jpeg_info.c:31-61), for callers that
 * do not link the reference objects.  Buffers are 16-byte aligned. */
int jgpu_image_init(image *img, jpeg_header *header);
/* JGPU_IMAGE_PINNED: page-locked `pixels` (cudaHostAlloc; ordinary memory when that fails), which
 * the backend reads back into a millisecond faster per 4K frame.  jgpu_image_clear frees either
 * kind (it asks the CUDA runtime what the pointer is). */
#define JGPU_IMAGE_PINNED 1u
int jgpu_image_init_ex(image *img, jpeg_header *header, unsigned flags);
/* For callers that keep the two-argument shape: on != 0 makes the surfaces THIS THREAD initialises
 * afterwards with jgpu_image_init page-locked.  Off by default. */
void jgpu_image_set_pinned(int on);
/* 1 when p points into page-locked host memory known to the CUDA runtime, else 0. */
int jgpu_host_is_pinned(const void *p);
void jgpu_image_zero(image *img);
void jgpu_image_clear(image *img);
int jgpu_info_init(jpeg_info *info, const char *name);
void jgpu_info_clear(jpeg_info *info);

/* ------------------------------------------------------------------------
 * (2) batch API
 * --------------------------------------------------------------------- */

typedef struct jgpu_ctx jgpu_ctx;
typedef struct jgpu_plan jgpu_plan;

/* One image of a batch.  Geometry is derived exactly as the reference does
 * (nhmb/nvmb src/xjpeg.c:403-407, hblocks/vblocks src/jpeg_wrap.c:305-306,
 * plane size / xdec / ydec / cstride / coefficient offsets src/image.c:38-95).
 * Requires hsamp[0] == max(hsamp) (the reference's coefficient layout is only
 * self-consistent in that case, src/xjpeg.c:558-560). */
typedef struct jgpu_image_desc {
  int32_t width, height;      /* visible size, 1..65535 */
  int32_t ncomps;             /* 1 or 3 */
  int32_t hsamp[3], vsamp[3]; /* 1, 2 or 4 */
  int32_t tq[3];              /* DQT slot 0..3 of each component */
  int32_t qtab_set;           /* index of this image's 4x64 table set */
  int32_t reserved;
  int64_t coef_off;           /* first int16 of this image in the coef buffer
                                 (multiple of 8); layout = image.coef */
  int64_t rgb_off;            /* first byte of this image in the rgb buffer */
  int64_t yuv_off;            /* first byte of the padded Y|Cb|Cr planes in the
                                 yuv buffer, or -1 */
} jgpu_image_desc;

/* What the reference's image_init would allocate for this image. */
typedef struct jgpu_plane_layout {
  int32_t hblocks, vblocks;   /* coded blocks */
  int32_t width, height;      /* padded plane size */
  int32_t xdec, ydec, cstride, reserved;
  int64_t coef_off;           /* plane.coef - image.coef, in int16 */
  int64_t data_off;           /* offset in a packed Y|Cb|Cr buffer, bytes */
} jgpu_plane_layout;

typedef struct jgpu_layout {
  int32_t nhmb, nvmb, hmax, vmax;
  int64_t coef_len;           /* int16 elements incl. the reference's padding */
  int64_t coded_blocks;       /* sum hblocks*vblocks */
  int64_t data_len;           /* bytes of the three padded planes */
  int64_t rgb_len;            /* width*height*(ncomps==1 ? 1 : 3) */
  jgpu_plane_layout plane[3];
} jgpu_layout;

/* Host-only; needs no GPU. */
int jgpu_layout_query(const jgpu_image_desc *desc, jgpu_layout *out);

const char *jgpu_last_error(void);
int jgpu_device_count(void);

jgpu_ctx *jgpu_create(int device);
void jgpu_destroy(jgpu_ctx *ctx);

#define JGPU_OUT_RGB 1u          /* write interleaved RGB8 / grey8 */
#define JGPU_OUT_YUV 2u          /* write the padded u8 Y|Cb|Cr planes */
#define JGPU_FORCE_GENERIC 4u    /* use the two-kernel generic path (tests) */

/* A plan holds the device-side work lists for one batch shape; running it
 * launches kernels only (no allocation, no host<->device copies). */
jgpu_plan *jgpu_plan_create(jgpu_ctx *ctx, const jgpu_image_desc *descs, int n,
                            unsigned flags);
void jgpu_plan_destroy(jgpu_plan *plan);
/* Kernels one jgpu_plan_run launches. */
int jgpu_plan_launches(const jgpu_plan *plan);
/* Total algorithmic bytes (128 B per coded block + output bytes) of one run. */
int64_t jgpu_plan_bytes(const jgpu_plan *plan);

/* All pointers are DEVICE pointers; d_qtabs is n_sets x 4 x 64 uint16 in
 * natural order (jpeg_quant.tbl); stream is a cudaStream_t (NULL = default).
 * Asynchronous: returns after enqueueing. */
int jgpu_plan_run(jgpu_plan *plan, const int16_t *d_coef,
                  const uint16_t *d_qtabs, int n_sets, uint8_t *d_rgb,
                  uint8_t *d_yuv, void *stream);

/* HOST buffers in, HOST buffers out: pinned staging, chunked H2D / kernel /
 * D2H overlapped on internal streams; synchronous.  h_rgb / h_yuv may be NULL
 * according to flags. */
int jgpu_decode_batch_host(jgpu_ctx *ctx, const jgpu_image_desc *descs, int n,
                           unsigned flags, const int16_t *h_coef,
                           const uint16_t *h_qtabs, int n_sets, uint8_t *h_rgb,
                           uint8_t *h_yuv);

/* One image on the reference's own structs: img->coef holds QUANT planes
 * (from any backend that honours src/xjpeg.c:550-563), header holds the
 * tables.  out = JPEG_DECODE_YUV fills img->plane[i].data, JPEG_DECODE_RGB
 * fills img->pixels.  Synchronous. */
int jgpu_decode_image(jgpu_ctx *ctx, const jpeg_header *header, image *img,
                      jpeg_decode_out out);

/* ------------------------------------------------------------------------
 * (3) PACK input: the reference's zero-run packed coefficient stream
 * --------------------------------------------------------------------- */

/* JPEG_DECODE_PACK as the reference's reader writes it (src/xjpeg.c:484-496,
 * 513-519,531-535) and its GL path consumes it (res/horz_pack_yuv.fs.glsl:
 * 105-127, upload at src/jpeg_gpu.c:1286-1287):
 *   pack[]   16-bit words in scan (MCU-interleaved) order; per block a DC word
 *            `dc & 0xfff`, then one word `run << 12 | value & 0xfff` per coded AC
 *            symbol (ZRL = 0xf000), then an end-of-block word 0 unless the block
 *            ran to coefficient 63;
 *   index[]  per block the position of its DC word, laid out like image.index
 *            (src/image.c:85-95): entry k belongs to the block whose dense
 *            coefficients start at image.coef + 64*k.
 * A batch holds image i's words at pack[pack_off[i] .. pack_off[i+1]) (index
 * values are relative to pack_off[i]) and its index entries at
 * index[coef_off/64 ...]; coef_off must be a multiple of 64.  Typical blocks
 * take ~20 bytes instead of 128, which is what the host->device link carries. */

/* Worst case words one image can need: 64 per coded block. */
int64_t jgpu_pack_bound(const jgpu_image_desc *desc);

/* Host utility (tests, benches, callers that hold dense planes): QUANT planes
 * of ONE image (coef = that image's image.coef) -> the PACK words and index the
 * reference's reader produces for a baseline scan carrying the same
 * coefficients.  Values are truncated to 12 bits exactly as the reference
 * does.  index must have coef_len/64 entries (unused ones are set to 0).
 * Returns the number of words written, or -1 (pack_cap too small / bad desc). */
int64_t jgpu_pack_from_quant(const jgpu_image_desc *desc, const int16_t *coef,
                             uint16_t *pack, int64_t pack_cap, int32_t *index);

/* DEVICE pointers.  Expands the PACK stream of every image of the plan into
 * dense QUANT planes at d_coef (the layout jgpu_plan_run reads), i.e. what
 * res/horz_pack_*.fs.glsl does per fragment before its row transform.
 * d_pack_off: n+1 int64 on the device.  Blocks whose words are cut off by
 * pack_off[i+1] or overrun coefficient 63 are truncated, never read out of
 * bounds.  Asynchronous. */
int jgpu_plan_unpack(jgpu_plan *plan, const uint16_t *d_pack,
                     const int64_t *d_pack_off, const int32_t *d_index,
                     int16_t *d_coef, void *stream);

/* HOST buffers: PACK words + index in, RGB / YUV out (jgpu_decode_batch_host
 * with the packed stream on the link instead of the dense planes).
 * pack_off: n+1 host int64. */
int jgpu_decode_batch_host_packed(jgpu_ctx *ctx, const jgpu_image_desc *descs,
                                  int n, unsigned flags, const uint16_t *h_pack,
                                  const int64_t *pack_off, const int32_t *h_index,
                                  const uint16_t *h_qtabs, int n_sets,
                                  uint8_t *h_rgb, uint8_t *h_yuv);

/* ------------------------------------------------------------------------
 * (4) JPEG files in, RGB out: multi-threaded entropy front end + GPU back end
 * --------------------------------------------------------------------- */

/* The reference decodes one file per call on one thread (src/jpeg_gpu.c:1228-1237) and its
 * Huffman reader is >99 % of the time once the back half runs on the GPU.
 * This entry point runs JFRONT_DECODE_CTX_VTBL's scan decoder on `nthreads` host threads --
 * one task per image, and per group of restart intervals for files that carry DRI (RSTn
 * markers make the pieces independent, src/xjpeg.c:593-629) -- writing QUANT planes straight
 * into pinned staging, and overlaps it with the upload, the fused kernel and the read-back of
 * the images already finished. */
typedef struct jgpu_jpeg {
  const unsigned char *data;  /* a whole baseline JPEG file in memory (caller-owned) */
  int64_t size;
} jgpu_jpeg;

typedef struct jgpu_jpeg_info {
  int32_t status;             /* 0 = ok, 1 = rejected (see message) */
  int32_t width, height, ncomps;
  int32_t hsamp0, vsamp0;     /* luma sampling factors */
  int32_t restart_interval;
  int32_t tasks;              /* entropy-decode tasks the image was split into */
  int64_t rgb_off, rgb_len;   /* where its pixels go in the output buffer */
  const char *message;        /* static string, NULL when ok */
} jgpu_jpeg_info;

/* Host only.  Parses the headers, assigns rgb_off back to back (256-byte aligned) and
 * returns the output buffer size in bytes (rejected files take no space), or -1. */
int64_t jgpu_jpegs_probe(const jgpu_jpeg *files, int n, jgpu_jpeg_info *info);
/* The same for the output selected by flags (0 or JGPU_JPEGS_OUT_YUV; other bits are ignored). */
int64_t jgpu_jpegs_probe_ex(const jgpu_jpeg *files, int n, unsigned flags, jgpu_jpeg_info *info);

/* Decodes every accepted file into h_rgb + info[i].rgb_off (interleaved RGB8, or grey8 for
 * 1-component files).  h_rgb may be pageable or pinned.  nthreads <= 0: all host cores.
 * info is filled as by jgpu_jpegs_probe.  Returns EXIT_SUCCESS when every file decoded;
 * EXIT_FAILURE otherwise (info[i].status / message say which and why; the others are still
 * decoded). */
int jgpu_decode_jpegs(jgpu_ctx *ctx, const jgpu_jpeg *files, int n, int nthreads,
                      uint8_t *h_rgb, int64_t rgb_cap, jgpu_jpeg_info *info);

/* The same with the entropy decoder chosen explicitly:
 *   JGPU_ENTROPY_GPU   Huffman decoding on the GPU (jgpu_huff.cu: self-synchronising parallel
 *                      decoding of 1024-bit subsequences, replaces the scan loop of
 *                      src/xjpeg.c:449-632).  Host threads only strip the byte stuffing and cut
 *                      the scan at its restart markers; the compressed scan is what crosses
 *                      the link.  Every file is verified on the device; a file the decoder does
 *                      not take or flags (corrupt, truncated) goes through the sequential
 *                      reader, so results and error reports are those of JGPU_ENTROPY_CPU.
 *                      info[i].tasks = 1024-bit subsequences decoded in parallel (1 = the file
 *                      went through the sequential reader).
 *   JGPU_ENTROPY_CPU   the host-thread reader described above.
 *   JGPU_ENTROPY_AUTO  $JGPU_ENTROPY ("cpu" / "gpu"), default gpu; what jgpu_decode_jpegs uses. */
#define JGPU_ENTROPY_AUTO 0u
#define JGPU_ENTROPY_CPU 1u
#define JGPU_ENTROPY_GPU 2u
/* OR-ed into flags: h_rgb is DEVICE memory (256-byte aligned) and the pixels stay there -- for
 * callers whose next stage runs on the GPU.  Implies the GPU entropy decoder; the call still
 * returns only when the pixels are complete.  The call is then as long as the chain of kernels,
 * so nthreads <= 0 means HALF the host cores here (the thread that feeds the GPU must not be
 * starved by the unstuffing workers), and the block decoder runs once, after the last group of
 * files, instead of once per group.  (One process per GPU on a shared host: pass nthreads
 * explicitly, cores / processes, or half of that here -- the default knows of one process only.) */
#define JGPU_JPEGS_DEVICE_OUT 0x100u
/* OR-ed into flags: the output is what the reference's xjpeg backend produces for JPEG_DECODE_YUV
 * (src/xjpeg.c:565-584) instead of pixels -- per file the padded u8 planes Y | Cb | Cr back to
 * back (jgpu_layout.plane[i].data_off / width / height, jgpu_layout.data_len bytes), bit-exact
 * with it, at info[i].rgb_off with info[i].rgb_len = data_len.  Half the read-back of RGB for
 * 4:2:0.  Implies the GPU entropy decoder.  jgpu_jpegs_probe_ex sizes the buffer. */
#define JGPU_JPEGS_OUT_YUV 0x200u
int jgpu_decode_jpegs_ex(jgpu_ctx *ctx, const jgpu_jpeg *files, int n, int nthreads, unsigned flags,
                         uint8_t *h_rgb, int64_t rgb_cap, jgpu_jpeg_info *info);

/* One image on the reference's structs with the PACK stream in img->coef / img->index, as a
 * reader leaves them after decode_image(..., JPEG_DECODE_PACK); words = sum of
 * img->plane[i].packed.  out = JPEG_DECODE_YUV or JPEG_DECODE_RGB.  Synchronous. */
int jgpu_decode_image_packed(jgpu_ctx *ctx, const jpeg_header *header, image *img,
                             int64_t words, jpeg_decode_out out);

/* Page-locked host memory for the batch entry points. */
void *jgpu_host_alloc(size_t bytes);
void jgpu_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* JPEG_GPU_B200_H */
