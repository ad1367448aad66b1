/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_idct.c header).
 *
 * CPU restatement of the reference's coefficient -> RGB block-decode path.
 * Every function cites the reference file:line it follows.  The product
 * library (jpeg_gpu_b200/csrc) never includes this header.
 */
#ifndef JGPU_ORACLE_H
#define JGPU_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define JGO_MAX_PLANES 3
#define JGO_MAX_QTABS 4

/* Per-plane geometry as derived by the reference's image_init
 * (src/image.c:24-97) from a header filled the way xjpeg_decode_header_ does
 * (src/jpeg_wrap.c:285-316, src/xjpeg.c:400-407). */
typedef struct jgo_plane {
  int hsamp, vsamp;
  int hblocks, vblocks;   /* nhmb*hsamp, nvmb*vsamp */
  int width, height;      /* MCU-padded plane size in samples */
  int xdec, ydec;         /* log2 decimation against the largest sampling */
  int cstride;            /* block rows of this plane "at luma width" */
  long long coef_off;     /* offset of the plane inside image.coef, in shorts */
  long long data_off;     /* offset of the plane inside a packed Y|Cb|Cr buffer */
} jgo_plane;

typedef struct jgo_geom {
  int width, height, ncomps;
  int hmax, vmax, nhmb, nvmb;
  jgo_plane plane[JGO_MAX_PLANES];
  long long coef_len;     /* shorts allocated for image.coef (incl. padding) */
  long long data_len;     /* bytes of all padded planes back to back */
  long long rgb_len;      /* width*height*(ncomps==1 ? 1 : 3) */
} jgo_geom;

/* dct.c restatement */
void jgo_idct8x8(short *x, int xstride, const short *y, int ystride);
void jgo_idct_constants(float out[12]);

/* geometry */
int jgo_geometry(int width, int height, int ncomps, const int *hsamp,
                 const int *vsamp, jgo_geom *g);
/* flat int64 view for ctypes: see oracle_pipeline.c */
int jgo_geometry_flat(int width, int height, int ncomps, const int *hsamp,
                      const int *vsamp, long long *out);

/* stages */
int jgo_coef_to_yuv(const jgo_geom *g, const short *coef,
                    const unsigned short *qtabs, const int *tq,
                    unsigned char *planes);
int jgo_yuv_to_rgb(const jgo_geom *g, const unsigned char *planes,
                   unsigned char *rgb);
void jgo_colour_offsets(int cb, int cr, int out[3]);

/* PACK format (oracle_pack.c) */
int jgo_unpack_image(const jgo_geom *g, const unsigned short *pack, long long pack_len,
                     const int *index, short *coef);
long long jgo_pack_image(const jgo_geom *g, const short *coef, unsigned short *pack,
                         long long pack_cap, int *index, int *packed);

/* the reference's OpenGL float path, emulated (oracle_glsl.c): a comparator, not the ground truth */
void jgo_glsl_scales2d(float out[64]);
int jgo_glsl_decode_image(const jgo_geom *g, const short *coef, const unsigned short *qtabs, const int *tq,
                          int floor_mode, unsigned char *rgb, int *samples);
int jgo_glsl_decode_image_flat(int width, int height, int ncomps, const int *hsamp, const int *vsamp,
                               const short *coef, const unsigned short *qtabs, const int *tq, int floor_mode,
                               unsigned char *rgb, int *samples);

/* whole path over a batch; the CPU baseline */
int jgo_decode_batch(int n, const long long *desc, const short *coef,
                     const unsigned short *qtabs, unsigned char *rgb,
                     unsigned char *planes, int nthreads);

/* helpers that operate on bare blocks */
void jgo_idct_blocks(short *blocks, long long nblocks);
void jgo_dequant_idct_blocks(short *blocks, long long nblocks,
                             const unsigned short *q);
void jgo_ieee1180_gen(unsigned *state, int low, int high, int sign,
                      long long nblocks, short *coef, short *ref);
const char *jgo_idct_name(void);

#ifdef __cplusplus
}
#endif
#endif
