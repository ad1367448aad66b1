/* jgpu_ref_abi.h — the reference's decoder-plugin interface, restated.
 *
 * The CUDA backend is a drop-in third backend next to the reference's
 * LIBJPEG_DECODE_CTX_VTBL and XJPEG_DECODE_CTX_VTBL.  When it is compiled
 * INSIDE the reference tree the reference's own headers are used (define
 * JGPU_USE_REFERENCE_HEADERS and put src/ on the include path).  Standalone,
 * this header declares the same types with the same names, member order and
 * enum values so that objects are link- and layout-compatible:
 *
 *   jpeg_subsamp, jpeg_quant, jpeg_component,
 *   jpeg_header, jpeg_info                     <- src/jpeg_info.h:22-76
 *   image_plane, image                         <- src/image.h:21-55
 *   jpeg_decode_out, the five function
 *   typedefs, jpeg_decode_ctx_vtbl             <- src/jpeg_wrap.h:22-54
 *
 * tests/test_abi.py compiles one probe against this header and one
 * against the reference's headers and compares sizeof/offsetof of every
 * member.  The include guards below are deliberately the reference's own, so
 * that including both sets in one translation unit cannot redeclare a type.
 */
#ifndef JGPU_REF_ABI_H
#define JGPU_REF_ABI_H

#ifdef JGPU_USE_REFERENCE_HEADERS
#include "jpeg_wrap.h" /* pulls image.h and jpeg_info.h */
#else

/* ---- src/jpeg_info.h ---------------------------------------------------- */
#if !defined(_jpeg_info_H)
#define _jpeg_info_H (1)

#define NCOMPS_MAX (3)
#define NQUANT_MAX (4)

typedef enum jpeg_subsamp {
  JPEG_SUBSAMP_UNKNOWN, /* 0 */
  JPEG_SUBSAMP_444,     /* 1 */
  JPEG_SUBSAMP_422,     /* 2 */
  JPEG_SUBSAMP_420,     /* 3 */
  JPEG_SUBSAMP_440,     /* 4 */
  JPEG_SUBSAMP_411,     /* 5 */
  JPEG_SUBSAMP_MONO,    /* 6 */
  JPEG_SUBSAMP_MAX
} jpeg_subsamp;

typedef struct jpeg_quant jpeg_quant;
typedef struct jpeg_component jpeg_component;
typedef struct jpeg_header jpeg_header;
typedef struct jpeg_info jpeg_info;

/* One DQT table.  tbl[] is in NATURAL (row-major) order: the reference stores
 * tbl[DE_ZIG_ZAG[i]] while parsing (src/xjpeg.c:238-247). */
struct jpeg_quant {
  int valid;
  unsigned char bits;
  unsigned short tbl[64];
};

struct jpeg_component {
  int hblocks; /* nhmb*hsamp */
  int vblocks; /* nvmb*vsamp */
  int hsamp;
  int vsamp;
  jpeg_quant *quant; /* points into the owning header's quant[] */
};

struct jpeg_header {
  int bits;
  int width;
  int height;
  int ncomps;
  jpeg_subsamp subsamp;
  int restart_interval;
  jpeg_component comp[NCOMPS_MAX];
  jpeg_quant quant[NQUANT_MAX];
};

/* An in-memory JPEG file; owned by the caller. */
struct jpeg_info {
  int size;
  unsigned char *buf;
};

#endif /* _jpeg_info_H */

/* ---- src/image.h -------------------------------------------------------- */
#if !defined(_image_H)
#define _image_H (1)

#define NPLANES_MAX (3)

typedef struct image_plane image_plane;
typedef struct image image;

struct image_plane {
  int bitdepth;
  unsigned char xdec;
  unsigned char ydec;
  int xstride;
  int ystride;
  unsigned short width;  /* MCU-padded, hblocks*8 */
  unsigned short height; /* MCU-padded, vblocks*8 */
  unsigned char *data;   /* u8 samples, YUV output */
  short *coef;           /* this plane's slice of image.coef */
  int cstride;
  int packed;
  int *index;
};

struct image {
  unsigned short width;
  unsigned short height;
  int nplanes;
  image_plane plane[NPLANES_MAX];
  short *coef;           /* all planes, block-contiguous (src/xjpeg.c:550-563) */
  int packed;
  int *index;
  unsigned char *pixels; /* width*height*3 bytes, RGB output */
};

#endif /* _image_H */

/* ---- src/jpeg_wrap.h ---------------------------------------------------- */
#if !defined(_jpeg_wrap_H)
#define _jpeg_wrap_H (1)

typedef struct jpeg_decode_ctx jpeg_decode_ctx;

/* How far the backend decodes.  The numeric values are part of the ABI:
 * xjpeg casts them to its own enum (src/jpeg_wrap.c:328). */
typedef enum jpeg_decode_out {
  JPEG_DECODE_PACK,  /* 0: run/level stream + per-block index */
  JPEG_DECODE_QUANT, /* 1: quantised, de-zigzagged coefficient planes */
  JPEG_DECODE_DCT,   /* 2: dequantised coefficient planes */
  JPEG_DECODE_YUV,   /* 3: u8 Y/Cb/Cr planes */
  JPEG_DECODE_RGB,   /* 4: interleaved RGB8 */
  JPEG_DECODE_OUT_MAX
} jpeg_decode_out;

typedef jpeg_decode_ctx *(*jpeg_decode_alloc_func)(jpeg_info *info);
typedef int (*jpeg_decode_header_func)(jpeg_decode_ctx *dec,
                                       jpeg_header *header);
typedef int (*jpeg_decode_image_func)(jpeg_decode_ctx *dec, image *img,
                                      jpeg_decode_out out);
typedef void (*jpeg_decode_reset_func)(jpeg_decode_ctx *dec, jpeg_info *info);
typedef void (*jpeg_decode_free_func)(jpeg_decode_ctx *dec);

typedef struct jpeg_decode_ctx_vtbl jpeg_decode_ctx_vtbl;

struct jpeg_decode_ctx_vtbl {
  jpeg_decode_alloc_func decode_alloc;
  jpeg_decode_header_func decode_header;
  jpeg_decode_image_func decode_image;
  jpeg_decode_reset_func decode_reset;
  jpeg_decode_free_func decode_free;
};

#endif /* _jpeg_wrap_H */

#endif /* JGPU_USE_REFERENCE_HEADERS */
#endif /* JGPU_REF_ABI_H */
