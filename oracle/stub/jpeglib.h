/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Declaration-only stand-in for <jpeglib.h>, which this image does not have.
 * Its only purpose: let the reference's src/jpeg_wrap.c compile UNMODIFIED
 * (oracle/Makefile `ref`), so that the reference's own XJPEG_DECODE_CTX_VTBL
 * (src/jpeg_wrap.c:254-358) exists in oracle/_ref/libjgpu_ref.so and can be
 * plugged into the CUDA backend with cuda_decode_set_frontend().
 *
 * The names below are the public libjpeg API that src/jpeg_wrap.c:54-252
 * (the libjpeg backend, a sibling of the path and out of scope) mentions; only
 * the members that file touches are declared.  The functions are defined in
 * oracle/jpeglib_stub.c and abort: LIBJPEG_DECODE_CTX_VTBL links, but is
 * never usable here.  Nothing of libjpeg's implementation is restated.
 */
#ifndef JGPU_ORACLE_JPEGLIB_STUB_H
#define JGPU_ORACLE_JPEGLIB_STUB_H

#include <stddef.h>

typedef int boolean;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

#define NUM_QUANT_TBLS 4
#define DCTSIZE2 64
#define JPEG_HEADER_OK 1

typedef unsigned int JDIMENSION;
typedef unsigned char JSAMPLE;
typedef JSAMPLE *JSAMPROW;
typedef JSAMPROW *JSAMPARRAY;
typedef JSAMPARRAY *JSAMPIMAGE;
typedef short JCOEF;
typedef JCOEF JBLOCK[DCTSIZE2];
typedef JBLOCK *JBLOCKROW;
typedef JBLOCKROW *JBLOCKARRAY;
typedef struct jvirt_barray_control *jvirt_barray_ptr;

typedef enum { JDCT_ISLOW, JDCT_IFAST, JDCT_FLOAT } J_DCT_METHOD;

typedef struct { unsigned short quantval[DCTSIZE2]; boolean sent_table; } JQUANT_TBL;

typedef struct {
  int h_samp_factor;
  int v_samp_factor;
  int quant_tbl_no;
  JDIMENSION width_in_blocks;
  JDIMENSION height_in_blocks;
} jpeg_component_info;

struct jpeg_error_mgr { int msg_code; };

struct jpeg_common_struct;
typedef struct jpeg_common_struct *j_common_ptr;

struct jpeg_memory_mgr {
  JBLOCKARRAY (*access_virt_barray)(j_common_ptr cinfo, jvirt_barray_ptr ptr, JDIMENSION start_row,
                                    JDIMENSION num_rows, boolean writable);
};

struct jpeg_common_struct {
  struct jpeg_error_mgr *err;
  struct jpeg_memory_mgr *mem;
};

struct jpeg_decompress_struct {
  struct jpeg_error_mgr *err;
  struct jpeg_memory_mgr *mem;
  JDIMENSION image_width;
  JDIMENSION image_height;
  int num_components;
  int data_precision;
  unsigned int restart_interval;
  int max_h_samp_factor;
  int max_v_samp_factor;
  JQUANT_TBL *quant_tbl_ptrs[NUM_QUANT_TBLS];
  jpeg_component_info *comp_info;
  boolean raw_data_out;
  boolean do_fancy_upsampling;
  J_DCT_METHOD dct_method;
  JDIMENSION output_scanline;
  JDIMENSION output_height;
};
typedef struct jpeg_decompress_struct *j_decompress_ptr;

struct jpeg_error_mgr *jpeg_std_error(struct jpeg_error_mgr *err);
void jpeg_create_decompress(j_decompress_ptr cinfo);
void jpeg_destroy_decompress(j_decompress_ptr cinfo);
void jpeg_mem_src(j_decompress_ptr cinfo, const unsigned char *inbuffer, unsigned long insize);
int jpeg_read_header(j_decompress_ptr cinfo, boolean require_image);
jvirt_barray_ptr *jpeg_read_coefficients(j_decompress_ptr cinfo);
boolean jpeg_start_decompress(j_decompress_ptr cinfo);
JDIMENSION jpeg_read_scanlines(j_decompress_ptr cinfo, JSAMPARRAY scanlines, JDIMENSION max_lines);
JDIMENSION jpeg_read_raw_data(j_decompress_ptr cinfo, JSAMPIMAGE data, JDIMENSION max_lines);
boolean jpeg_finish_decompress(j_decompress_ptr cinfo);

#endif
