/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Link-time stand-ins for the libjpeg entry points src/jpeg_wrap.c:54-252 calls
 * (see oracle/stub/jpeglib.h).  libjpeg is a third-party sibling backend that
 * is not present in this image and not on the coefficient->RGB path; whoever
 * selects LIBJPEG_DECODE_CTX_VTBL from oracle/_ref gets a message and abort().
 */
#include <stdio.h>
#include <stdlib.h>
#include <jpeglib.h>

static void absent(const char *what) {
  fprintf(stderr, "oracle/_ref: %s called, but libjpeg is not part of this build (stub)\n", what);
  abort();
}

struct jpeg_error_mgr *jpeg_std_error(struct jpeg_error_mgr *err) { absent("jpeg_std_error"); return err; }
void jpeg_create_decompress(j_decompress_ptr cinfo) { (void)cinfo; absent("jpeg_create_decompress"); }
void jpeg_destroy_decompress(j_decompress_ptr cinfo) { (void)cinfo; absent("jpeg_destroy_decompress"); }
void jpeg_mem_src(j_decompress_ptr cinfo, const unsigned char *b, unsigned long n) {
  (void)cinfo; (void)b; (void)n; absent("jpeg_mem_src");
}
int jpeg_read_header(j_decompress_ptr cinfo, boolean r) { (void)cinfo; (void)r; absent("jpeg_read_header"); return 0; }
jvirt_barray_ptr *jpeg_read_coefficients(j_decompress_ptr cinfo) {
  (void)cinfo; absent("jpeg_read_coefficients"); return NULL;
}
boolean jpeg_start_decompress(j_decompress_ptr cinfo) { (void)cinfo; absent("jpeg_start_decompress"); return 0; }
JDIMENSION jpeg_read_scanlines(j_decompress_ptr cinfo, JSAMPARRAY s, JDIMENSION m) {
  (void)cinfo; (void)s; (void)m; absent("jpeg_read_scanlines"); return 0;
}
JDIMENSION jpeg_read_raw_data(j_decompress_ptr cinfo, JSAMPIMAGE d, JDIMENSION m) {
  (void)cinfo; (void)d; (void)m; absent("jpeg_read_raw_data"); return 0;
}
boolean jpeg_finish_decompress(j_decompress_ptr cinfo) { (void)cinfo; absent("jpeg_finish_decompress"); return 0; }
