/* jpeg_gpu_cli — the reference's command line (src/jpeg_gpu.c:473-506) without the window.
 *
 * Same options, same -H / -d text formats (src/jpeg_gpu.c:614-700), so a dump of this tool
 * diffs against a dump of the reference's binary; the frame loop
 * decode_reset -> decode_header -> decode_image (src/jpeg_gpu.c:1231-1237) runs headless for
 * --frames N iterations and reports the per-frame time once at the end instead of in a window
 * title (src/jpeg_gpu.c:1446-1458).  Backends: `cuda` (CUDA_DECODE_CTX_VTBL) and `jfront`
 * (our CPU entropy front end; accepts pack/quant/dct like the reference's `xjpeg`).
 *
 * Plain C against include/jpeg_gpu_b200.h only: this is also the smallest example of a host
 * program using the plugin boundary.
 */
#define _POSIX_C_SOURCE 200809L
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "jpeg_gpu_b200.h"

#define NAME "jpeg_gpu_cli"

static const char *const SUBSAMP_NAMES[JPEG_SUBSAMP_MAX] = {
    "Unknown", "4:4:4", "4:2:2", "4:2:0", "4:4:0", "4:1:1", "Mono"};
static const char *const OUT_NAMES[JPEG_DECODE_OUT_MAX] = {"pack", "quant", "dct", "yuv", "rgb"};

static const struct option OPTIONS[] = {
    {"help", no_argument, NULL, 'h'},     {"impl", required_argument, NULL, 'i'},
    {"out", required_argument, NULL, 'o'}, {"dump", no_argument, NULL, 'd'},
    {"header", no_argument, NULL, 'H'},   {"frames", required_argument, NULL, 'n'},
    {"entropy", required_argument, NULL, 'e'}, {NULL, 0, NULL, 0}};

static void usage(void) {
  fprintf(stderr,
          "Usage: %s [options] jpeg_file\n\n"
          "Options:\n\n"
          "  -h --help                      Display this help and exit.\n"
          "  -i --impl <decoder>            Decoder backend to use.\n"
          "                                 cuda (default) => CPU entropy decode, B200 back half\n"
          "                                 jfront => CPU entropy decode only (pack, quant, dct)\n"
          "  -o --out <format>              Format the decoder should output.\n"
          "                                 pack => RLC zero packed and quantized.\n"
          "                                 quant => quantized but de-zigzaged.\n"
          "                                 dct => DCT (12-bit dequantized)\n"
          "                                 yuv => YUV (4:4:4 or 4:2:0)\n"
          "                                 rgb (default) => RGB (4:4:4)\n"
          "  -d --dump                      Dump jpeg data in the output format.\n"
          "  -H --header                    Print the jpeg header.\n"
          "  -n --frames <count>            Run the decode loop <count> times and report\n"
          "                                  the time per frame.\n"
          "  -e --entropy <where>           Where -i cuda -o rgb decodes the Huffman scan.\n"
          "                                 cpu (default) => in the front end, on the host\n"
          "                                 gpu => on the device; the file is uploaded as it is\n\n"
          " %s accepts only 8-bit non-hierarchical baseline JPEG files.\n\n",
          NAME, NAME);
}

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int main(int argc, char *argv[]) {
  jpeg_decode_ctx_vtbl vtbl = CUDA_DECODE_CTX_VTBL;
  jpeg_decode_out out = JPEG_DECODE_RGB;
  int dump = 0, head = 0, frames = 0, c;
  jpeg_info info;
  jpeg_header header;
  image img;
  jpeg_decode_ctx *dec;

  while ((c = getopt_long(argc, argv, "hi:o:dHn:e:", OPTIONS, NULL)) != EOF) {
    switch (c) {
      case 'i':
        if (strcmp("cuda", optarg) == 0) vtbl = CUDA_DECODE_CTX_VTBL;
        else if (strcmp("jfront", optarg) == 0) vtbl = JFRONT_DECODE_CTX_VTBL;
        else {
          fprintf(stderr, "Invalid decoder implementation: %s\n", optarg);
          usage();
          return EXIT_FAILURE;
        }
        break;
      case 'o': {
        int k;
        for (k = 0; k < JPEG_DECODE_OUT_MAX && strcmp(OUT_NAMES[k], optarg) != 0; k++) {}
        if (k == JPEG_DECODE_OUT_MAX) {
          fprintf(stderr, "Invalid decoder output format: %s\n", optarg);
          usage();
          return EXIT_FAILURE;
        }
        out = (jpeg_decode_out)k;
        break;
      }
      case 'd': dump = 1; break;
      case 'H': head = 1; break;
      case 'n': frames = atoi(optarg); break;
      case 'e':
        if (strcmp("cpu", optarg) != 0 && strcmp("gpu", optarg) != 0) {
          fprintf(stderr, "Invalid entropy decoder: %s\n", optarg);
          usage();
          return EXIT_FAILURE;
        }
        cuda_decode_set_entropy(strcmp("gpu", optarg) == 0);
        break;
      case 'h':
      default: usage(); return EXIT_FAILURE;
    }
  }
  info.buf = NULL;
  info.size = 0;
  for (; optind < argc; optind++) {
    if (jgpu_info_init(&info, argv[optind]) != EXIT_SUCCESS) return EXIT_FAILURE;
  }
  if (info.buf == NULL) {
    usage();
    return EXIT_FAILURE;
  }

  /* the CUDA backend reads pixels back into the surface: give it page-locked memory */
  jgpu_image_set_pinned(vtbl.decode_image == CUDA_DECODE_CTX_VTBL.decode_image && out == JPEG_DECODE_RGB);
  dec = (*vtbl.decode_alloc)(&info);
  if (dec == NULL) return EXIT_FAILURE;
  if ((*vtbl.decode_header)(dec, &header) != EXIT_SUCCESS) return EXIT_FAILURE;
  if (head) {
    int i, j;
    printf("Image Size         : %ix%i\n", header.width, header.height);
    printf("Bits Per Pixel     : %i\n", header.bits);
    printf("Components         : %i\n", header.ncomps);
    printf("Chroma Subsampling : %s\n", SUBSAMP_NAMES[header.subsamp]);
    printf("Minimum Coded Unit : ");
    for (i = 0; i < header.ncomps; i++) {
      printf("%s%ix%i", i > 0 ? " " : "", header.comp[i].hsamp, header.comp[i].vsamp);
    }
    printf("\n");
    printf("Restart Interval   : %i\n", header.restart_interval);
    for (i = 0; i < NQUANT_MAX; i++) {
      if (header.quant[i].valid) {
        printf("Quant Table %i Bits : %i\n", i, header.quant[i].bits);
        for (j = 1; j <= 64; j++) {
          printf("%4i%s", header.quant[i].tbl[j - 1], j & 0x7 ? "" : "\n");
        }
      }
    }
    return EXIT_SUCCESS;
  }
  if (jgpu_image_init(&img, &header) != EXIT_SUCCESS) {
    fprintf(stderr, "Error initializing image\n");
    return EXIT_FAILURE;
  }
  jgpu_image_zero(&img);
  if ((*vtbl.decode_image)(dec, &img, out) != EXIT_SUCCESS) return EXIT_FAILURE;

  if (dump) {
    int i, j, k;
    if (out == JPEG_DECODE_PACK) {
      img.packed = 0;
      for (i = 0; i < img.nplanes; i++) {
        printf("Plane %i Packed Data: %i\n", i, img.plane[i].packed);
        img.packed += img.plane[i].packed;
      }
      printf("Packed Data : %i\n", img.packed);
      return EXIT_SUCCESS;
    }
    for (i = 0; i < img.nplanes; i++) {
      image_plane *plane = &img.plane[i];
      printf("Plane %i\n", i);
      switch (out) {
        case JPEG_DECODE_QUANT:
        case JPEG_DECODE_DCT:
          for (k = 0; k < plane->height; k++) {
            for (j = 0; j < plane->width; j++) printf("%4i ", plane->coef[k * plane->width + j]);
            printf("\n");
          }
          break;
        case JPEG_DECODE_YUV:
          for (k = 0; k < plane->height; k++) {
            for (j = 0; j < plane->width; j++) printf("%4i ", plane->data[k * plane->width + j]);
            printf("\n");
          }
          break;
        case JPEG_DECODE_RGB:
          /* one channel per "plane"; rows are the image's (the reference indexes rows with the
           * padded plane width here, src/jpeg_gpu.c:683, which skews its dump for widths that
           * are not a multiple of the MCU; this tool uses the real row stride) */
          for (k = 0; k < img.height; k++) {
            for (j = 0; j < img.width; j++) {
              printf("%4i ", img.nplanes == 1 ? img.pixels[k * img.width + j]
                                                : img.pixels[(k * img.width + j) * 3 + i]);
            }
            printf("\n");
          }
          break;
        default:
          fprintf(stderr, "Unsupported output %i.\n", (int)out);
          return EXIT_FAILURE;
      }
      printf("\n");
    }
    return EXIT_SUCCESS;
  }

  if (frames > 0) {
    /* the reference's steady-state loop, src/jpeg_gpu.c:1231-1237 */
    double t0 = now_ms(), dt;
    int f;
    for (f = 0; f < frames; f++) {
      (*vtbl.decode_reset)(dec, &info);
      (*vtbl.decode_header)(dec, &header);
      if ((*vtbl.decode_image)(dec, &img, out) != EXIT_SUCCESS) return EXIT_FAILURE;
    }
    dt = now_ms() - t0;
    printf("%i frames, %0.3f ms per frame, %.1f Mpixels/s (%s, %ix%i %s)\n", frames, dt / frames,
           frames * (double)header.width * header.height / dt / 1e3, OUT_NAMES[out], header.width,
           header.height, SUBSAMP_NAMES[header.subsamp]);
  }
  (*vtbl.decode_free)(dec);
  jgpu_image_clear(&img);
  jgpu_info_clear(&info);
  return EXIT_SUCCESS;
}
