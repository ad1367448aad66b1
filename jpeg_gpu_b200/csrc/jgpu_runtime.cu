/* jgpu_runtime.cu — contexts, plans and the batch entry points of the C ABI
 * (include/jpeg_gpu_b200.h).  Host code only; the kernels live in
 * jgpu_kernels.cu / jgpu_fused.cu. */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>

#include "jgpu_huff.h"
#include "jgpu_internal.h"
#include "jgpu_launch.h"

using namespace jgpu;

#define CU_TRY(expr)                                                          \
  do {                                                                        \
    cudaError_t e_ = (expr);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      return jgpu_fail("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_),   \
                       __FILE__, __LINE__, cudaGetErrorString(e_));           \
    }                                                                         \
  } while (0)

namespace {

#ifndef JGPU_HOST_STREAMS
#define JGPU_HOST_STREAMS 3
#endif
constexpr int kHostStreams = JGPU_HOST_STREAMS;

/* A grow-only device or pinned-host buffer. */
struct Buffer {
  void *ptr = nullptr;
  size_t cap = 0;
  bool pinned_host = false;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    release();
    cudaError_t e = pinned_host ? cudaHostAlloc(&ptr, bytes, cudaHostAllocDefault)
                                : cudaMalloc(&ptr, bytes);
    if (e != cudaSuccess) {
      ptr = nullptr;
      return jgpu_fail("could not allocate %zu bytes of %s memory (%s)", bytes,
                       pinned_host ? "pinned host" : "device", cudaGetErrorString(e));
    }
    cap = bytes;
    return 0;
  }
  void release() {
    if (ptr) {
      if (pinned_host) cudaFreeHost(ptr); else cudaFree(ptr);
    }
    ptr = nullptr;
    cap = 0;
  }
};

template <typename T>
int upload(Buffer &buf, const std::vector<T> &v) {
  size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
  if (buf.reserve(bytes)) return 1;
  if (!v.empty()) {
    cudaError_t e = cudaMemcpy(buf.ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return jgpu_fail("descriptor upload failed (%s)", cudaGetErrorString(e));
  }
  return 0;
}

}  // namespace

struct jgpu_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t streams[kHostStreams] = {};
  cudaEvent_t events[kHostStreams] = {};
  /* jgpu_decode_jpegs_ex, entropy decoding on the device: the compressed scans go up on a stream of
   * their own, one event per group of files, so that a group's kernels never wait behind the copy
   * of a later group (created on first use) */
  cudaStream_t up_stream = nullptr;
  std::vector<cudaEvent_t> up_events;
  /* device mirrors used by the host-buffer entry points (grow-only) */
  Buffer d_coef, d_qtabs, d_rgb, d_yuv;
  Buffer d_pack, d_index, d_pack_off; /* jgpu_decode_batch_host_packed */
  /* pinned bounce buffers for pageable host memory */
  Buffer h_in[kHostStreams], h_out[kHostStreams];
  /* GPU entropy decoder (jgpu_decode_jpegs_ex): pinned staging and device arrays */
  Buffer hz_stream, hz_files, hz_tables, hz_segs, hz_status, hz_coef;
  Buffer dz_stream, dz_files, dz_tables, dz_segs, dz_status, dz_sub[4], dz_carry[2], dz_dc[kHostStreams];
  /* last plan built by jgpu_decode_batch_host, reused while descs match */
  jgpu_plan *cached_plan = nullptr;
  std::vector<jgpu_image_desc> cached_descs;
  unsigned cached_flags = 0;
};

struct jgpu_plan {
  jgpu_ctx *ctx = nullptr;
  unsigned flags = 0;
  int n = 0;
  std::vector<jgpu_image_desc> descs;
  std::vector<jgpu_layout> layouts;
  int64_t bytes = 0;
  int64_t scratch_len = 0;          /* planes scratch when the caller has no yuv */
  std::vector<int64_t> scratch_off; /* per image offset into the scratch */
  bool use_scratch = false;
  /* generic path */
  Buffer d_segs, d_pair_work, d_cimgs, d_colour_work, d_scratch;
  std::vector<int> img_first_pair_cta, img_first_colour_cta; /* size n+1 */
  /* fused path: jgpu_mcu.cu (pixels or planes), or its predecessor jgpu_fused.cu (pixels; JGPU_FUSED_IMPL=v9) */
  bool fused = false;
  bool mcu = false;
  FusedPlan fp;
  /* PACK expansion: built when every coef_off is a multiple of 64 */
  bool can_unpack = false;
  Buffer d_unpack_segs, d_unpack_work;
  std::vector<int> img_first_unpack_cta; /* size n+1 */
};

/* -------------------------------------------------------------------------- */

extern "C" int jgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" jgpu_ctx *jgpu_create(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    jgpu_fail("no CUDA device available (%s)", cudaGetErrorString(e));
    cudaGetLastError();
    return nullptr;
  }
  if (device < 0 || device >= n) {
    jgpu_fail("CUDA device %d out of range (have %d)", device, n);
    return nullptr;
  }
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(device)) != cudaSuccess ||
      (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    jgpu_fail("cudaSetDevice(%d) failed (%s)", device, cudaGetErrorString(e));
    return nullptr;
  }
  if (prop.major != 10) {
    jgpu_fail("device %d is sm_%d%d; this library contains sm_100a code only",
              device, prop.major, prop.minor);
    return nullptr;
  }
  jgpu_ctx *ctx = new (std::nothrow) jgpu_ctx();
  if (!ctx) {
    jgpu_fail("out of host memory");
    return nullptr;
  }
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  for (int i = 0; i < kHostStreams; i++) {
    ctx->h_in[i].pinned_host = ctx->h_out[i].pinned_host = true;
    ctx->hz_stream.pinned_host = ctx->hz_files.pinned_host = ctx->hz_tables.pinned_host = true;
    ctx->hz_segs.pinned_host = ctx->hz_status.pinned_host = ctx->hz_coef.pinned_host = true;
    if (cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->events[i], cudaEventDisableTiming) != cudaSuccess) {
      jgpu_fail("could not create CUDA streams");
      jgpu_destroy(ctx);
      return nullptr;
    }
  }
  if (mcu_configure(device) != cudaSuccess) {
    jgpu_fail("could not configure the fused kernel (%s)", cudaGetErrorString(cudaGetLastError()));
    jgpu_destroy(ctx);
    return nullptr;
  }
  if (fused_configure(device) != cudaSuccess) {
    jgpu_fail("could not configure the fused kernel (%s)", cudaGetErrorString(cudaGetLastError()));
    jgpu_destroy(ctx);
    return nullptr;
  }
  if (huff_configure() != cudaSuccess) {
    jgpu_fail("could not configure the entropy decoder (%s)", cudaGetErrorString(cudaGetLastError()));
    jgpu_destroy(ctx);
    return nullptr;
  }
  return ctx;
}

extern "C" void jgpu_destroy(jgpu_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->cached_plan) jgpu_plan_destroy(ctx->cached_plan);
  if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
  for (cudaEvent_t e : ctx->up_events) cudaEventDestroy(e);
  for (int i = 0; i < kHostStreams; i++) {
    if (ctx->streams[i]) cudaStreamDestroy(ctx->streams[i]);
    if (ctx->events[i]) cudaEventDestroy(ctx->events[i]);
    ctx->h_in[i].release();
    ctx->h_out[i].release();
  }
  ctx->d_coef.release();
  ctx->d_qtabs.release();
  ctx->d_rgb.release();
  ctx->d_yuv.release();
  ctx->d_pack.release();
  ctx->d_index.release();
  ctx->d_pack_off.release();
  for (Buffer *b : {&ctx->hz_stream, &ctx->hz_files, &ctx->hz_tables, &ctx->hz_segs, &ctx->hz_status, &ctx->hz_coef,
                    &ctx->dz_stream, &ctx->dz_files, &ctx->dz_tables, &ctx->dz_segs, &ctx->dz_status,
                    &ctx->dz_sub[0], &ctx->dz_sub[1], &ctx->dz_sub[2], &ctx->dz_sub[3],
                    &ctx->dz_carry[0], &ctx->dz_carry[1]}) {
    b->release();
  }
  for (int i = 0; i < kHostStreams; i++) ctx->dz_dc[i].release();
  delete ctx;
}

extern "C" void *jgpu_host_alloc(size_t bytes) {
  void *p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    jgpu_fail("cudaHostAlloc(%zu) failed (%s)", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}

extern "C" void jgpu_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

/* -------------------------------------------------------------------------- */
/* plans                                                                      */

/* JGPU_FUSED_IMPL=v9 selects the warp-role kernel of jgpu_fused.cu (A/B runs); default: jgpu_mcu.cu */
static bool use_v9() {
  const char *e = getenv("JGPU_FUSED_IMPL");
  return e != nullptr && strcmp(e, "v9") == 0;
}

static bool fused_eligible(const jgpu_image_desc &d, const jgpu_layout &lay, unsigned flags,
                           int *mode) {
  (void)lay;
  if (d.ncomps == 1) {
    *mode = kModeGray;
  } else {
    if (d.hsamp[1] != 1 || d.vsamp[1] != 1 || d.hsamp[2] != 1 || d.vsamp[2] != 1) return false;
    if (d.hsamp[0] == 1 && d.vsamp[0] == 1) *mode = kMode444;
    else if (d.hsamp[0] == 2 && d.vsamp[0] == 1) *mode = kMode422;
    else if (d.hsamp[0] == 2 && d.vsamp[0] == 2) *mode = kMode420;
    else if (d.hsamp[0] == 1 && d.vsamp[0] == 2) *mode = kMode440;
    else if (d.hsamp[0] == 4 && d.vsamp[0] == 1) *mode = kMode411;
    else return false;
    if (*mode >= kMode411 && (use_v9() || !mcu_has_mode(*mode))) return false;
  }
  /* the fused kernels write pixels OR planes and address coefficients by 128-byte row */
  const unsigned out = flags & (JGPU_OUT_RGB | JGPU_OUT_YUV);
  if (out != JGPU_OUT_RGB && out != JGPU_OUT_YUV) return false;
  if (out == JGPU_OUT_YUV && use_v9()) return false;
  if (d.coef_off & 63) return false;
  return true;
}

/* Work lists of k_unpack: one segment per (image, plane), one CTA per 256 blocks. */
static int build_unpack_lists(jgpu_plan *plan) {
  plan->can_unpack = true;
  for (int i = 0; i < plan->n; i++) {
    if (plan->descs[i].coef_off & 63) plan->can_unpack = false;
  }
  if (!plan->can_unpack) return 0;
  std::vector<UnpackSeg> segs;
  std::vector<UnpackWork> work;
  plan->img_first_unpack_cta.assign(plan->n + 1, 0);
  for (int i = 0; i < plan->n; i++) {
    const jgpu_image_desc &d = plan->descs[i];
    const jgpu_layout &lay = plan->layouts[i];
    plan->img_first_unpack_cta[i] = (int)work.size();
    for (int p = 0; p < d.ncomps; p++) {
      UnpackSeg s;
      s.block0 = (d.coef_off + lay.plane[p].coef_off) / 64;
      s.nblocks = lay.plane[p].hblocks * lay.plane[p].vblocks;
      s.img = i;
      for (int first = 0; first < s.nblocks; first += kUnpackThreads) {
        UnpackWork w = {(int32_t)segs.size(), first};
        work.push_back(w);
      }
      segs.push_back(s);
    }
  }
  plan->img_first_unpack_cta[plan->n] = (int)work.size();
  return upload(plan->d_unpack_segs, segs) || upload(plan->d_unpack_work, work);
}

extern "C" jgpu_plan *jgpu_plan_create(jgpu_ctx *ctx, const jgpu_image_desc *descs, int n,
                                       unsigned flags) {
  if (!ctx || !descs || n <= 0) {
    jgpu_fail("jgpu_plan_create: bad arguments");
    return nullptr;
  }
  if (!(flags & (JGPU_OUT_RGB | JGPU_OUT_YUV))) {
    jgpu_fail("jgpu_plan_create: flags select no output");
    return nullptr;
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) {
    jgpu_fail("cudaSetDevice(%d) failed", ctx->device);
    return nullptr;
  }
  jgpu_plan *plan = new (std::nothrow) jgpu_plan();
  if (!plan) {
    jgpu_fail("out of host memory");
    return nullptr;
  }
  plan->ctx = ctx;
  plan->flags = flags;
  plan->n = n;
  plan->descs.assign(descs, descs + n);
  plan->layouts.resize(n);
  bool all_fused = fused_available() && !(flags & JGPU_FORCE_GENERIC);
  std::vector<int> modes(n, 0);
  for (int i = 0; i < n; i++) {
    const jgpu_image_desc &d = descs[i];
    if (jgpu_layout_query(&d, &plan->layouts[i])) goto fail;
    if (d.coef_off < 0 || (d.coef_off & 7)) {
      jgpu_fail("image %d: coef_off must be a non-negative multiple of 8 int16", i);
      goto fail;
    }
    if ((flags & JGPU_OUT_RGB) && d.rgb_off < 0) {
      jgpu_fail("image %d: rgb_off must be non-negative", i);
      goto fail;
    }
    if ((flags & JGPU_OUT_YUV) && (d.yuv_off < 0 || (d.yuv_off & 15))) {
      jgpu_fail("image %d: yuv_off must be a non-negative multiple of 16", i);
      goto fail;
    }
    if (d.qtab_set < 0) {
      jgpu_fail("image %d: negative qtab_set", i);
      goto fail;
    }
    plan->bytes += 128 * plan->layouts[i].coded_blocks;
    if (flags & JGPU_OUT_RGB) plan->bytes += plan->layouts[i].rgb_len;
    if (flags & JGPU_OUT_YUV) plan->bytes += plan->layouts[i].data_len;
    if (!fused_eligible(d, plan->layouts[i], flags, &modes[i])) all_fused = false;
  }
  plan->fused = all_fused;
  if (build_unpack_lists(plan)) goto fail;

  if (plan->fused) {
    plan->mcu = !use_v9();
    if (plan->mcu ? mcu_plan_build(plan->fp, plan->descs.data(), plan->layouts.data(), modes.data(), n, flags,
                                   ctx->sm_count)
                  : fused_plan_build(plan->fp, plan->descs.data(), plan->layouts.data(), modes.data(), n, flags,
                                     ctx->sm_count)) {
      goto fail;
    }
    return plan;
  }

  {
    /* generic path: planes kernel + colour kernel */
    std::vector<PlaneSeg> segs;
    std::vector<PairWork> pair_work;
    std::vector<ColourImage> cimgs;
    std::vector<ColourWork> colour_work;
    plan->use_scratch = !(flags & JGPU_OUT_YUV);
    plan->scratch_off.assign(n, 0);
    plan->img_first_pair_cta.assign(n + 1, 0);
    plan->img_first_colour_cta.assign(n + 1, 0);
    for (int i = 0; i < n; i++) {
      const jgpu_image_desc &d = descs[i];
      const jgpu_layout &lay = plan->layouts[i];
      int64_t planes_base = d.yuv_off;
      if (plan->use_scratch) {
        plan->scratch_off[i] = plan->scratch_len;
        planes_base = plan->scratch_len;
        plan->scratch_len += (lay.data_len + 15) & ~(int64_t)15;
      }
      plan->img_first_pair_cta[i] = (int)pair_work.size();
      plan->img_first_colour_cta[i] = (int)colour_work.size();
      ColourImage ci;
      memset(&ci, 0, sizeof(ci));
      for (int p = 0; p < d.ncomps; p++) {
        const jgpu_plane_layout &pl = lay.plane[p];
        PlaneSeg s;
        s.coef_off = d.coef_off + pl.coef_off;
        s.out_off = planes_base + pl.data_off;
        s.hblocks = pl.hblocks;
        s.nblocks = pl.hblocks * pl.vblocks;
        s.pitch = pl.width;
        s.qidx = d.qtab_set * 4 + d.tq[p];
        int npairs = (s.nblocks + 1) / 2;
        for (int first = 0; first < npairs; first += kPairThreads) {
          PairWork w = {(int32_t)segs.size(), first};
          pair_work.push_back(w);
        }
        segs.push_back(s);
        ci.plane_off[p] = s.out_off;
        ci.pitch[p] = pl.width;
        ci.xdec[p] = pl.xdec;
        ci.ydec[p] = pl.ydec;
      }
      if (flags & JGPU_OUT_RGB) {
        ci.rgb_off = d.rgb_off;
        ci.width = d.width;
        ci.height = d.height;
        ci.ncomps = d.ncomps;
        ci.groups_per_row = (d.width + 3) / 4;
        int64_t items = (int64_t)ci.groups_per_row * d.height;
        for (int64_t first = 0; first < items; first += kColourThreads) {
          ColourWork w = {(int32_t)cimgs.size(), (int32_t)first};
          colour_work.push_back(w);
        }
        cimgs.push_back(ci);
      }
    }
    plan->img_first_pair_cta[n] = (int)pair_work.size();
    plan->img_first_colour_cta[n] = (int)colour_work.size();
    if (upload(plan->d_segs, segs) || upload(plan->d_pair_work, pair_work) ||
        upload(plan->d_cimgs, cimgs) || upload(plan->d_colour_work, colour_work)) {
      goto fail;
    }
    if (plan->use_scratch && plan->d_scratch.reserve((size_t)plan->scratch_len + 16)) goto fail;
  }
  return plan;

fail:
  jgpu_plan_destroy(plan);
  return nullptr;
}

extern "C" void jgpu_plan_destroy(jgpu_plan *plan) {
  if (!plan) return;
  if (plan->ctx) cudaSetDevice(plan->ctx->device);
  plan->d_segs.release();
  plan->d_pair_work.release();
  plan->d_cimgs.release();
  plan->d_colour_work.release();
  plan->d_scratch.release();
  plan->d_unpack_segs.release();
  plan->d_unpack_work.release();
  if (plan->mcu) mcu_plan_release(plan->fp);
  else fused_plan_release(plan->fp);
  delete plan;
}

extern "C" int jgpu_plan_launches(const jgpu_plan *plan) {
  if (!plan) return 0;
  if (plan->fused) return plan->mcu ? mcu_plan_launches(plan->fp) : fused_plan_launches(plan->fp);
  return (plan->flags & JGPU_OUT_RGB) ? 2 : 1;
}

extern "C" int64_t jgpu_plan_bytes(const jgpu_plan *plan) { return plan ? plan->bytes : 0; }

/* Runs images [i0, i1) of the plan. */
static int plan_run_range(jgpu_plan *plan, int i0, int i1, const int16_t *d_coef,
                          const uint16_t *d_qtabs, int n_sets, uint8_t *d_rgb, uint8_t *d_yuv,
                          cudaStream_t stream) {
  if (i0 >= i1) return 0;
  if (plan->fused) {
    if (plan->mcu && d_yuv && (reinterpret_cast<uintptr_t>(d_yuv) & 15)) {
      return jgpu_fail("the planes buffer must be 16-byte aligned");
    }
    if (plan->mcu) return mcu_plan_launch(plan->fp, i0, i1, d_coef, d_qtabs, n_sets, d_rgb, d_yuv, stream);
    return fused_plan_launch(plan->fp, i0, i1, d_coef, d_qtabs, n_sets, d_rgb, stream);
  }
  uint8_t *planes = plan->use_scratch ? (uint8_t *)plan->d_scratch.ptr : d_yuv;
  int c0 = plan->img_first_pair_cta[i0], c1 = plan->img_first_pair_cta[i1];
  CU_TRY(launch_coef_to_planes((const PlaneSeg *)plan->d_segs.ptr,
                               (const PairWork *)plan->d_pair_work.ptr + c0, c1 - c0, d_coef,
                               d_qtabs, planes, stream));
  if (plan->flags & JGPU_OUT_RGB) {
    c0 = plan->img_first_colour_cta[i0];
    c1 = plan->img_first_colour_cta[i1];
    CU_TRY(launch_planes_to_rgb((const ColourImage *)plan->d_cimgs.ptr,
                                (const ColourWork *)plan->d_colour_work.ptr + c0, c1 - c0,
                                planes, d_rgb, stream));
  }
  return 0;
}

extern "C" int jgpu_plan_run(jgpu_plan *plan, const int16_t *d_coef, const uint16_t *d_qtabs,
                             int n_sets, uint8_t *d_rgb, uint8_t *d_yuv, void *stream) {
  if (!plan || !d_coef || !d_qtabs) return jgpu_fail("jgpu_plan_run: NULL argument");
  if ((plan->flags & JGPU_OUT_RGB) && !d_rgb) return jgpu_fail("jgpu_plan_run: d_rgb is NULL");
  if ((plan->flags & JGPU_OUT_YUV) && !d_yuv) return jgpu_fail("jgpu_plan_run: d_yuv is NULL");
  for (int i = 0; i < plan->n; i++) {
    if (plan->descs[i].qtab_set >= n_sets) {
      return jgpu_fail("image %d uses table set %d but only %d were passed", i,
                       plan->descs[i].qtab_set, n_sets);
    }
  }
  CU_TRY(cudaSetDevice(plan->ctx->device));
  return plan_run_range(plan, 0, plan->n, d_coef, d_qtabs, n_sets, d_rgb, d_yuv, (cudaStream_t)stream);
}

/* -------------------------------------------------------------------------- */
/* host-buffer entry point                                                    */

extern "C" int jgpu_host_is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (p == nullptr || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return a.type == cudaMemoryTypeHost;
}

static bool is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

static bool same_descs(const std::vector<jgpu_image_desc> &a, const jgpu_image_desc *b, int n) {
  return (int)a.size() == n && memcmp(a.data(), b, sizeof(jgpu_image_desc) * n) == 0;
}

extern "C" int jgpu_decode_batch_host(jgpu_ctx *ctx, const jgpu_image_desc *descs, int n,
                                      unsigned flags, const int16_t *h_coef,
                                      const uint16_t *h_qtabs, int n_sets, uint8_t *h_rgb,
                                      uint8_t *h_yuv) {
  if (!ctx || !descs || n <= 0 || !h_coef || !h_qtabs || n_sets <= 0) {
    return jgpu_fail("jgpu_decode_batch_host: bad arguments");
  }
  if ((flags & JGPU_OUT_RGB) && !h_rgb) return jgpu_fail("jgpu_decode_batch_host: h_rgb is NULL");
  if ((flags & JGPU_OUT_YUV) && !h_yuv) return jgpu_fail("jgpu_decode_batch_host: h_yuv is NULL");
  CU_TRY(cudaSetDevice(ctx->device));

  if (!ctx->cached_plan || ctx->cached_flags != flags || !same_descs(ctx->cached_descs, descs, n)) {
    if (ctx->cached_plan) jgpu_plan_destroy(ctx->cached_plan);
    ctx->cached_plan = jgpu_plan_create(ctx, descs, n, flags);
    if (!ctx->cached_plan) return EXIT_FAILURE;
    ctx->cached_descs.assign(descs, descs + n);
    ctx->cached_flags = flags;
  }
  jgpu_plan *plan = ctx->cached_plan;
  for (int i = 0; i < n; i++) {
    if (descs[i].qtab_set >= n_sets) {
      return jgpu_fail("image %d uses table set %d but only %d were passed", i, descs[i].qtab_set,
                       n_sets);
    }
  }

  /* device mirrors with the same offsets as the host buffers */
  int64_t coef_end = 0, rgb_end = 0, yuv_end = 0;
  for (int i = 0; i < n; i++) {
    const jgpu_layout &lay = plan->layouts[i];
    coef_end = std::max(coef_end, descs[i].coef_off + lay.coef_len);
    if (flags & JGPU_OUT_RGB) rgb_end = std::max(rgb_end, descs[i].rgb_off + lay.rgb_len);
    if (flags & JGPU_OUT_YUV) yuv_end = std::max(yuv_end, descs[i].yuv_off + lay.data_len);
  }
  if (ctx->d_coef.reserve((size_t)coef_end * 2 + 256) ||
      ctx->d_qtabs.reserve((size_t)n_sets * 4 * 64 * 2) ||
      ((flags & JGPU_OUT_RGB) && ctx->d_rgb.reserve((size_t)rgb_end + 256)) ||
      ((flags & JGPU_OUT_YUV) && ctx->d_yuv.reserve((size_t)yuv_end + 256))) {
    return EXIT_FAILURE;
  }
  int16_t *d_coef = (int16_t *)ctx->d_coef.ptr;
  uint16_t *d_qtabs = (uint16_t *)ctx->d_qtabs.ptr;
  uint8_t *d_rgb = (uint8_t *)ctx->d_rgb.ptr;
  uint8_t *d_yuv = (uint8_t *)ctx->d_yuv.ptr;

  const bool in_pinned = is_pinned(h_coef);
  const bool rgb_pinned = !(flags & JGPU_OUT_RGB) || is_pinned(h_rgb);
  const bool yuv_pinned = !(flags & JGPU_OUT_YUV) || is_pinned(h_yuv);

  CU_TRY(cudaMemcpyAsync(d_qtabs, h_qtabs, (size_t)n_sets * 4 * 64 * 2, cudaMemcpyHostToDevice,
                         ctx->streams[0]));
  CU_TRY(cudaEventRecord(ctx->events[0], ctx->streams[0]));
  for (int s = 1; s < kHostStreams; s++) CU_TRY(cudaStreamWaitEvent(ctx->streams[s], ctx->events[0], 0));

  /* chunk the batch so that copies of chunk k+1 overlap the kernel and the
   * read-back of chunk k; ~48 MB of coefficients per chunk */
  const int64_t chunk_bytes = 48ll << 20;
  int i0 = 0, chunk = 0;
  struct Pending { int i0, i1; };
  std::vector<Pending> pending_out[kHostStreams];
  while (i0 < n) {
    int i1 = i0;
    int64_t acc = 0;
    while (i1 < n && (i1 == i0 || acc + plan->layouts[i1].coef_len * 2 <= chunk_bytes)) {
      acc += plan->layouts[i1].coef_len * 2;
      i1++;
    }
    const int s = chunk % kHostStreams;
    cudaStream_t st = ctx->streams[s];
    /* H2D */
    if (in_pinned) {
      for (int i = i0; i < i1; i++) {
        CU_TRY(cudaMemcpyAsync(d_coef + descs[i].coef_off, h_coef + descs[i].coef_off,
                               (size_t)plan->layouts[i].coef_len * 2, cudaMemcpyHostToDevice, st));
      }
    } else {
      /* pageable source: bounce through this stream's pinned buffer */
      CU_TRY(cudaStreamSynchronize(st));
      if (ctx->h_in[s].reserve((size_t)acc)) return EXIT_FAILURE;
      size_t off = 0;
      for (int i = i0; i < i1; i++) {
        size_t bytes = (size_t)plan->layouts[i].coef_len * 2;
        memcpy((char *)ctx->h_in[s].ptr + off, h_coef + descs[i].coef_off, bytes);
        CU_TRY(cudaMemcpyAsync(d_coef + descs[i].coef_off, (char *)ctx->h_in[s].ptr + off, bytes,
                               cudaMemcpyHostToDevice, st));
        off += bytes;
      }
    }
    /* kernels */
    if (plan_run_range(plan, i0, i1, d_coef, d_qtabs, n_sets, d_rgb, d_yuv, st)) return EXIT_FAILURE;
    /* D2H */
    if (rgb_pinned && yuv_pinned) {
      for (int i = i0; i < i1; i++) {
        if (flags & JGPU_OUT_RGB) {
          CU_TRY(cudaMemcpyAsync(h_rgb + descs[i].rgb_off, d_rgb + descs[i].rgb_off,
                                 (size_t)plan->layouts[i].rgb_len, cudaMemcpyDeviceToHost, st));
        }
        if (flags & JGPU_OUT_YUV) {
          CU_TRY(cudaMemcpyAsync(h_yuv + descs[i].yuv_off, d_yuv + descs[i].yuv_off,
                                 (size_t)plan->layouts[i].data_len, cudaMemcpyDeviceToHost, st));
        }
      }
    } else {
      /* pageable destination: stage per chunk, drain synchronously */
      size_t need = 0;
      for (int i = i0; i < i1; i++) {
        if (flags & JGPU_OUT_RGB) need += (size_t)plan->layouts[i].rgb_len;
        if (flags & JGPU_OUT_YUV) need += (size_t)plan->layouts[i].data_len;
      }
      if (ctx->h_out[s].reserve(need)) return EXIT_FAILURE;
      size_t off = 0;
      for (int i = i0; i < i1; i++) {
        if (flags & JGPU_OUT_RGB) {
          CU_TRY(cudaMemcpyAsync((char *)ctx->h_out[s].ptr + off, d_rgb + descs[i].rgb_off,
                                 (size_t)plan->layouts[i].rgb_len, cudaMemcpyDeviceToHost, st));
          off += (size_t)plan->layouts[i].rgb_len;
        }
        if (flags & JGPU_OUT_YUV) {
          CU_TRY(cudaMemcpyAsync((char *)ctx->h_out[s].ptr + off, d_yuv + descs[i].yuv_off,
                                 (size_t)plan->layouts[i].data_len, cudaMemcpyDeviceToHost, st));
          off += (size_t)plan->layouts[i].data_len;
        }
      }
      CU_TRY(cudaStreamSynchronize(st));
      off = 0;
      for (int i = i0; i < i1; i++) {
        if (flags & JGPU_OUT_RGB) {
          memcpy(h_rgb + descs[i].rgb_off, (char *)ctx->h_out[s].ptr + off,
                 (size_t)plan->layouts[i].rgb_len);
          off += (size_t)plan->layouts[i].rgb_len;
        }
        if (flags & JGPU_OUT_YUV) {
          memcpy(h_yuv + descs[i].yuv_off, (char *)ctx->h_out[s].ptr + off,
                 (size_t)plan->layouts[i].data_len);
          off += (size_t)plan->layouts[i].data_len;
        }
      }
    }
    i0 = i1;
    chunk++;
  }
  for (int s = 0; s < kHostStreams; s++) CU_TRY(cudaStreamSynchronize(ctx->streams[s]));
  return EXIT_SUCCESS;
}

/* -------------------------------------------------------------------------- */
/* PACK input                                                                 */

static int plan_unpack_range(jgpu_plan *plan, int i0, int i1, const uint16_t *d_pack,
                             const int64_t *d_pack_off, const int32_t *d_index, int16_t *d_coef,
                             cudaStream_t stream) {
  if (!plan->can_unpack) {
    return jgpu_fail("PACK input needs every coef_off to be a multiple of 64 int16");
  }
  const int c0 = plan->img_first_unpack_cta[i0], c1 = plan->img_first_unpack_cta[i1];
  CU_TRY(launch_unpack((const UnpackSeg *)plan->d_unpack_segs.ptr,
                       (const UnpackWork *)plan->d_unpack_work.ptr + c0, c1 - c0, d_pack, d_pack_off,
                       d_index, d_coef, stream));
  return 0;
}

extern "C" int jgpu_plan_unpack(jgpu_plan *plan, const uint16_t *d_pack, const int64_t *d_pack_off,
                                const int32_t *d_index, int16_t *d_coef, void *stream) {
  if (!plan || !d_pack || !d_pack_off || !d_index || !d_coef) {
    return jgpu_fail("jgpu_plan_unpack: NULL argument");
  }
  if (reinterpret_cast<uintptr_t>(d_coef) & 15) {
    return jgpu_fail("jgpu_plan_unpack: d_coef must be 16-byte aligned");
  }
  CU_TRY(cudaSetDevice(plan->ctx->device));
  return plan_unpack_range(plan, 0, plan->n, d_pack, d_pack_off, d_index, d_coef, (cudaStream_t)stream);
}

extern "C" int jgpu_decode_batch_host_packed(jgpu_ctx *ctx, const jgpu_image_desc *descs, int n,
                                             unsigned flags, const uint16_t *h_pack,
                                             const int64_t *pack_off, const int32_t *h_index,
                                             const uint16_t *h_qtabs, int n_sets, uint8_t *h_rgb,
                                             uint8_t *h_yuv) {
  if (!ctx || !descs || n <= 0 || !h_pack || !pack_off || !h_index || !h_qtabs || n_sets <= 0) {
    return jgpu_fail("jgpu_decode_batch_host_packed: bad arguments");
  }
  if ((flags & JGPU_OUT_RGB) && !h_rgb) return jgpu_fail("jgpu_decode_batch_host_packed: h_rgb is NULL");
  if ((flags & JGPU_OUT_YUV) && !h_yuv) return jgpu_fail("jgpu_decode_batch_host_packed: h_yuv is NULL");
  for (int i = 0; i < n; i++) {
    if (pack_off[i] < 0 || pack_off[i + 1] < pack_off[i]) {
      return jgpu_fail("jgpu_decode_batch_host_packed: pack_off must be non-decreasing");
    }
  }
  CU_TRY(cudaSetDevice(ctx->device));
  if (!ctx->cached_plan || ctx->cached_flags != flags || !same_descs(ctx->cached_descs, descs, n)) {
    if (ctx->cached_plan) jgpu_plan_destroy(ctx->cached_plan);
    ctx->cached_plan = jgpu_plan_create(ctx, descs, n, flags);
    if (!ctx->cached_plan) return EXIT_FAILURE;
    ctx->cached_descs.assign(descs, descs + n);
    ctx->cached_flags = flags;
  }
  jgpu_plan *plan = ctx->cached_plan;
  if (!plan->can_unpack) return jgpu_fail("PACK input needs every coef_off to be a multiple of 64 int16");
  for (int i = 0; i < n; i++) {
    if (descs[i].qtab_set >= n_sets) {
      return jgpu_fail("image %d uses table set %d but only %d were passed", i, descs[i].qtab_set, n_sets);
    }
  }
  int64_t coef_end = 0, rgb_end = 0, yuv_end = 0;
  for (int i = 0; i < n; i++) {
    const jgpu_layout &lay = plan->layouts[i];
    coef_end = std::max(coef_end, descs[i].coef_off + lay.coef_len);
    if (flags & JGPU_OUT_RGB) rgb_end = std::max(rgb_end, descs[i].rgb_off + lay.rgb_len);
    if (flags & JGPU_OUT_YUV) yuv_end = std::max(yuv_end, descs[i].yuv_off + lay.data_len);
  }
  const int64_t pack_words = pack_off[n];
  if (ctx->d_coef.reserve((size_t)coef_end * 2 + 256) ||
      ctx->d_pack.reserve((size_t)pack_words * 2 + 256) ||
      ctx->d_index.reserve((size_t)(coef_end / 64 + 1) * 4 + 256) ||
      ctx->d_pack_off.reserve((size_t)(n + 1) * 8) ||
      ctx->d_qtabs.reserve((size_t)n_sets * 4 * 64 * 2) ||
      ((flags & JGPU_OUT_RGB) && ctx->d_rgb.reserve((size_t)rgb_end + 256)) ||
      ((flags & JGPU_OUT_YUV) && ctx->d_yuv.reserve((size_t)yuv_end + 256))) {
    return EXIT_FAILURE;
  }
  int16_t *d_coef = (int16_t *)ctx->d_coef.ptr;
  uint16_t *d_pack = (uint16_t *)ctx->d_pack.ptr;
  int32_t *d_index = (int32_t *)ctx->d_index.ptr;
  int64_t *d_pack_off = (int64_t *)ctx->d_pack_off.ptr;
  uint16_t *d_qtabs = (uint16_t *)ctx->d_qtabs.ptr;
  uint8_t *d_rgb = (uint8_t *)ctx->d_rgb.ptr;
  uint8_t *d_yuv = (uint8_t *)ctx->d_yuv.ptr;

  /* pageable sources and destinations go through the driver's own staging here (cudaMemcpyAsync
   * from pageable memory is synchronous with respect to the host); pinned ones overlap */
  CU_TRY(cudaMemcpyAsync(d_qtabs, h_qtabs, (size_t)n_sets * 4 * 64 * 2, cudaMemcpyHostToDevice,
                         ctx->streams[0]));
  CU_TRY(cudaMemcpyAsync(d_pack_off, pack_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice,
                         ctx->streams[0]));
  CU_TRY(cudaEventRecord(ctx->events[0], ctx->streams[0]));
  for (int s = 1; s < kHostStreams; s++) CU_TRY(cudaStreamWaitEvent(ctx->streams[s], ctx->events[0], 0));

  /* chunks of ~48 MB of dense coefficients, as in jgpu_decode_batch_host, so that the read-back
   * of chunk k overlaps the upload and kernels of chunk k+1 and the expanded planes of a chunk
   * are still in L2 when the decode kernel reads them */
  const int64_t chunk_bytes = 48ll << 20;
  int i0 = 0, chunk = 0;
  while (i0 < n) {
    int i1 = i0;
    int64_t acc = 0;
    while (i1 < n && (i1 == i0 || acc + plan->layouts[i1].coef_len * 2 <= chunk_bytes)) {
      acc += plan->layouts[i1].coef_len * 2;
      i1++;
    }
    cudaStream_t st = ctx->streams[chunk % kHostStreams];
    if (pack_off[i1] > pack_off[i0]) {
      CU_TRY(cudaMemcpyAsync(d_pack + pack_off[i0], h_pack + pack_off[i0],
                             (size_t)(pack_off[i1] - pack_off[i0]) * 2, cudaMemcpyHostToDevice, st));
    }
    for (int i = i0; i < i1; i++) {
      CU_TRY(cudaMemcpyAsync(d_index + descs[i].coef_off / 64, h_index + descs[i].coef_off / 64,
                             (size_t)(plan->layouts[i].coef_len / 64) * 4, cudaMemcpyHostToDevice, st));
    }
    if (plan_unpack_range(plan, i0, i1, d_pack, d_pack_off, d_index, d_coef, st)) return EXIT_FAILURE;
    if (plan_run_range(plan, i0, i1, d_coef, d_qtabs, n_sets, d_rgb, d_yuv, st)) return EXIT_FAILURE;
    for (int i = i0; i < i1; i++) {
      if (flags & JGPU_OUT_RGB) {
        CU_TRY(cudaMemcpyAsync(h_rgb + descs[i].rgb_off, d_rgb + descs[i].rgb_off,
                               (size_t)plan->layouts[i].rgb_len, cudaMemcpyDeviceToHost, st));
      }
      if (flags & JGPU_OUT_YUV) {
        CU_TRY(cudaMemcpyAsync(h_yuv + descs[i].yuv_off, d_yuv + descs[i].yuv_off,
                               (size_t)plan->layouts[i].data_len, cudaMemcpyDeviceToHost, st));
      }
    }
    i0 = i1;
    chunk++;
  }
  for (int s = 0; s < kHostStreams; s++) CU_TRY(cudaStreamSynchronize(ctx->streams[s]));
  return EXIT_SUCCESS;
}

/* -------------------------------------------------------------------------- */
/* one image on the reference's structs                                       */

extern "C" int jgpu_decode_image(jgpu_ctx *ctx, const jpeg_header *header, image *img,
                                 jpeg_decode_out out) {
  if (!ctx || !header || !img) return jgpu_fail("jgpu_decode_image: NULL argument");
  if (out != JPEG_DECODE_YUV && out != JPEG_DECODE_RGB) {
    return jgpu_fail("jgpu_decode_image: output must be yuv or rgb");
  }
  jgpu_image_desc d;
  jgpu_layout lay;
  if (jgpu_desc_from_header(header, &d) || jgpu_layout_query(&d, &lay)) return EXIT_FAILURE;
  if (img->nplanes != d.ncomps || img->width != d.width || img->height != d.height ||
      !img->coef) {
    return jgpu_fail("jgpu_decode_image: image surface does not match the header");
  }
  for (int p = 0; p < d.ncomps; p++) {
    const image_plane *pl = &img->plane[p];
    if (pl->width != lay.plane[p].width || pl->height != lay.plane[p].height ||
        pl->coef - img->coef != lay.plane[p].coef_off) {
      return jgpu_fail("jgpu_decode_image: plane %d layout differs from image_init's", p);
    }
  }
  uint16_t qt[NQUANT_MAX * 64];
  for (int t = 0; t < NQUANT_MAX; t++) memcpy(qt + 64 * t, header->quant[t].tbl, 128);
  d.coef_off = 0;
  d.qtab_set = 0;
  if (out == JPEG_DECODE_RGB) {
    d.rgb_off = 0;
    d.yuv_off = -1;
    return jgpu_decode_batch_host(ctx, &d, 1, JGPU_OUT_RGB, img->coef, qt, 1, img->pixels, nullptr);
  }
  /* YUV: planes are separate allocations on the image surface */
  d.yuv_off = 0;
  Buffer &stage = ctx->h_out[kHostStreams - 1];
  if (stage.reserve((size_t)lay.data_len)) return EXIT_FAILURE;
  if (jgpu_decode_batch_host(ctx, &d, 1, JGPU_OUT_YUV, img->coef, qt, 1, nullptr,
                             (uint8_t *)stage.ptr)) {
    return EXIT_FAILURE;
  }
  for (int p = 0; p < d.ncomps; p++) {
    memcpy(img->plane[p].data, (uint8_t *)stage.ptr + lay.plane[p].data_off,
           (size_t)lay.plane[p].width * lay.plane[p].height);
  }
  return EXIT_SUCCESS;
}

/* Same, with the PACK stream in img->coef / img->index (what a front end leaves there after
 * decode_image(..., JPEG_DECODE_PACK), src/xjpeg.c:484-496): `words` = sum of plane[i].packed. */
extern "C" int jgpu_decode_image_packed(jgpu_ctx *ctx, const jpeg_header *header, image *img,
                                        int64_t words, jpeg_decode_out out) {
  if (!ctx || !header || !img) return jgpu_fail("jgpu_decode_image_packed: NULL argument");
  if (out != JPEG_DECODE_YUV && out != JPEG_DECODE_RGB) {
    return jgpu_fail("jgpu_decode_image_packed: output must be yuv or rgb");
  }
  jgpu_image_desc d;
  jgpu_layout lay;
  if (jgpu_desc_from_header(header, &d) || jgpu_layout_query(&d, &lay)) return EXIT_FAILURE;
  if (img->nplanes != d.ncomps || img->width != d.width || img->height != d.height || !img->coef ||
      !img->index) {
    return jgpu_fail("jgpu_decode_image_packed: image surface does not match the header");
  }
  if (words < 0 || words > lay.coef_len) {
    return jgpu_fail("jgpu_decode_image_packed: %lld words do not fit image.coef", (long long)words);
  }
  for (int p = 0; p < d.ncomps; p++) {
    if (img->plane[p].index - img->index != lay.plane[p].coef_off / 64) {
      return jgpu_fail("jgpu_decode_image_packed: plane %d index layout differs from image_init's", p);
    }
  }
  uint16_t qt[NQUANT_MAX * 64];
  for (int t = 0; t < NQUANT_MAX; t++) memcpy(qt + 64 * t, header->quant[t].tbl, 128);
  const int64_t pack_off[2] = {0, words};
  d.coef_off = 0;
  d.qtab_set = 0;
  if (out == JPEG_DECODE_RGB) {
    d.rgb_off = 0;
    d.yuv_off = -1;
    return jgpu_decode_batch_host_packed(ctx, &d, 1, JGPU_OUT_RGB, (const uint16_t *)img->coef, pack_off,
                                         img->index, qt, 1, img->pixels, nullptr);
  }
  d.yuv_off = 0;
  Buffer &stage = ctx->h_out[kHostStreams - 1];
  if (stage.reserve((size_t)lay.data_len)) return EXIT_FAILURE;
  if (jgpu_decode_batch_host_packed(ctx, &d, 1, JGPU_OUT_YUV, (const uint16_t *)img->coef, pack_off, img->index,
                                    qt, 1, nullptr, (uint8_t *)stage.ptr)) {
    return EXIT_FAILURE;
  }
  for (int p = 0; p < d.ncomps; p++) {
    memcpy(img->plane[p].data, (uint8_t *)stage.ptr + lay.plane[p].data_off,
           (size_t)lay.plane[p].width * lay.plane[p].height);
  }
  return EXIT_SUCCESS;
}

/* -------------------------------------------------------------------------- */
/* JPEG files in, RGB out                                                     */

#include <atomic>
#include <chrono>
#include <thread>

#include "jgpu_front.h"

namespace {

/* One file after header parsing. */
struct JpegItem {
  jpeg_info info;                 /* points at the caller's bytes */
  jpeg_decode_ctx *front = nullptr;
  jpeg_header header;
  jgpu_image_desc desc;
  jgpu_layout lay;
  jfront_segments segs;
  bool have_segs = false;
  int first_task = 0, ntasks = 0;
  std::atomic<int> tasks_left{0};
  std::atomic<int> failed{0};
  const char *message = nullptr;
};

struct JpegTask {
  int item;
  int s0, s1;   /* restart intervals; s0 < 0: whole image through the sequential reader */
};

/* An `image` whose coefficient planes live in caller-provided memory (what image_init would
 * lay out, src/image.c:24-97, minus the allocations the QUANT path never touches). */
void bind_image(image *img, const jgpu_image_desc &d, const jgpu_layout &lay, short *coef) {
  memset(img, 0, sizeof(*img));
  img->width = (unsigned short)d.width;
  img->height = (unsigned short)d.height;
  img->nplanes = d.ncomps;
  img->coef = coef;
  for (int p = 0; p < d.ncomps; p++) {
    image_plane *ip = &img->plane[p];
    ip->width = (unsigned short)lay.plane[p].width;
    ip->height = (unsigned short)lay.plane[p].height;
    ip->xstride = 1;
    ip->ystride = ip->width;
    ip->xdec = (unsigned char)lay.plane[p].xdec;
    ip->ydec = (unsigned char)lay.plane[p].ydec;
    ip->cstride = lay.plane[p].cstride;
    ip->coef = coef + lay.plane[p].coef_off;
  }
}

/* Parses one header; on rejection sets message and returns false. */
bool probe_one(const jgpu_jpeg &f, JpegItem &it, jgpu_jpeg_info &out, bool keep_front) {
  const jpeg_decode_ctx_vtbl &v = JFRONT_DECODE_CTX_VTBL;
  memset(&out, 0, sizeof(out));
  out.status = 1;
  if (!f.data || f.size <= 0 || f.size > 0x7fffffff) {
    out.message = "Error, empty or oversized jpeg buffer";
    return false;
  }
  it.info.buf = const_cast<unsigned char *>(f.data);
  it.info.size = (int)f.size;
  it.front = v.decode_alloc(&it.info);
  if (!it.front) {
    out.message = "Error, out of memory";
    return false;
  }
  bool ok = v.decode_header(it.front, &it.header) == EXIT_SUCCESS;
  if (!ok) out.message = "Error reading jpeg headers";
  if (ok && (jgpu_desc_from_header(&it.header, &it.desc) || jgpu_layout_query(&it.desc, &it.lay))) {
    ok = false;
    out.message = "Unsupported component layout";
  }
  if (ok) {
    out.status = 0;
    out.width = it.desc.width;
    out.height = it.desc.height;
    out.ncomps = it.desc.ncomps;
    out.hsamp0 = it.desc.hsamp[0];
    out.vsamp0 = it.desc.vsamp[0];
    out.restart_interval = it.header.restart_interval;
    out.rgb_len = it.lay.rgb_len;
  }
  if (!ok || !keep_front) {
    v.decode_free(it.front);
    it.front = nullptr;
  }
  return ok;
}

}  // namespace

extern "C" int64_t jgpu_jpegs_probe_ex(const jgpu_jpeg *files, int n, unsigned flags, jgpu_jpeg_info *info) {
  if (!files || !info || n <= 0) {
    jgpu_fail("jgpu_jpegs_probe: bad arguments");
    return -1;
  }
  int64_t off = 0;
  for (int i = 0; i < n; i++) {
    JpegItem it;
    if (probe_one(files[i], it, info[i], false)) {
      if (flags & JGPU_JPEGS_OUT_YUV) info[i].rgb_len = it.lay.data_len;   /* padded Y|Cb|Cr planes */
      info[i].rgb_off = off;
      off += (info[i].rgb_len + 255) & ~(int64_t)255;
    }
  }
  return off;
}

extern "C" int64_t jgpu_jpegs_probe(const jgpu_jpeg *files, int n, jgpu_jpeg_info *info) {
  return jgpu_jpegs_probe_ex(files, n, 0, info);
}

static int decode_jpegs_cpu_entropy(jgpu_ctx *ctx, const jgpu_jpeg *files, int n, int nthreads,
                                    uint8_t *h_rgb, int64_t rgb_cap, jgpu_jpeg_info *info) {
  CU_TRY(cudaSetDevice(ctx->device));
  const jpeg_decode_ctx_vtbl &v = JFRONT_DECODE_CTX_VTBL;
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads <= 0) nthreads = 1;

  /* ---- headers, layout ------------------------------------------------------ */
  std::vector<JpegItem> items(n);
  std::vector<int> ok;              /* indices of accepted files, in order */
  std::vector<jgpu_image_desc> descs;
  std::vector<uint16_t> qtabs;
  int64_t rgb_off = 0, coef_off = 0;
  for (int i = 0; i < n; i++) {
    JpegItem &it = items[i];
    if (!probe_one(files[i], it, info[i], true)) continue;
    info[i].rgb_off = rgb_off;
    it.desc.rgb_off = rgb_off;
    it.desc.coef_off = coef_off;
    it.desc.yuv_off = -1;
    it.desc.qtab_set = (int32_t)ok.size();
    rgb_off += (it.lay.rgb_len + 255) & ~(int64_t)255;
    coef_off += it.lay.coef_len;   /* a multiple of 64 */
    for (int t = 0; t < NQUANT_MAX; t++) {
      qtabs.insert(qtabs.end(), it.header.quant[t].tbl, it.header.quant[t].tbl + 64);
    }
    descs.push_back(it.desc);
    ok.push_back(i);
  }
  auto release_fronts = [&]() {
    for (JpegItem &it : items) {
      if (it.have_segs) jfront_segments_free(&it.segs);
      if (it.front) v.decode_free(it.front);
      it.front = nullptr;
      it.have_segs = false;
    }
  };
  if (ok.empty()) {
    release_fronts();
    return jgpu_fail("jgpu_decode_jpegs: no decodable file in the batch");
  }
  {
    /* the last image ends at its own length, not at the next 256-byte boundary */
    const JpegItem &last = items[ok.back()];
    const int64_t need = last.desc.rgb_off + last.lay.rgb_len;
    if (need > rgb_cap) {
      release_fronts();
      return jgpu_fail("jgpu_decode_jpegs: output needs %lld bytes, buffer has %lld", (long long)need,
                       (long long)rgb_cap);
    }
  }
  const int m = (int)ok.size();

  /* ---- plan and buffers ------------------------------------------------------- */
  const unsigned flags = JGPU_OUT_RGB;
  if (!ctx->cached_plan || ctx->cached_flags != flags || !same_descs(ctx->cached_descs, descs.data(), m)) {
    if (ctx->cached_plan) jgpu_plan_destroy(ctx->cached_plan);
    ctx->cached_plan = jgpu_plan_create(ctx, descs.data(), m, flags);
    if (!ctx->cached_plan) {
      release_fronts();
      return EXIT_FAILURE;
    }
    ctx->cached_descs = descs;
    ctx->cached_flags = flags;
  }
  jgpu_plan *plan = ctx->cached_plan;
  Buffer &stage = ctx->h_in[0];   /* pinned: the front end decodes straight into it */
  if (stage.reserve((size_t)coef_off * 2 + 256) || ctx->d_coef.reserve((size_t)coef_off * 2 + 256) ||
      ctx->d_qtabs.reserve(qtabs.size() * 2) || ctx->d_rgb.reserve((size_t)rgb_off + 256)) {
    release_fronts();
    return EXIT_FAILURE;
  }
  int16_t *h_coef = (int16_t *)stage.ptr;
  int16_t *d_coef = (int16_t *)ctx->d_coef.ptr;
  uint16_t *d_qtabs = (uint16_t *)ctx->d_qtabs.ptr;
  uint8_t *d_rgb = (uint8_t *)ctx->d_rgb.ptr;

  /* ---- entropy-decode tasks ---------------------------------------------------- */
  std::vector<JpegTask> tasks;
  const int kMinMcusPerTask = 1024;
  for (int k = 0; k < m; k++) {
    JpegItem &it = items[ok[k]];
    it.first_task = (int)tasks.size();
    if (jfront_find_segments(it.front, &it.segs) == 0) {
      it.have_segs = true;
      int pieces = std::max(1, std::min(it.segs.nseg, it.segs.total_mcus / kMinMcusPerTask));
      for (int p = 0; p < pieces; p++) {
        JpegTask t = {ok[k], (int)((int64_t)it.segs.nseg * p / pieces), (int)((int64_t)it.segs.nseg * (p + 1) / pieces)};
        tasks.push_back(t);
      }
    } else {
      JpegTask t = {ok[k], -1, -1};   /* malformed restart markers: let the sequential reader judge */
      tasks.push_back(t);
    }
    it.ntasks = (int)tasks.size() - it.first_task;
    it.tasks_left.store(it.ntasks, std::memory_order_relaxed);
    info[ok[k]].tasks = it.ntasks;
  }
  std::atomic<int> next_task{0};
  auto worker = [&]() {
    for (;;) {
      const int ti = next_task.fetch_add(1, std::memory_order_relaxed);
      if (ti >= (int)tasks.size()) return;
      const JpegTask &t = tasks[ti];
      JpegItem &it = items[t.item];
      image img;
      bind_image(&img, it.desc, it.lay, h_coef + it.desc.coef_off);
      const char *err = nullptr;
      int rc;
      if (t.s0 >= 0) {
        rc = jfront_decode_segments(it.front, &img, JPEG_DECODE_QUANT, &it.segs, t.s0, t.s1, &err);
      } else {
        /* the sequential reader stops at a missing restart marker and leaves the rest of the
         * planes alone; the reference hands it zeroed planes (image_zero, src/jpeg_gpu.c:1227) */
        memset(h_coef + it.desc.coef_off, 0, (size_t)it.lay.coef_len * 2);
        rc = v.decode_image(it.front, &img, JPEG_DECODE_QUANT);
        if (rc) err = "Error decoding scan";
      }
      if (rc) {
        it.message = err;
        it.failed.store(1, std::memory_order_relaxed);
      }
      it.tasks_left.fetch_sub(1, std::memory_order_release);
    }
  };
  std::vector<std::thread> pool;
  const int nworkers = std::min<int>(nthreads, (int)tasks.size());
  for (int t = 0; t < nworkers; t++) pool.emplace_back(worker);
  auto join_all = [&]() {
    for (std::thread &t : pool) t.join();
    pool.clear();
  };

  /* ---- GPU pipeline: chunk k is uploaded as soon as its images are decoded ------- */
  int rc = EXIT_SUCCESS;
  auto gpu_part = [&]() -> int {
    CU_TRY(cudaMemcpyAsync(d_qtabs, qtabs.data(), qtabs.size() * 2, cudaMemcpyHostToDevice, ctx->streams[0]));
    CU_TRY(cudaEventRecord(ctx->events[0], ctx->streams[0]));
    for (int s = 1; s < kHostStreams; s++) CU_TRY(cudaStreamWaitEvent(ctx->streams[s], ctx->events[0], 0));
    /* the tables come from pageable memory: the copy above has completed on the host side */
    const int64_t chunk_bytes = 48ll << 20;
    int i0 = 0, chunk = 0;
    while (i0 < m) {
      int i1 = i0;
      int64_t acc = 0;
      while (i1 < m && (i1 == i0 || acc + plan->layouts[i1].coef_len * 2 <= chunk_bytes)) {
        acc += plan->layouts[i1].coef_len * 2;
        i1++;
      }
      for (int k = i0; k < i1; k++) {
        JpegItem &it = items[ok[k]];
        while (it.tasks_left.load(std::memory_order_acquire) > 0) std::this_thread::yield();
      }
      cudaStream_t st = ctx->streams[chunk % kHostStreams];
      CU_TRY(cudaMemcpyAsync(d_coef + descs[i0].coef_off, h_coef + descs[i0].coef_off, (size_t)acc,
                             cudaMemcpyHostToDevice, st));
      if (plan_run_range(plan, i0, i1, d_coef, d_qtabs, m, d_rgb, nullptr, st)) return EXIT_FAILURE;
      {
        /* pinned destination: asynchronous; pageable: the call returns when the data is there */
        const int64_t lo = descs[i0].rgb_off;
        const int64_t hi = descs[i1 - 1].rgb_off + plan->layouts[i1 - 1].rgb_len;
        CU_TRY(cudaMemcpyAsync(h_rgb + lo, d_rgb + lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost, st));
      }
      i0 = i1;
      chunk++;
    }
    for (int s = 0; s < kHostStreams; s++) CU_TRY(cudaStreamSynchronize(ctx->streams[s]));
    return EXIT_SUCCESS;
  };
  rc = gpu_part();
  join_all();
  if (rc != EXIT_SUCCESS) {
    /* drain whatever is in flight before the staging buffers can be reused */
    cudaDeviceSynchronize();
  }
  for (int k = 0; k < m; k++) {
    JpegItem &it = items[ok[k]];
    if (it.failed.load()) {
      info[ok[k]].status = 1;
      info[ok[k]].message = it.message ? it.message : "Error decoding scan";
      rc = EXIT_FAILURE;
    }
  }
  if (m != n) rc = EXIT_FAILURE;
  release_fronts();
  if (rc != EXIT_SUCCESS) {
    /* name the first bad file (a CUDA failure has left its own message and no bad file) */
    for (int i = 0; i < n; i++) {
      if (info[i].status) {
        jgpu_fail("file %d: %s", i, info[i].message ? info[i].message : "rejected");
        break;
      }
    }
  }
  return rc;
}

/* -------------------------------------------------------------------------- */
/* JPEG files in, RGB out, Huffman decoding on the GPU           */

/* Host threads only strip the byte stuffing and cut the scan at its restart markers
 * (jfront_huff_prepare, a memchr/memcpy pass); the entropy decoding itself runs in
 * jgpu_huff.cu, followed on the same stream by the fused block decoder.  What crosses the
 * link is the compressed scan in, the pixels out.  Files the GPU decoder does not take
 * (more than 10 blocks per MCU, restart markers out of place) or flags (corrupt or truncated
 * scans, states that did not settle) go through the sequential reader afterwards, which
 * decodes or rejects them exactly as before. */
static int decode_jpegs_gpu_entropy(jgpu_ctx *ctx, const jgpu_jpeg *files, int n, int nthreads,
                                    uint8_t *h_rgb, int64_t rgb_cap, jgpu_jpeg_info *info, bool device_out,
                                    bool yuv_out) {
  CU_TRY(cudaSetDevice(ctx->device));
  const jpeg_decode_ctx_vtbl &v = JFRONT_DECODE_CTX_VTBL;
  constexpr int S = kHuffSubseqWords;
  /* JGPU_TRACE=1: host-side phase times of this call on stderr */
  const bool trace = getenv("JGPU_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  double t_setup = 0, t_prepared = 0, t_enqueued = 0, t_synced = 0;
  if (nthreads <= 0) {
    nthreads = (int)std::thread::hardware_concurrency();
    /* With the pixels staying on the device the call is as long as the chain of groups on the GPU, and this
     * thread feeds that chain: workers on every core starve it (16-core box, 128 4K files: 19.9 ms with 16
     * workers, 17.2 with 14, 15.9-17.4 with 8; profiles/r2_notes.md 14).  With the read-back in the pipeline
     * the link is the limit and all cores are marginally better (66.2 against 67.9 ms). */
    if (device_out && nthreads > 2) nthreads /= 2;
  }
  if (nthreads <= 0) nthreads = 1;

  /* ---- headers, layout ------------------------------------------------------ */
  std::vector<JpegItem> items(n);
  std::vector<int> ok;
  std::vector<jgpu_image_desc> descs;
  std::vector<uint16_t> qtabs;
  int64_t rgb_off = 0, coef_off = 0;
  for (int i = 0; i < n; i++) {
    JpegItem &it = items[i];
    if (!probe_one(files[i], it, info[i], true)) continue;
    /* one output buffer either way: pixels, or (JGPU_JPEGS_OUT_YUV) the padded Y|Cb|Cr planes */
    const int64_t out_len = yuv_out ? it.lay.data_len : it.lay.rgb_len;
    info[i].rgb_off = rgb_off;
    info[i].rgb_len = out_len;
    it.desc.rgb_off = yuv_out ? 0 : rgb_off;
    it.desc.coef_off = coef_off;
    it.desc.yuv_off = yuv_out ? rgb_off : -1;
    it.desc.qtab_set = (int32_t)ok.size();
    rgb_off += (out_len + 255) & ~(int64_t)255;
    coef_off += it.lay.coef_len;
    for (int t = 0; t < NQUANT_MAX; t++) {
      qtabs.insert(qtabs.end(), it.header.quant[t].tbl, it.header.quant[t].tbl + 64);
    }
    descs.push_back(it.desc);
    ok.push_back(i);
  }
  auto release_fronts = [&]() {
    for (JpegItem &it : items) {
      if (it.front) v.decode_free(it.front);
      it.front = nullptr;
    }
  };
  if (ok.empty()) {
    release_fronts();
    return jgpu_fail("jgpu_decode_jpegs: no decodable file in the batch");
  }
  {
    /* the last image ends at its own length, not at the next 256-byte boundary */
    const JpegItem &last = items[ok.back()];
    const int64_t need = yuv_out ? last.desc.yuv_off + last.lay.data_len : last.desc.rgb_off + last.lay.rgb_len;
    if (need > rgb_cap) {
      release_fronts();
      return jgpu_fail("jgpu_decode_jpegs: output needs %lld bytes, buffer has %lld", (long long)need,
                       (long long)rgb_cap);
    }
  }
  const int m = (int)ok.size();
  const unsigned flags = yuv_out ? JGPU_OUT_YUV : JGPU_OUT_RGB;
  if (!ctx->cached_plan || ctx->cached_flags != flags || !same_descs(ctx->cached_descs, descs.data(), m)) {
    if (ctx->cached_plan) jgpu_plan_destroy(ctx->cached_plan);
    ctx->cached_plan = jgpu_plan_create(ctx, descs.data(), m, flags);
    if (!ctx->cached_plan) {
      release_fronts();
      return EXIT_FAILURE;
    }
    ctx->cached_descs = descs;
    ctx->cached_flags = flags;
  }
  jgpu_plan *plan = ctx->cached_plan;

  /* ---- where each file's pieces go (from upper bounds, so the files prepare independently) */
  struct Slot {
    int64_t stream_off;   /* bytes, multiple of 16 */
    int64_t stream_cap;
    uint32_t seg0, seg_cap, subseq0, cta0;
  };
  std::vector<Slot> slot(m + 1);
  {
    int64_t so = 0;
    uint64_t seg = 0, sub = 0, cta = 0;
    for (int k = 0; k < m; k++) {
      int nseg = 0;
      const int64_t bound = (jfront_huff_bound(items[ok[k]].front, S, &nseg) + 15) & ~(int64_t)15;
      const uint64_t nsub = (uint64_t)(bound / (4 * S)) + 1;
      slot[k] = {so, bound, (uint32_t)seg, 2 * (uint32_t)nseg + 3, (uint32_t)sub, (uint32_t)cta};
      so += bound;
      seg += 2 * (uint64_t)nseg + 3;
      sub += nsub;
      cta += (nsub + JGPU_HUFF_OWN - 1) / JGPU_HUFF_OWN + 1;   /* carry slots of the sync kernel's CTAs */
    }
    slot[m] = {so, 0, (uint32_t)seg, 0, (uint32_t)sub, (uint32_t)cta};
    if (so / 4 >= 0xffffffffll || sub >= 0xffffffffull) {
      release_fronts();
      return jgpu_fail("jgpu_decode_jpegs: batch too large for 32-bit stream offsets; split it");
    }
  }
  const size_t n_sub = slot[m].subseq0, n_seg = slot[m].seg0, n_cta = slot[m].cta0;
  if (ctx->hz_stream.reserve((size_t)slot[m].stream_off + 16) || ctx->dz_stream.reserve((size_t)slot[m].stream_off + 16) ||
      ctx->hz_files.reserve(sizeof(jgpu_huff_file) * m) || ctx->dz_files.reserve(sizeof(jgpu_huff_file) * m) ||
      ctx->hz_tables.reserve(sizeof(jgpu_huff_table) * JGPU_HUFF_TABLES * m) ||
      ctx->dz_tables.reserve(sizeof(jgpu_huff_table) * JGPU_HUFF_TABLES * m) ||
      ctx->hz_segs.reserve(4 * n_seg + 16) || ctx->dz_segs.reserve(4 * n_seg + 16) ||
      ctx->hz_status.reserve(4 * (size_t)m + 16) || ctx->dz_status.reserve(4 * (size_t)m + 16) ||
      ctx->dz_sub[0].reserve(4 * n_sub + 16) || ctx->dz_sub[1].reserve(4 * n_sub + 16) ||
      ctx->dz_sub[2].reserve(4 * n_sub + 16) || ctx->dz_sub[3].reserve(4 * n_sub + 16) ||
      ctx->dz_carry[0].reserve(4 * n_cta + 16) || ctx->dz_carry[1].reserve(4 * n_cta + 16) ||
      ctx->d_coef.reserve((size_t)coef_off * 2 + 256) || ctx->d_qtabs.reserve(qtabs.size() * 2) ||
      (!device_out && ctx->d_rgb.reserve((size_t)rgb_off + 256))) {
    release_fronts();
    return EXIT_FAILURE;
  }
  unsigned char *h_stream = (unsigned char *)ctx->hz_stream.ptr;
  jgpu_huff_file *h_files = (jgpu_huff_file *)ctx->hz_files.ptr;
  jgpu_huff_table *h_tables = (jgpu_huff_table *)ctx->hz_tables.ptr;
  uint32_t *h_segs = (uint32_t *)ctx->hz_segs.ptr;
  uint32_t *h_status = (uint32_t *)ctx->hz_status.ptr;
  int16_t *d_coef = (int16_t *)ctx->d_coef.ptr;
  uint16_t *d_qtabs = (uint16_t *)ctx->d_qtabs.ptr;
  /* JGPU_JPEGS_DEVICE_OUT: the caller's buffer is device memory and the kernels write into it */
  uint8_t *d_rgb = device_out ? h_rgb : (uint8_t *)ctx->d_rgb.ptr;

  t_setup = since();
  /* ---- host threads: unstuff, cut at the restart markers, build the tables ------- */
  std::vector<char> on_gpu(m, 0);
  for (int k = 0; k < m; k++) items[ok[k]].tasks_left.store(1, std::memory_order_relaxed);
  std::atomic<int> next_file{0};
  auto worker = [&]() {
    for (;;) {
      const int k = next_file.fetch_add(1, std::memory_order_relaxed);
      if (k >= m) return;
      JpegItem &it = items[ok[k]];
      jgpu_huff_file &f = h_files[k];
      memset(&f, 0, sizeof(f));
      const char *why = nullptr;
      const long long bytes = jfront_huff_prepare(it.front, S, h_stream + slot[k].stream_off, slot[k].stream_cap,
                                                  h_segs + slot[k].seg0, (int)slot[k].seg_cap,
                                                  h_tables + (size_t)JGPU_HUFF_TABLES * k, &f, &why);
      if (bytes < 0) {
        memset(&f, 0, sizeof(f));   /* no subsequences: the kernels skip it */
      } else {
        on_gpu[k] = 1;
        for (int p = 0; p < it.desc.ncomps; p++) {
          f.hblocks[p] = it.lay.plane[p].hblocks;
          f.plane_off[p] = it.desc.coef_off + it.lay.plane[p].coef_off;
        }
        jgpu_huff_file_finish(&f);
      }
      f.word0 = (uint32_t)(slot[k].stream_off / 4);
      f.subseq0 = slot[k].subseq0;
      f.seg0 = slot[k].seg0;
      f.cta0 = slot[k].cta0;
      f.table0 = (uint32_t)(JGPU_HUFF_TABLES * k);
      f.status_slot = k;
      it.tasks_left.store(0, std::memory_order_release);
    }
  };
  std::vector<std::thread> pool;
  const int nworkers = std::min(nthreads, m);
  for (int t = 0; t < nworkers; t++) pool.emplace_back(worker);
  auto join_all = [&]() {
    for (std::thread &t : pool) t.join();
    pool.clear();
  };

  /* ---- GPU pipeline ------------------------------------------------------------------ */
  auto gpu_part = [&]() -> int {
    CU_TRY(cudaMemcpyAsync(d_qtabs, qtabs.data(), qtabs.size() * 2, cudaMemcpyHostToDevice, ctx->streams[0]));
    CU_TRY(cudaEventRecord(ctx->events[0], ctx->streams[0]));
    for (int s = 1; s < kHostStreams; s++) CU_TRY(cudaStreamWaitEvent(ctx->streams[s], ctx->events[0], 0));
    /* Groups of files.  The first one is small (48 MB of coefficients: two 4K files) so that the GPU -- and with
     * host output the read-back, the slowest stage -- gets going early.  The later ones are sized by the sync
     * kernel's grid: its CTAs run rounds of a serial chain and take about as long whatever their number, so a
     * grid of 1.3 waves costs two.  Groups two and three take the files that fill half a wave of resident CTAs,
     * the rest one wave (4K files: 2, 3, 3, then 6 or 7 per group), within 512 MB of coefficients.  Measured on
     * 128 / 32 4K files, pixels left on the device (profiles/r2_notes.md 14): one wave 15.2 / 6.0 ms, two waves
     * 15.7 / 6.4, half waves throughout 17.5 / 5.65. */
    const int64_t first_bytes = 48ll << 20, max_bytes = 512ll << 20;
    long wave = (long)ctx->sm_count * kHuffSyncCtasPerSm;   /* CTAs of k_huff_sync resident at once */
    if (getenv("JGPU_HUFF_WAVE_PER_SM")) wave = (long)ctx->sm_count * std::max(1, atoi(getenv("JGPU_HUFF_WAVE_PER_SM")));
    int i0 = 0, chunk = 0;
    const bool blocks_at_end = device_out && !getenv("JGPU_BLOCKS_PER_GROUP");

    /* JGPU_HUFF_WAVES / JGPU_HUFF_WAVE_PER_SM: experiment knobs, waves of sync CTAs per group and CTAs per SM in a wave */
    const int max_waves = getenv("JGPU_HUFF_WAVES") ? std::max(1, atoi(getenv("JGPU_HUFF_WAVES"))) : 1;
    /* JGPU_TRACE: device-side times of every group (events on its stream) */
    std::vector<cudaEvent_t> tev;
    std::vector<int> tfiles;
    auto mark = [&](cudaStream_t s) {
      if (!trace) return;
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, s);
      tev.push_back(e);
    };
    while (i0 < m) {
      int i1 = i0;
      int64_t acc = 0;
      long ctas = 0;
      const long target = chunk == 0 ? 0 : chunk < 3 ? wave / 2 : max_waves * wave;
      /* (a group's file count is gridDim.y of the entropy kernels: at most 65535) */
      while (i1 < m && i1 - i0 < 65535) {
        JpegItem &it = items[ok[i1]];
        while (it.tasks_left.load(std::memory_order_acquire) > 0) std::this_thread::yield();
        const int64_t bytes = plan->layouts[i1].coef_len * 2;
        const long c = on_gpu[i1] ? (long)((h_files[i1].n_subseq + JGPU_HUFF_OWN - 1) / JGPU_HUFF_OWN) : 0;
        if (i1 > i0 && (acc + bytes > max_bytes || (target > 0 ? ctas + c > target : acc + bytes > first_bytes))) break;
        acc += bytes;
        ctas += c;
        i1++;
      }
      HuffLaunch l;
      for (int k = i0; k < i1; k++) {
        JpegItem &it = items[ok[k]];
        while (it.tasks_left.load(std::memory_order_acquire) > 0) std::this_thread::yield();
        l.max_subseq = std::max(l.max_subseq, (int)h_files[k].n_subseq);
        l.max_dc_jobs = std::max(l.max_dc_jobs, (int)h_files[k].n_seg * h_files[k].ncomps);
        l.max_dc_chain = std::max(l.max_dc_chain, h_files[k].mcus_per_seg * h_files[k].hs[0] * h_files[k].vs[0]);
        info[ok[k]].tasks = on_gpu[k] ? (int32_t)h_files[k].n_subseq : 1;
      }
      cudaStream_t st = ctx->streams[chunk % kHostStreams];
      mark(st);
      tfiles.push_back(i1 - i0);
      /* The group's input goes up on the upload stream; its kernels wait for that copy alone.  (With the
       * copies on the groups' own streams the three streams fell into step -- all copying, then all
       * computing -- and the SMs sat idle a third of the time, profiles/r2_notes.md 14.) */
      if (!ctx->up_stream) CU_TRY(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
      while ((int)ctx->up_events.size() <= chunk) {
        cudaEvent_t e;
        CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->up_events.push_back(e);
      }
      cudaStream_t up = ctx->up_stream;
      /* each file's stream is a little shorter than its slot (headers, stuffing and markers are
       * gone): large files are copied slot by slot, many small ones in one go, gaps included */
      if (i1 - i0 > 16) {
        CU_TRY(cudaMemcpyAsync((unsigned char *)ctx->dz_stream.ptr + slot[i0].stream_off, h_stream + slot[i0].stream_off,
                               (size_t)(slot[i1].stream_off - slot[i0].stream_off), cudaMemcpyHostToDevice, up));
      } else {
        for (int k = i0; k < i1; k++) {
          if (!on_gpu[k]) continue;
          const size_t bytes = (size_t)h_files[k].n_subseq * 4 * S + 16;
          CU_TRY(cudaMemcpyAsync((unsigned char *)ctx->dz_stream.ptr + slot[k].stream_off, h_stream + slot[k].stream_off,
                                 bytes, cudaMemcpyHostToDevice, up));
        }
      }
      CU_TRY(cudaMemcpyAsync((jgpu_huff_file *)ctx->dz_files.ptr + i0, h_files + i0, sizeof(jgpu_huff_file) * (i1 - i0),
                             cudaMemcpyHostToDevice, up));
      CU_TRY(cudaMemcpyAsync((jgpu_huff_table *)ctx->dz_tables.ptr + (size_t)JGPU_HUFF_TABLES * i0,
                             h_tables + (size_t)JGPU_HUFF_TABLES * i0,
                             sizeof(jgpu_huff_table) * JGPU_HUFF_TABLES * (i1 - i0), cudaMemcpyHostToDevice, up));
      CU_TRY(cudaMemcpyAsync((uint32_t *)ctx->dz_segs.ptr + slot[i0].seg0, h_segs + slot[i0].seg0,
                             4 * (size_t)(slot[i1].seg0 - slot[i0].seg0), cudaMemcpyHostToDevice, up));
      CU_TRY(cudaEventRecord(ctx->up_events[chunk], up));
      CU_TRY(cudaStreamWaitEvent(st, ctx->up_events[chunk], 0));
      mark(st);
      /* (zeroing on the upload stream instead, ahead of the group's turn, holds up the copies behind it:
       * 16.4 against 15.8 ms for 128 files) */
      CU_TRY(cudaMemsetAsync(d_coef + descs[i0].coef_off, 0, (size_t)acc, st));
      mark(st);
      {
        /* scratch of the DC pass, one buffer per stream (groups on one stream run in order) */
        Buffer &dc = ctx->dz_dc[chunk % kHostStreams];
        l.dc_partial_stride = huff_dc_partial_ints(l.max_dc_jobs, l.max_dc_chain);
        if (dc.cap < l.dc_partial_stride * 4 * (size_t)(i1 - i0)) {
          /* growing means freeing: wait for the groups that may still use the old buffer */
          CU_TRY(cudaStreamSynchronize(st));
          if (dc.reserve(l.dc_partial_stride * 4 * (size_t)(i1 - i0) + 16)) return EXIT_FAILURE;
        }
        l.d_dc_partial = (int *)dc.ptr;
      }
      l.n_files = i1 - i0;
      l.carry0 = slot[i0].cta0;
      l.n_carry = slot[i1].cta0 - slot[i0].cta0;
      l.status0 = i0;
      l.d_files = (const jgpu_huff_file *)ctx->dz_files.ptr + i0;
      l.d_stream = (const uint32_t *)ctx->dz_stream.ptr;
      l.d_tables = (const jgpu_huff_table *)ctx->dz_tables.ptr;
      l.d_seg_first = (const uint32_t *)ctx->dz_segs.ptr;
      l.d_state = (uint32_t *)ctx->dz_sub[0].ptr;
      l.d_nslots = (uint32_t *)ctx->dz_sub[1].ptr;
      l.d_slots = (uint32_t *)ctx->dz_sub[2].ptr;
      l.d_segid = (uint32_t *)ctx->dz_sub[3].ptr;
      l.d_carry[0] = (uint32_t *)ctx->dz_carry[0].ptr;
      l.d_carry[1] = (uint32_t *)ctx->dz_carry[1].ptr;
      l.d_status = (uint32_t *)ctx->dz_status.ptr;
      l.d_coef = d_coef;
      if (huff_launch(l, st)) return EXIT_FAILURE;
      mark(st);
      /* The block decoder takes whole SMs (one CTA of 227 KB each): launched between the entropy kernels of
       * three streams it waits for SMs to drain and they wait for it (0.4-1.2 ms per group against 0.11 ms
       * alone, profiles/r2_notes.md 14).  When the pixels stay on the device nothing waits for a group's
       * pixels, so one launch at the end decodes the blocks of all groups; with the read-back in the
       * pipeline every group is decoded as soon as its coefficients are there. */
      if (!blocks_at_end &&
          plan_run_range(plan, i0, i1, d_coef, d_qtabs, m, yuv_out ? nullptr : d_rgb, yuv_out ? d_rgb : nullptr, st)) {
        return EXIT_FAILURE;
      }
      mark(st);
      CU_TRY(cudaMemcpyAsync(h_status + i0, (uint32_t *)ctx->dz_status.ptr + i0, 4 * (size_t)(i1 - i0),
                             cudaMemcpyDeviceToHost, st));
      if (!device_out) {
        const int64_t lo = info[ok[i0]].rgb_off;
        const int64_t hi = info[ok[i1 - 1]].rgb_off + info[ok[i1 - 1]].rgb_len;
        CU_TRY(cudaMemcpyAsync(h_rgb + lo, d_rgb + lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost, st));
      }
      i0 = i1;
      chunk++;
      if (i0 >= m) t_prepared = since();
    }
    if (blocks_at_end) {
      for (int s = 1; s < kHostStreams; s++) {
        CU_TRY(cudaEventRecord(ctx->events[s], ctx->streams[s]));
        CU_TRY(cudaStreamWaitEvent(ctx->streams[0], ctx->events[s], 0));
      }
      if (plan_run_range(plan, 0, m, d_coef, d_qtabs, m, yuv_out ? nullptr : d_rgb, yuv_out ? d_rgb : nullptr,
                         ctx->streams[0])) {
        return EXIT_FAILURE;
      }
    }
    t_enqueued = since();
    for (int s = 0; s < kHostStreams; s++) CU_TRY(cudaStreamSynchronize(ctx->streams[s]));
    CU_TRY(cudaStreamSynchronize(ctx->up_stream));
    t_synced = since();
    for (size_t g = 0; g + 4 < tev.size() + 1 && g / 5 < tfiles.size(); g += 5) {
      float a = 0, b = 0, c = 0, d = 0, e = 0;
      cudaEventElapsedTime(&a, tev[0], tev[g]);
      cudaEventElapsedTime(&b, tev[g], tev[g + 1]);
      cudaEventElapsedTime(&c, tev[g + 1], tev[g + 2]);
      cudaEventElapsedTime(&d, tev[g + 2], tev[g + 3]);
      cudaEventElapsedTime(&e, tev[g + 3], tev[g + 4]);
      fprintf(stderr, "  group %2d (%2d files, stream %d): starts %.2f ms, copy in %.2f, zero %.2f, entropy %.2f, blocks %.2f\n",
              (int)(g / 5), tfiles[g / 5], (int)((g / 5) % kHostStreams), a, b, c, d, e);
    }
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
    return EXIT_SUCCESS;
  };
  int rc = gpu_part();
  join_all();
  if (rc != EXIT_SUCCESS) {
    cudaDeviceSynchronize();
    release_fronts();
    return rc;
  }
  if (trace) {
    fprintf(stderr, "jgpu_decode_jpegs[gpu entropy]: %d files, setup %.2f ms, last file prepared %.2f, enqueued %.2f, "
            "streams drained %.2f\n", m, t_setup, t_prepared, t_enqueued, t_synced);
  }

  /* ---- files the GPU decoder left to the sequential reader ---------------------------- */
  auto leftovers = [&]() -> int {   /* returns through CU_TRY on a CUDA failure; the caller cleans up */
  for (int k = 0; k < m; k++) {
    if (on_gpu[k] && h_status[k] == 0) continue;
    JpegItem &it = items[ok[k]];
    info[ok[k]].tasks = 1;
    if (ctx->hz_coef.reserve((size_t)it.lay.coef_len * 2 + 256)) return EXIT_FAILURE;
    int16_t *h_coef = (int16_t *)ctx->hz_coef.ptr;
    memset(h_coef, 0, (size_t)it.lay.coef_len * 2);
    image img;
    bind_image(&img, it.desc, it.lay, h_coef);
    {
      /* exactly what the host-thread path does with this file: restart intervals cut at their
       * markers when those are in place, else the sequential reader and its error reports */
      jfront_segments segs;
      const char *err = nullptr;
      int frc;
      if (jfront_find_segments(it.front, &segs) == 0) {
        frc = jfront_decode_segments(it.front, &img, JPEG_DECODE_QUANT, &segs, 0, segs.nseg, &err);
        jfront_segments_free(&segs);
      } else {
        frc = v.decode_image(it.front, &img, JPEG_DECODE_QUANT);
        if (frc) err = "Error decoding scan";
      }
      if (frc) {
        info[ok[k]].status = 1;
        info[ok[k]].message = err ? err : "Error decoding scan";
        rc = EXIT_FAILURE;
        continue;
      }
    }
    cudaStream_t st = ctx->streams[0];
    CU_TRY(cudaMemcpyAsync(d_coef + it.desc.coef_off, h_coef, (size_t)it.lay.coef_len * 2, cudaMemcpyHostToDevice, st));
    if (plan_run_range(plan, k, k + 1, d_coef, d_qtabs, m, yuv_out ? nullptr : d_rgb, yuv_out ? d_rgb : nullptr, st)) {
      return EXIT_FAILURE;
    }
    if (!device_out) {
      CU_TRY(cudaMemcpyAsync(h_rgb + info[ok[k]].rgb_off, d_rgb + info[ok[k]].rgb_off, (size_t)info[ok[k]].rgb_len,
                             cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaStreamSynchronize(st));
  }
  return EXIT_SUCCESS;
  };
  if (leftovers() != EXIT_SUCCESS) {
    release_fronts();
    return EXIT_FAILURE;
  }
  if (m != n) rc = EXIT_FAILURE;
  release_fronts();
  if (rc != EXIT_SUCCESS) {
    for (int i = 0; i < n; i++) {
      if (info[i].status) {
        jgpu_fail("file %d: %s", i, info[i].message ? info[i].message : "rejected");
        break;
      }
    }
  }
  return rc;
}

extern "C" int jgpu_decode_jpegs_ex(jgpu_ctx *ctx, const jgpu_jpeg *files, int n, int nthreads, unsigned flags,
                                    uint8_t *h_rgb, int64_t rgb_cap, jgpu_jpeg_info *info) {
  if (!ctx || !files || n <= 0 || !h_rgb || !info) return jgpu_fail("jgpu_decode_jpegs: bad arguments");
  const bool device_out = (flags & JGPU_JPEGS_DEVICE_OUT) != 0;
  const bool yuv_out = (flags & JGPU_JPEGS_OUT_YUV) != 0;
  unsigned entropy = flags & ~(JGPU_JPEGS_DEVICE_OUT | JGPU_JPEGS_OUT_YUV);
  if (entropy == JGPU_ENTROPY_AUTO) {
    const char *env = getenv("JGPU_ENTROPY");
    entropy = env && !strcmp(env, "cpu") && !device_out && !yuv_out ? JGPU_ENTROPY_CPU : JGPU_ENTROPY_GPU;
  }
  if (entropy == JGPU_ENTROPY_CPU && !device_out && !yuv_out) {
    return decode_jpegs_cpu_entropy(ctx, files, n, nthreads, h_rgb, rgb_cap, info);
  }
  if (entropy == JGPU_ENTROPY_GPU) {
    if (device_out && (reinterpret_cast<uintptr_t>(h_rgb) & 255)) {
      return jgpu_fail("jgpu_decode_jpegs_ex: a device output buffer must be 256-byte aligned");
    }
    return decode_jpegs_gpu_entropy(ctx, files, n, nthreads, h_rgb, rgb_cap, info, device_out, yuv_out);
  }
  return jgpu_fail("jgpu_decode_jpegs_ex: unsupported flags %u", flags);
}

extern "C" int jgpu_decode_jpegs(jgpu_ctx *ctx, const jgpu_jpeg *files, int n, int nthreads,
                                 uint8_t *h_rgb, int64_t rgb_cap, jgpu_jpeg_info *info) {
  return jgpu_decode_jpegs_ex(ctx, files, n, nthreads, JGPU_ENTROPY_AUTO, h_rgb, rgb_cap, info);
}
