/* Launch entry points of jgpu_kernels.cu, called by jgpu_runtime.cu. */
#ifndef JGPU_LAUNCH_H
#define JGPU_LAUNCH_H

#include <cuda_runtime.h>
#include "jgpu_device_types.h"
#include "jpeg_gpu_b200.h"

namespace jgpu {

constexpr int kPairThreads = 128;   /* block pairs per CTA, generic IDCT kernel */
constexpr int kColourThreads = 256; /* 4-pixel groups per CTA, colour kernel   */

cudaError_t launch_coef_to_planes(const PlaneSeg *segs, const PairWork *work,
                                  int ncta, const int16_t *coef,
                                  const uint16_t *qtabs, uint8_t *planes,
                                  cudaStream_t stream);

cudaError_t launch_planes_to_rgb(const ColourImage *imgs, const ColourWork *work,
                                 int ncta, const uint8_t *planes, uint8_t *rgb,
                                 cudaStream_t stream);

/* ---- PACK expansion (jgpu_unpack.cu) -------------------------------------- */

constexpr int kUnpackThreads = 256; /* blocks per CTA */

cudaError_t launch_unpack(const UnpackSeg *segs, const UnpackWork *work, int ncta,
                          const uint16_t *pack, const int64_t *pack_off, const int32_t *index,
                          int16_t *coef, cudaStream_t stream);

/* ---- fused path (jgpu_fused.cu) ------------------------------------------ */

struct FusedPlan {
  void *impl = nullptr; /* FusedPlanImpl, jgpu_fused.cu */
};

bool fused_available();
cudaError_t fused_configure(int device);
/* modes[i] is the FusedMode of image i.  Returns 0 or 1 (jgpu_fail). */
int fused_plan_build(FusedPlan &fp, const jgpu_image_desc *descs, const jgpu_layout *layouts,
                     const int *modes, int n, unsigned flags, int sm_count);
void fused_plan_release(FusedPlan &fp);
int fused_plan_launches(const FusedPlan &fp);
/* Enqueues images [i0, i1).  Returns 0 or 1 (jgpu_fail). */
int fused_plan_launch(FusedPlan &fp, int i0, int i1, const int16_t *coef, const uint16_t *qtabs,
                      int n_sets, uint8_t *rgb, cudaStream_t stream);

cudaError_t launch_prep_qtabs(const uint16_t *qtabs, uint32_t *qint, int n_tables, uint32_t *wide_flag,
                              cudaStream_t stream);

/* ---- fused path, one MCU column per thread (jgpu_mcu.cu) -------------------- */
/* Same plan interface; `flags` selects pixels (JGPU_OUT_RGB) or planes (JGPU_OUT_YUV), not both. */
cudaError_t mcu_configure(int device);
/* whether the kernel in use (k_tk, or k_mcu with JGPU_KERNEL=mcu) has this FusedMode */
bool mcu_has_mode(int mode);
int mcu_plan_build(FusedPlan &fp, const jgpu_image_desc *descs, const jgpu_layout *layouts,
                   const int *modes, int n, unsigned flags, int sm_count);
void mcu_plan_release(FusedPlan &fp);
int mcu_plan_launches(const FusedPlan &fp);
int mcu_plan_launch(FusedPlan &fp, int i0, int i1, const int16_t *coef, const uint16_t *qtabs,
                    int n_sets, uint8_t *rgb, uint8_t *yuv, cudaStream_t stream);

}  // namespace jgpu
#endif
