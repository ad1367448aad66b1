/* jgpu_fused.cu — the fused coefficient -> RGB kernel (placeholder while the
 * generic path is brought up; fused_available() gates it). */
#include "jgpu_launch.h"

namespace jgpu {

bool fused_available() { return false; }
cudaError_t fused_configure(int) { return cudaSuccess; }
int fused_plan_build(FusedPlan &, const jgpu_image_desc *, const jgpu_layout *, const int *, int,
                     unsigned, int) {
  return 1;
}
void fused_plan_release(FusedPlan &) {}
cudaError_t fused_plan_launch(const FusedPlan &, int, int, const int16_t *, const uint16_t *,
                              uint8_t *, uint8_t *, cudaStream_t) {
  return cudaErrorNotSupported;
}

}  // namespace jgpu
