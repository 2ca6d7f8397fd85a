// fp32 instantiations of the fused stage (separate TU: parallel compilation).
#include "fused_stage_impl.cuh"

namespace psb {
int fused_dispatch_f32(const psb_scalar_stage_args* a, cudaStream_t s, int64_t* count) {
  return dispatch_h<float>(a, s, count);
}
}  // namespace psb
