// fp64 instantiations of the fused stage (separate TU: parallel compilation).
#include "fused_stage_impl.cuh"

namespace psb {
int fused_dispatch_f64(const psb_scalar_stage_args* a, cudaStream_t s, int64_t* count) {
  return dispatch_h<double>(a, s, count);
}
}  // namespace psb
