// Error channel, ABI version and the host-side TMA descriptor encoder.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace psb {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t,
                                  void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault,
                                &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

int encode_field_tensor_map(CUtensorMap* map, int dtype, const void* base,
                            int64_t nfields, int64_t field_stride_elems, int64_t Px,
                            int64_t Py, int64_t Pz, int box_x, int box_y,
                            int box_z) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return 1;
  const int64_t es = (dtype == PSB_F64) ? 8 : 4;
  if (reinterpret_cast<uintptr_t>(base) % 16) return 2;
  if ((Pz * es) % 16 || (Py * Pz * es) % 16 || (field_stride_elems * es) % 16)
    return 3;
  if ((box_z * es) % 16 || box_z > 256 || box_y > 256 || box_x > 256) return 4;
  cuuint64_t gdim[4] = {(cuuint64_t)Pz, (cuuint64_t)Py, (cuuint64_t)Px,
                        (cuuint64_t)nfields};
  cuuint64_t gstride[3] = {(cuuint64_t)(Pz * es), (cuuint64_t)(Py * Pz * es),
                           (cuuint64_t)(field_stride_elems * es)};
  if (nfields == 1) gstride[2] = (cuuint64_t)(Px * Py * Pz * es);
  cuuint32_t box[4] = {(cuuint32_t)box_z, (cuuint32_t)box_y, (cuuint32_t)box_x, 1};
  cuuint32_t estride[4] = {1, 1, 1, 1};
  CUresult r = fn(map,
                  dtype == PSB_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                                   : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  4, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 5;
}

}  // namespace psb

extern "C" {

const char* psb_last_error(void) { return psb::g_error; }
int psb_abi_version(void) { return PSB_ABI_VERSION; }
int psb_strict_fp(void) {
#ifdef PSB_STRICT_FP
  return 1;
#else
  return 0;
#endif
}

}  // extern "C"
