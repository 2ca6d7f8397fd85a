// Error channel, ABI version and the host-side TMA descriptor encoder.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "stage_tail.cuh"

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

int validate_tail(const psb_stage_tail* a, int64_t nblocks, bool has_partials,
                  const char* who) {
  if (a == nullptr || a->mode == 0) return 0;
  PSB_REQUIRE((a->mode & ~(PSB_TAIL_FOLD | PSB_TAIL_EXPAND | PSB_TAIL_PEER)) == 0,
              "%s: unknown stage-tail mode bits 0x%x", who, a->mode);
  PSB_REQUIRE(a->nscalars >= 0 && a->nscalars <= PSB_MAX_SCALARS,
              "%s: stage tail nscalars=%d out of range", who, a->nscalars);
  if (a->mode & PSB_TAIL_FOLD) {
    PSB_REQUIRE(has_partials, "%s: PSB_TAIL_FOLD needs energy_partials", who);
    PSB_REQUIRE(a->nscalars >= 1, "%s: PSB_TAIL_FOLD needs nscalars >= 1", who);
  }
  if (!has_partials)
    PSB_REQUIRE(a->rows != nullptr && a->nrows >= nblocks,
                "%s: stage tail needs %lld marker rows (has %lld)", who, (long long)nblocks,
                (long long)(a->rows ? a->nrows : 0));
  if (a->nscalars > 0)
    PSB_REQUIRE(a->sums != nullptr, "%s: stage tail without a sums buffer", who);
  if (a->mode & PSB_TAIL_EXPAND) {
    PSB_REQUIRE(a->nscalars >= 1 && a->state && a->energies,
                "%s: PSB_TAIL_EXPAND needs nscalars >= 1, state and energies", who);
    PSB_REQUIRE(a->grid_size > 0 && a->mpl > 0, "%s: stage tail: bad grid_size/mpl", who);
  }
  if (a->mode & PSB_TAIL_PEER) {
    PSB_REQUIRE(a->peers.nranks >= 1 && a->peers.nranks <= 64 && a->peers.rank >= 0 &&
                    a->peers.rank < a->peers.nranks,
                "%s: stage tail: bad rank/nranks", who);
    PSB_REQUIRE(a->peers.peer_slots && a->peers.peer_flags && a->peers.epoch,
                "%s: stage tail: null peer tables", who);
  }
  return 0;
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
