// C-ABI entry points of the fused stage; the kernels live in fused_stage_impl.cuh
// and are instantiated per dtype in fused_stage_f64.cu / fused_stage_f32.cu.
#include "common.cuh"
#include "stage_tail.cuh"

namespace psb {
thread_local psb_fused_plan* t_fused_plan = nullptr;
int fused_dispatch_f64(const psb_scalar_stage_args* a, cudaStream_t s, int64_t* count);
int fused_dispatch_f32(const psb_scalar_stage_args* a, cudaStream_t s, int64_t* count);

static int validate(const psb_scalar_stage_args* a) {
  PSB_REQUIRE(a != nullptr, "psb_scalar_stage_fused: null args");
  PSB_REQUIRE(a->f_in && a->f_out && a->dfdt && a->kf && a->kdfdt,
              "psb_scalar_stage_fused: null array pointer");
  PSB_REQUIRE(a->f_in != a->f_out,
              "psb_scalar_stage_fused: f_out must not alias f_in (ping-pong)");
  PSB_REQUIRE(a->Nx > 0 && a->Ny > 0 && a->Nz > 0, "psb_scalar_stage_fused: empty grid");
  PSB_REQUIRE(a->h >= 1 && a->h <= 4, "psb_scalar_stage_fused: h=%d out of range 1..4", a->h);
  PSB_REQUIRE(a->Nx >= a->h && a->Ny >= a->h && a->Nz >= a->h,
              "psb_scalar_stage_fused: grid smaller than the halo");
  PSB_REQUIRE(a->x_begin >= 0 && a->x_end <= a->Nx && a->x_begin <= a->x_end,
              "psb_scalar_stage_fused: bad x range [%lld, %lld)", (long long)a->x_begin,
              (long long)a->x_end);
  PSB_REQUIRE(a->x2_begin >= 0 && a->x2_end <= a->Nx && a->x2_begin <= a->x2_end,
              "psb_scalar_stage_fused: bad second x range");
  PSB_REQUIRE(a->x2_end == a->x2_begin || a->x2_begin >= a->x_end || a->x2_end <= a->x_begin,
              "psb_scalar_stage_fused: the two x ranges overlap");
  PSB_REQUIRE(a->nscalars >= 1 && a->nscalars <= PSB_MAX_SCALARS,
              "psb_scalar_stage_fused: nscalars=%d out of range", a->nscalars);
  for (int f = 0; f < a->nscalars; ++f)
    PSB_REQUIRE(a->nterms[f] >= 0 && a->nterms[f] <= PSB_MAX_TERMS,
                "psb_scalar_stage_fused: too many potential terms");
  PSB_REQUIRE(a->nvterms >= 0 && a->nvterms <= PSB_MAX_TERMS,
              "psb_scalar_stage_fused: too many terms in V");
  for (int f = 0; f < a->nscalars; ++f)
    for (int t = 0; t < a->nterms[f]; ++t)
      for (int m = 0; m < a->nscalars; ++m)
        PSB_REQUIRE(a->power[f][t][m] >= 0,
                    "psb_scalar_stage_fused: negative power %d of field %d (dV/df_%d, "
                    "term %d)", a->power[f][t][m], m, f, t);
  for (int t = 0; t < a->nvterms; ++t)
    for (int m = 0; m < a->nscalars; ++m)
      PSB_REQUIRE(a->vpower[t][m] >= 0,
                  "psb_scalar_stage_fused: negative power %d of field %d (V, term %d)",
                  a->vpower[t][m], m, t);
  for (int d = 0; d < 3; ++d)
    PSB_REQUIRE(a->dx[d] > 0.0 && a->dx[d] < 1e300,
                "psb_scalar_stage_fused: grid spacing dx[%d]=%g must be positive and finite",
                d, a->dx[d]);
  {
    const int64_t lim = (int64_t)1 << 30;   // plane/row indices are 32-bit in the kernels
    PSB_REQUIRE(a->Nx < lim && a->Ny < lim && a->Nz < lim,
                "psb_scalar_stage_fused: an extent exceeds 2^30 sites");
    const int64_t vol = (a->Nx + 2 * a->h) * (a->Ny + 2 * a->h) * (a->Nz + 2 * a->h);
    PSB_REQUIRE(a->nscalars == 1 || a->field_stride >= vol,
                "psb_scalar_stage_fused: field_stride %lld smaller than one padded field "
                "(%lld elements)", (long long)a->field_stride, (long long)vol);
  }
  PSB_REQUIRE(!(a->wrap_x && (a->peer_lo_f_out || a->peer_hi_f_out)),
              "psb_scalar_stage_fused: wrap_x and peer halos are mutually exclusive");
  PSB_REQUIRE(a->dtype == PSB_F64 || a->dtype == PSB_F32,
              "psb_scalar_stage_fused: unknown dtype %d", a->dtype);
  return 0;
}
}  // namespace psb

extern "C" int psb_scalar_stage_fused(const psb_scalar_stage_args* a, void* stream) {
  using namespace psb;
  if (int rc = validate(a)) return rc;
  if (a->x_begin == a->x_end && a->x2_begin == a->x2_end) {
    PSB_REQUIRE(a->tail == nullptr || a->tail->mode == 0,
                "psb_scalar_stage_fused: a stage tail needs a non-empty launch");
    return 0;
  }
  if (a->tail != nullptr && a->tail->mode != 0) {
    int64_t nblocks = 0;
    const int rc0 = a->dtype == PSB_F64 ? fused_dispatch_f64(a, nullptr, &nblocks)
                                        : fused_dispatch_f32(a, nullptr, &nblocks);
    if (rc0) return rc0;
    if (int rc = validate_tail(a->tail, nblocks, a->energy_partials != nullptr,
                               "psb_scalar_stage_fused"))
      return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return a->dtype == PSB_F64 ? fused_dispatch_f64(a, s, nullptr)
                             : fused_dispatch_f32(a, s, nullptr);
}

extern "C" int psb_scalar_stage_fused_plan(const psb_scalar_stage_args* a,
                                           psb_fused_plan* plan) {
  using namespace psb;
  PSB_REQUIRE(plan != nullptr, "psb_scalar_stage_fused_plan: null plan");
  if (int rc = validate(a)) return rc;
  PSB_REQUIRE(a->x_begin != a->x_end || a->x2_begin != a->x2_end,
              "psb_scalar_stage_fused_plan: empty x range");
  if (a->tail != nullptr && a->tail->mode != 0) {
    int64_t nblocks = 0;
    const int rc0 = a->dtype == PSB_F64 ? fused_dispatch_f64(a, nullptr, &nblocks)
                                        : fused_dispatch_f32(a, nullptr, &nblocks);
    if (rc0) return rc0;
    if (int rc = validate_tail(a->tail, nblocks, a->energy_partials != nullptr,
                               "psb_scalar_stage_fused_plan"))
      return rc;
  }
  t_fused_plan = plan;
  const int rc = a->dtype == PSB_F64 ? fused_dispatch_f64(a, nullptr, nullptr)
                                     : fused_dispatch_f32(a, nullptr, nullptr);
  t_fused_plan = nullptr;
  return rc;
}

extern "C" int64_t psb_scalar_stage_energy_slots(const psb_scalar_stage_args* a) {
  using namespace psb;
  if (validate(a) != 0) return -1;
  int64_t count = 0;
  int rc = a->dtype == PSB_F64 ? fused_dispatch_f64(a, nullptr, &count)
                               : fused_dispatch_f32(a, nullptr, &count);
  return rc == 0 ? count : -1;
}
