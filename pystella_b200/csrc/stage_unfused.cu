// Element-wise scalar-sector LSRK stage reading a precomputed Laplacian: the
// reference's own stage kernel (pystella/step.py:473-491 built on
// pystella/elementwise.py:164-235), kept for user loops written exactly like the
// reference's (`derivs(...)` then `stepper(...)`).  In place, like the reference.
#include "common.cuh"

namespace psb {

template <typename T, int NF> struct UnfusedParams {
  T* f;
  T* dfdt;
  T* kf;
  T* kdfdt;
  const T* lap;
  int64_t field_stride, lap_stride;
  int Nx, Ny, Nz, h;
  int x_begin, x_end;
  T dt, A, B, a, hubble;
  const double* a_dev;
  const double* hubble_dev;
  int nterms[NF];
  T coef[NF][PSB_MAX_TERMS];
  int8_t power[NF][PSB_MAX_TERMS][NF];
};

template <typename T, int NF>
__global__ void __launch_bounds__(256)
    unfused_stage_kernel(const __grid_constant__ UnfusedParams<T, NF> p) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int i = p.x_begin + blockIdx.z;
  if (k >= p.Nz || j >= p.Ny || i >= p.x_end) return;
  const int h = p.h;
  const int64_t Pz = p.Nz + 2 * h, Py = p.Ny + 2 * h;
  const int64_t o = ((int64_t)(i + h) * Py + (j + h)) * Pz + (k + h);
  const int64_t ol = ((int64_t)i * p.Ny + j) * p.Nz + k;

  T a = p.a, hubble = p.hubble;
  if (p.a_dev) a = (T)(*p.a_dev);
  if (p.hubble_dev) hubble = (T)(*p.hubble_dev);
  const T two_h = (T)2 * hubble;
  const T a_sq = a * a;

  T fv[NF], dv[NF], lv[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    fv[f] = p.f[o + f * p.field_stride];
    dv[f] = p.dfdt[o + f * p.field_stride];
    lv[f] = p.lap[ol + f * p.lap_stride];
  }
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    T dV = (T)0;
    for (int t = 0; t < p.nterms[f]; ++t) {
      T term = p.coef[f][t];
#pragma unroll
      for (int m = 0; m < NF; ++m)
        for (int e = 0; e < p.power[f][t][m]; ++e) term = term * fv[m];
      dV = dV + term;
    }
    const T rhs_f = dv[f];
    const T rhs_d = (lv[f] - two_h * dv[f]) - a_sq * dV;
    const int64_t of = o + f * p.field_stride;
    T kfv, kdv;
    if (p.A != (T)0) {
      kfv = p.A * p.kf[of] + p.dt * rhs_f;
      kdv = p.A * p.kdfdt[of] + p.dt * rhs_d;
    } else {
      kfv = p.dt * rhs_f;
      kdv = p.dt * rhs_d;
    }
    p.kf[of] = kfv;
    p.kdfdt[of] = kdv;
    p.f[of] = fv[f] + p.B * kfv;
    p.dfdt[of] = dv[f] + p.B * kdv;
  }
}

template <typename T, int NF>
static int launch_unfused(const psb_scalar_stage_args* a, cudaStream_t stream) {
  UnfusedParams<T, NF> p;
  p.f = static_cast<T*>(a->f_out);
  p.dfdt = static_cast<T*>(a->dfdt);
  p.kf = static_cast<T*>(a->kf);
  p.kdfdt = static_cast<T*>(a->kdfdt);
  p.lap = static_cast<const T*>(a->lap_f);
  p.field_stride = a->field_stride;
  p.lap_stride = a->lap_stride;
  p.Nx = (int)a->Nx;
  p.Ny = (int)a->Ny;
  p.Nz = (int)a->Nz;
  p.h = a->h;
  p.x_begin = (int)a->x_begin;
  p.x_end = (int)a->x_end;
  p.dt = (T)a->dt;
  p.A = (T)a->A;
  p.B = (T)a->B;
  p.a = (T)a->a;
  p.hubble = (T)a->hubble;
  p.a_dev = a->a_dev;
  p.hubble_dev = a->hubble_dev;
  for (int f = 0; f < NF; ++f) {
    p.nterms[f] = a->nterms[f];
    for (int t = 0; t < PSB_MAX_TERMS; ++t) {
      p.coef[f][t] = (T)a->coef[f][t];
      for (int m = 0; m < NF; ++m) p.power[f][t][m] = a->power[f][t][m];
    }
  }
  dim3 block(64, 4, 1);
  dim3 grid((unsigned)((a->Nz + 63) / 64), (unsigned)((a->Ny + 3) / 4),
            (unsigned)(a->x_end - a->x_begin));
  unfused_stage_kernel<T, NF><<<grid, block, 0, stream>>>(p);
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int dispatch_unfused(const psb_scalar_stage_args* a, cudaStream_t s) {
  switch (a->nscalars) {
    case 1: return launch_unfused<T, 1>(a, s);
    case 2: return launch_unfused<T, 2>(a, s);
    case 3: return launch_unfused<T, 3>(a, s);
    case 4: return launch_unfused<T, 4>(a, s);
  }
  set_error("psb_scalar_stage_unfused: nscalars=%d out of range", a->nscalars);
  return 1;
}

}  // namespace psb

extern "C" int psb_scalar_stage_unfused(const psb_scalar_stage_args* a, void* stream) {
  using namespace psb;
  PSB_REQUIRE(a != nullptr, "psb_scalar_stage_unfused: null args");
  PSB_REQUIRE(a->f_out && a->dfdt && a->kf && a->kdfdt && a->lap_f,
              "psb_scalar_stage_unfused: null array pointer");
  PSB_REQUIRE(a->f_in == nullptr || a->f_in == a->f_out,
              "psb_scalar_stage_unfused: the unfused stage updates f in place");
  PSB_REQUIRE(a->Nx > 0 && a->Ny > 0 && a->Nz > 0, "psb_scalar_stage_unfused: empty grid");
  PSB_REQUIRE(a->h >= 0, "psb_scalar_stage_unfused: negative halo");
  PSB_REQUIRE(a->x_begin >= 0 && a->x_end <= a->Nx && a->x_begin <= a->x_end,
              "psb_scalar_stage_unfused: bad x range");
  PSB_REQUIRE(a->nscalars >= 1 && a->nscalars <= PSB_MAX_SCALARS,
              "psb_scalar_stage_unfused: nscalars=%d out of range", a->nscalars);
  for (int f = 0; f < a->nscalars; ++f) {
    PSB_REQUIRE(a->nterms[f] >= 0 && a->nterms[f] <= PSB_MAX_TERMS,
                "psb_scalar_stage_unfused: too many potential terms");
    for (int t = 0; t < a->nterms[f]; ++t)
      for (int m = 0; m < a->nscalars; ++m)
        PSB_REQUIRE(a->power[f][t][m] >= 0,
                    "psb_scalar_stage_unfused: negative power of field %d (dV/df_%d, term %d)",
                    m, f, t);
  }
  if (a->x_begin == a->x_end) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == PSB_F64) return dispatch_unfused<double>(a, s);
  if (a->dtype == PSB_F32) return dispatch_unfused<float>(a, s);
  set_error("psb_scalar_stage_unfused: unknown dtype %d", a->dtype);
  return 1;
}
