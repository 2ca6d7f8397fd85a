// Classical RK stage with storage copies for the scalar sector (element-wise):
// RungeKuttaStepper.make_steps + step_statements (pystella/step.py:202-411).
#include "common.cuh"

namespace psb {

template <typename T, int NF> struct RKParams {
  T* f;
  T* dfdt;
  const T* lap;
  int Nx, Ny, Nz, h, ncopies, q, nst;
  int dst[2], ncomb[2], src[2][2];
  T coef[2][2], dtc[2];
  T dt, a, hubble;
  int nterms[NF];
  T pcoef[NF][PSB_MAX_TERMS];
  int8_t power[NF][PSB_MAX_TERMS][NF];
};

template <typename T, int NF>
__global__ void __launch_bounds__(256)
    rk_stage_kernel(const __grid_constant__ RKParams<T, NF> p) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int i = blockIdx.z;
  if (k >= p.Nz || j >= p.Ny) return;
  const int h = p.h;
  const int64_t Pz = p.Nz + 2 * h, Py = p.Ny + 2 * h, Px = p.Nx + 2 * h;
  const int64_t vol_p = Px * Py * Pz, vol = (int64_t)p.Nx * p.Ny * p.Nz;
  const int64_t copy_stride = (int64_t)NF * vol_p;
  const int64_t o = ((int64_t)(i + h) * Py + (j + h)) * Pz + (k + h);
  const int64_t ou = ((int64_t)i * p.Ny + j) * p.Nz + k;
  const T two_h = (T)2 * p.hubble, a_sq = p.a * p.a;

  T fq[NF], dq[NF];
#pragma unroll
  for (int n = 0; n < NF; ++n) {
    fq[n] = p.f[p.q * copy_stride + n * vol_p + o];
    dq[n] = p.dfdt[p.q * copy_stride + n * vol_p + o];
  }
#pragma unroll
  for (int n = 0; n < NF; ++n) {
    T dV = (T)0;
    for (int t = 0; t < p.nterms[n]; ++t) {
      T term = p.pcoef[n][t];
#pragma unroll
      for (int m = 0; m < NF; ++m)
        for (int e = 0; e < p.power[n][t][m]; ++e) term = term * fq[m];
      dV = dV + term;
    }
    const T rhs_f = dq[n];
    const T rhs_d = (p.lap[n * vol + ou] - two_h * dq[n]) - a_sq * dV;
    const int64_t on = n * vol_p + o;
    T newf[2], newd[2];
    for (int s = 0; s < p.nst; ++s) {
      T vf = (T)0, vd = (T)0;
      for (int m = 0; m < p.ncomb[s]; ++m) {
        const int64_t os = p.src[s][m] * copy_stride + on;
        vf = (m == 0) ? p.coef[s][m] * p.f[os] : vf + p.coef[s][m] * p.f[os];
        vd = (m == 0) ? p.coef[s][m] * p.dfdt[os] : vd + p.coef[s][m] * p.dfdt[os];
      }
      newf[s] = vf + (p.dtc[s] * p.dt) * rhs_f;
      newd[s] = vd + (p.dtc[s] * p.dt) * rhs_d;
    }
    for (int s = 0; s < p.nst; ++s) {
      p.f[p.dst[s] * copy_stride + on] = newf[s];
      p.dfdt[p.dst[s] * copy_stride + on] = newd[s];
    }
  }
}

template <typename T, int NF>
static int rk_launch(const psb_rk_stage_args* a, cudaStream_t s) {
  RKParams<T, NF> p;
  p.f = static_cast<T*>(a->f);
  p.dfdt = static_cast<T*>(a->dfdt);
  p.lap = static_cast<const T*>(a->lap_f);
  p.Nx = (int)a->Nx;
  p.Ny = (int)a->Ny;
  p.Nz = (int)a->Nz;
  p.h = a->h;
  p.ncopies = a->ncopies;
  p.q = a->q_read;
  p.nst = a->nstatements;
  for (int st = 0; st < 2; ++st) {
    p.dst[st] = a->dst[st];
    p.ncomb[st] = a->ncomb[st];
    p.dtc[st] = (T)a->dt_coef[st];
    for (int m = 0; m < 2; ++m) {
      p.src[st][m] = a->src[st][m];
      p.coef[st][m] = (T)a->coef[st][m];
    }
  }
  p.dt = (T)a->dt;
  p.a = (T)a->a;
  p.hubble = (T)a->hubble;
  for (int n = 0; n < NF; ++n) {
    p.nterms[n] = a->nterms[n];
    for (int t = 0; t < PSB_MAX_TERMS; ++t) {
      p.pcoef[n][t] = (T)a->pcoef[n][t];
      for (int m = 0; m < NF; ++m) p.power[n][t][m] = a->power[n][t][m];
    }
  }
  dim3 block(64, 4, 1);
  dim3 grid((unsigned)((a->Nz + 63) / 64), (unsigned)((a->Ny + 3) / 4), (unsigned)a->Nx);
  rk_stage_kernel<T, NF><<<grid, block, 0, s>>>(p);
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int rk_typed(const psb_rk_stage_args* a, cudaStream_t s) {
  switch (a->nscalars) {
    case 1: return rk_launch<T, 1>(a, s);
    case 2: return rk_launch<T, 2>(a, s);
    case 3: return rk_launch<T, 3>(a, s);
    case 4: return rk_launch<T, 4>(a, s);
  }
  set_error("psb_scalar_rk_stage: nscalars=%d out of range", a->nscalars);
  return 1;
}

}  // namespace psb

extern "C" int psb_scalar_rk_stage(const psb_rk_stage_args* a, void* stream) {
  using namespace psb;
  PSB_REQUIRE(a != nullptr, "psb_scalar_rk_stage: null args");
  PSB_REQUIRE(a->f && a->dfdt && a->lap_f, "psb_scalar_rk_stage: null array pointer");
  PSB_REQUIRE(a->Nx > 0 && a->Ny > 0 && a->Nz > 0, "psb_scalar_rk_stage: empty grid");
  PSB_REQUIRE(a->ncopies >= 1 && a->ncopies <= 4, "psb_scalar_rk_stage: bad ncopies");
  PSB_REQUIRE(a->q_read >= 0 && a->q_read < a->ncopies, "psb_scalar_rk_stage: bad q_read");
  PSB_REQUIRE(a->nstatements >= 1 && a->nstatements <= 2,
              "psb_scalar_rk_stage: 1 or 2 statements per stage");
  for (int s = 0; s < a->nstatements; ++s) {
    PSB_REQUIRE(a->dst[s] >= 0 && a->dst[s] < a->ncopies, "psb_scalar_rk_stage: bad dst");
    PSB_REQUIRE(a->ncomb[s] >= 1 && a->ncomb[s] <= 2, "psb_scalar_rk_stage: bad ncomb");
    for (int m = 0; m < a->ncomb[s]; ++m)
      PSB_REQUIRE(a->src[s][m] >= 0 && a->src[s][m] < a->ncopies,
                  "psb_scalar_rk_stage: bad src");
  }
  PSB_REQUIRE(a->nscalars >= 1 && a->nscalars <= PSB_MAX_SCALARS,
              "psb_scalar_rk_stage: nscalars=%d out of range", a->nscalars);
  PSB_REQUIRE(a->h >= 0, "psb_scalar_rk_stage: negative halo");
  for (int f = 0; f < a->nscalars; ++f) {
    PSB_REQUIRE(a->nterms[f] >= 0 && a->nterms[f] <= PSB_MAX_TERMS,
                "psb_scalar_rk_stage: too many potential terms");
    for (int t = 0; t < a->nterms[f]; ++t)
      for (int m = 0; m < a->nscalars; ++m)
        PSB_REQUIRE(a->power[f][t][m] >= 0,
                    "psb_scalar_rk_stage: negative power of field %d (dV/df_%d, term %d)", m,
                    f, t);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == PSB_F64) return rk_typed<double>(a, s);
  if (a->dtype == PSB_F32) return rk_typed<float>(a, s);
  set_error("psb_scalar_rk_stage: unknown dtype %d", a->dtype);
  return 1;
}
