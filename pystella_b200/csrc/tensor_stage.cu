// Element-wise low-storage RK stage of the tensor-perturbation (GW) sector:
// the stage kernel the reference generates from TensorPerturbationSector.rhs_dict
// (pystella/sectors.py:187-204, pystella/step.py:473-491).  One thread per site
// updates all six packed components; the gradient products are formed once per site.
#include "common.cuh"

namespace psb {

template <typename T> struct TensorParams {
  T* hij;
  T* dhijdt;
  T* khij;
  T* kdhijdt;
  const T* lap;
  const T* dfdx;
  int Nx, Ny, Nz, h, ns;
  T dt, A, B, hubble, coef;
  const double* hubble_dev;
};

template <typename T>
__global__ void __launch_bounds__(256)
    tensor_stage_kernel(const __grid_constant__ TensorParams<T> p) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int i = blockIdx.z;
  if (k >= p.Nz || j >= p.Ny) return;
  const int h = p.h;
  const int64_t Pz = p.Nz + 2 * h, Py = p.Ny + 2 * h, Px = p.Nx + 2 * h;
  const int64_t vol_p = Px * Py * Pz, vol = (int64_t)p.Nx * p.Ny * p.Nz;
  const int64_t o = ((int64_t)(i + h) * Py + (j + h)) * Pz + (k + h);
  const int64_t ou = ((int64_t)i * p.Ny + j) * p.Nz + k;
  T hubble = p.hubble;
  if (p.hubble_dev) hubble = (T)(*p.hubble_dev);
  const T two_h = (T)2 * hubble;

  // S_ij = sum_n d_i f_n d_j f_n, packed (1,1),(1,2),(1,3),(2,2),(2,3),(3,3)
  T S[6] = {0, 0, 0, 0, 0, 0};
  for (int n = 0; n < p.ns; ++n) {
    const T gx = p.dfdx[(n * 3 + 0) * vol + ou];
    const T gy = p.dfdx[(n * 3 + 1) * vol + ou];
    const T gz = p.dfdx[(n * 3 + 2) * vol + ou];
    // first field initialises, later fields accumulate (sum(...) in the reference)
    if (n == 0) {
      S[0] = gx * gx; S[1] = gx * gy; S[2] = gx * gz;
      S[3] = gy * gy; S[4] = gy * gz; S[5] = gz * gz;
    } else {
      S[0] = S[0] + gx * gx; S[1] = S[1] + gx * gy; S[2] = S[2] + gx * gz;
      S[3] = S[3] + gy * gy; S[4] = S[4] + gy * gz; S[5] = S[5] + gz * gz;
    }
  }
  const bool read_k = (p.A != (T)0);
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const int64_t oc = o + c * vol_p;
    const T d_old = p.dhijdt[oc];
    const T rhs_d = (p.lap[c * vol + ou] - two_h * d_old) + p.coef * S[c];
    T kh, kd;
    if (read_k) {
      kh = p.A * p.khij[oc] + p.dt * d_old;
      kd = p.A * p.kdhijdt[oc] + p.dt * rhs_d;
    } else {
      kh = p.dt * d_old;
      kd = p.dt * rhs_d;
    }
    p.khij[oc] = kh;
    p.kdhijdt[oc] = kd;
    p.hij[oc] = p.hij[oc] + p.B * kh;
    p.dhijdt[oc] = d_old + p.B * kd;
  }
}

template <typename T>
static int tensor_typed(const psb_tensor_stage_args* a, cudaStream_t s) {
  TensorParams<T> p;
  p.hij = static_cast<T*>(a->hij);
  p.dhijdt = static_cast<T*>(a->dhijdt);
  p.khij = static_cast<T*>(a->khij);
  p.kdhijdt = static_cast<T*>(a->kdhijdt);
  p.lap = static_cast<const T*>(a->lap_hij);
  p.dfdx = static_cast<const T*>(a->dfdx);
  p.Nx = (int)a->Nx;
  p.Ny = (int)a->Ny;
  p.Nz = (int)a->Nz;
  p.h = a->h;
  p.ns = a->nscalars;
  p.dt = (T)a->dt;
  p.A = (T)a->A;
  p.B = (T)a->B;
  p.hubble = (T)a->hubble;
  p.coef = (T)a->source_coef;
  p.hubble_dev = a->hubble_dev;
  dim3 block(64, 4, 1);
  dim3 grid((unsigned)((a->Nz + 63) / 64), (unsigned)((a->Ny + 3) / 4), (unsigned)a->Nx);
  tensor_stage_kernel<T><<<grid, block, 0, s>>>(p);
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace psb

extern "C" int psb_tensor_stage_unfused(const psb_tensor_stage_args* a, void* stream) {
  using namespace psb;
  PSB_REQUIRE(a != nullptr, "psb_tensor_stage_unfused: null args");
  PSB_REQUIRE(a->hij && a->dhijdt && a->khij && a->kdhijdt && a->lap_hij && a->dfdx,
              "psb_tensor_stage_unfused: null array pointer");
  PSB_REQUIRE(a->Nx > 0 && a->Ny > 0 && a->Nz > 0, "psb_tensor_stage_unfused: empty grid");
  PSB_REQUIRE(a->nscalars >= 1, "psb_tensor_stage_unfused: needs at least one source field");
  PSB_REQUIRE(a->h >= 0, "psb_tensor_stage_unfused: negative halo");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == PSB_F64) return tensor_typed<double>(a, s);
  if (a->dtype == PSB_F32) return tensor_typed<float>(a, s);
  set_error("psb_tensor_stage_unfused: unknown dtype %d", a->dtype);
  return 1;
}
