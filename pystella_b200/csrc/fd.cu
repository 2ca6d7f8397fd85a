// Stand-alone finite-difference kernels: the lap / grad / grad_lap / pd{x,y,z} /
// pd*_incr kernels of FiniteDifferencer (pystella/derivs.py:291-337).
//
// 2.5-D streaming like the fused stage: each CTA owns a (BY x BZ) column and
// marches along x with a register queue for the x-neighbours and the centre
// plane staged in shared memory for y/z neighbours.  Outputs are unpadded.
#include <math.h>

#include <stdlib.h>

#include "common.cuh"

namespace psb {

template <typename T> struct FDParams {
  const T* fx;
  T* lap;
  T* pdx;
  T* pdy;
  T* pdz;
  int Nx, Ny, Nz, halo, xchunk, variant;
  T inv_dx2[3];
  T inv_dx[3];
};

// MODE bits: 1 = lap, 2 = pdx, 4 = pdy, 8 = pdz ; ACC: add into first derivatives
template <typename T, int H, int MODE, bool ACC, int BY, int BZ>
__global__ void __launch_bounds__(BY* BZ)
    fd_kernel(const __grid_constant__ FDParams<T> p) {
  constexpr int W = BZ + 2 * H, ROWS = BY + 2 * H;
  __shared__ T tile[2][ROWS][W];

  const int tz = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * BZ + tz;
  const int z0 = blockIdx.x * BZ, y0 = blockIdx.y * BY;
  const int xs = blockIdx.z * p.xchunk;
  const int xe = min(xs + p.xchunk, p.Nx);
  const int j = y0 + ty, k = z0 + tz;
  const bool active = (j < p.Ny) && (k < p.Nz);
  const int g = p.halo;  // padding of fx (>= H)
  const int64_t Pz = p.Nz + 2 * g, Py = p.Ny + 2 * g;
  const int64_t plane = Py * Pz;

  // own column pointer at padded plane index 0
  const int jj = min(j, p.Ny - 1), kk = min(k, p.Nz - 1);
  const T* col = p.fx + (int64_t)(jj + g) * Pz + (kk + g);

  T q[2 * H + 1];
#pragma unroll
  for (int s = 0; s < 2 * H; ++s)
    q[s + 1] = col[(int64_t)(xs + g - H + s) * plane];

  for (int i = xs; i < xe; ++i) {
#pragma unroll
    for (int s = 0; s < 2 * H; ++s) q[s] = q[s + 1];
    q[2 * H] = col[(int64_t)(i + g + H) * plane];

    const int b = i & 1;
    // stage the centre plane (with y/z halos) in shared memory
    const T* src = p.fx + (int64_t)(i + g) * plane;
    for (int e = tid; e < ROWS * W; e += BY * BZ) {
      const int r = e / W, c = e % W;
      const int gy = y0 + r + g - H, gz = z0 + c + g - H;
      T val = (T)0;
      if (gy < Py && gz < Pz) val = src[(int64_t)gy * Pz + gz];
      tile[b][r][c] = val;
    }
    __syncthreads();

    if (active) {
      const int r = ty + H, c = tz + H;
      const int64_t o = ((int64_t)i * p.Ny + j) * p.Nz + k;
      if constexpr (MODE & 1) {
        const T cc = q[H];
        T sx = (T)lap_coef<H>(0) * cc, sy = sx, sz = sx;
#pragma unroll
        for (int s = 1; s <= H; ++s) {
          const T cs = (T)lap_coef<H>(s);
          sx = sx + cs * q[H + s];
          sx = sx + cs * q[H - s];
          sy = sy + cs * tile[b][r + s][c];
          sy = sy + cs * tile[b][r - s][c];
          sz = sz + cs * tile[b][r][c + s];
          sz = sz + cs * tile[b][r][c - s];
        }
        p.lap[o] = sx * p.inv_dx2[0] + sy * p.inv_dx2[1] + sz * p.inv_dx2[2];
      }
      if constexpr (MODE & 2) {
        T s1 = (T)0;
#pragma unroll
        for (int s = 1; s <= H; ++s) {
          const T cs = (T)grad_coef<H>(s);
          s1 = (s == 1) ? cs * q[H + s] : s1 + cs * q[H + s];
          s1 = s1 + (-cs) * q[H - s];
        }
        const T val = s1 * p.inv_dx[0];
        p.pdx[o] = ACC ? p.pdx[o] + val : val;
      }
      if constexpr (MODE & 4) {
        T s1 = (T)0;
#pragma unroll
        for (int s = 1; s <= H; ++s) {
          const T cs = (T)grad_coef<H>(s);
          s1 = (s == 1) ? cs * tile[b][r + s][c] : s1 + cs * tile[b][r + s][c];
          s1 = s1 + (-cs) * tile[b][r - s][c];
        }
        const T val = s1 * p.inv_dx[1];
        p.pdy[o] = ACC ? p.pdy[o] + val : val;
      }
      if constexpr (MODE & 8) {
        T s1 = (T)0;
#pragma unroll
        for (int s = 1; s <= H; ++s) {
          const T cs = (T)grad_coef<H>(s);
          s1 = (s == 1) ? cs * tile[b][r][c + s] : s1 + cs * tile[b][r][c + s];
          s1 = s1 + (-cs) * tile[b][r][c - s];
        }
        const T val = s1 * p.inv_dx[2];
        p.pdz[o] = ACC ? p.pdz[o] + val : val;
      }
    }
    // double-buffered tile: one barrier per plane is enough
  }
}

}  // namespace psb
#include "fd_ws.cuh"
namespace psb {

template <typename T, int H, int MODE, bool ACC>
static int launch_fd(const FDParams<T>& p, cudaStream_t stream) {
  if constexpr (!ACC && (MODE == 1 || MODE == 14 || MODE == 15)) {
    if (p.variant != 1) {
      const int rc = launch_fd_ws<T, H, MODE>(p, stream);
      if (rc >= 0) return rc;
    }
  }
  constexpr int BY = 8, BZ = 64;
  FDParams<T> q = p;
  q.xchunk = 32;
  if (q.xchunk > p.Nx) q.xchunk = p.Nx;
  dim3 grid((p.Nz + BZ - 1) / BZ, (p.Ny + BY - 1) / BY, (p.Nx + q.xchunk - 1) / q.xchunk);
  dim3 block(BZ, BY, 1);
  fd_kernel<T, H, MODE, ACC, BY, BZ><<<grid, block, 0, stream>>>(q);
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T, int H>
static int dispatch_mode(const FDParams<T>& p, int mode, bool acc, cudaStream_t s) {
  if (acc) {
    switch (mode) {
      case 2: return launch_fd<T, H, 2, true>(p, s);
      case 4: return launch_fd<T, H, 4, true>(p, s);
      case 8: return launch_fd<T, H, 8, true>(p, s);
    }
    set_error("psb_finite_difference: accumulate needs exactly one of pdx/pdy/pdz");
    return 1;
  }
  switch (mode) {
    case 1: return launch_fd<T, H, 1, false>(p, s);
    case 2: return launch_fd<T, H, 2, false>(p, s);
    case 4: return launch_fd<T, H, 4, false>(p, s);
    case 8: return launch_fd<T, H, 8, false>(p, s);
    case 14: return launch_fd<T, H, 14, false>(p, s);
    case 15: return launch_fd<T, H, 15, false>(p, s);
  }
  set_error("psb_finite_difference: unsupported output combination (mask %d); valid: "
            "lap, pdx, pdy, pdz, pdx+pdy+pdz, lap+pdx+pdy+pdz", mode);
  return 1;
}

template <typename T>
static int fd_typed(const void* fx, void* lap, void* pdx, void* pdy, void* pdz,
                    int64_t Nx, int64_t Ny, int64_t Nz, int h, int halo,
                    const double dx[3], int accumulate, int variant, cudaStream_t s) {
  FDParams<T> p;
  p.variant = variant;
  p.fx = static_cast<const T*>(fx);
  p.lap = static_cast<T*>(lap);
  p.pdx = static_cast<T*>(pdx);
  p.pdy = static_cast<T*>(pdy);
  p.pdz = static_cast<T*>(pdz);
  p.Nx = (int)Nx;
  p.Ny = (int)Ny;
  p.Nz = (int)Nz;
  p.halo = halo;
  p.xchunk = 0;
  for (int d = 0; d < 3; ++d) {
    p.inv_dx2[d] = (T)pow(dx[d], -2.0);  // dxi**-2  (derivs.py:295-296)
    p.inv_dx[d] = (T)(1.0 / dx[d]);      // 1/dxi    (derivs.py:293-294)
  }
  const int mode = (lap ? 1 : 0) | (pdx ? 2 : 0) | (pdy ? 4 : 0) | (pdz ? 8 : 0);
  switch (h) {
    case 1: return dispatch_mode<T, 1>(p, mode, accumulate != 0, s);
    case 2: return dispatch_mode<T, 2>(p, mode, accumulate != 0, s);
    case 3: return dispatch_mode<T, 3>(p, mode, accumulate != 0, s);
    case 4: return dispatch_mode<T, 4>(p, mode, accumulate != 0, s);
  }
  set_error("psb_finite_difference: h=%d out of range 1..4", h);
  return 1;
}

}  // namespace psb

extern "C" int psb_finite_difference_v(const void* fx, void* lap, void* pdx, void* pdy,
                                       void* pdz, int dtype, int64_t Nx, int64_t Ny,
                                       int64_t Nz, int h, int halo, const double dx[3],
                                       int accumulate, int variant, void* stream) {
  using namespace psb;
  PSB_REQUIRE(fx != nullptr, "psb_finite_difference: null input");
  PSB_REQUIRE(Nx > 0 && Ny > 0 && Nz > 0, "psb_finite_difference: empty grid");
  PSB_REQUIRE(halo >= h, "psb_finite_difference: halo %d smaller than stencil radius %d",
              halo, h);
  PSB_REQUIRE(variant == 0 || variant == 1, "psb_finite_difference: unknown variant %d",
              variant);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == PSB_F64)
    return fd_typed<double>(fx, lap, pdx, pdy, pdz, Nx, Ny, Nz, h, halo, dx, accumulate,
                            variant, s);
  if (dtype == PSB_F32)
    return fd_typed<float>(fx, lap, pdx, pdy, pdz, Nx, Ny, Nz, h, halo, dx, accumulate,
                           variant, s);
  set_error("psb_finite_difference: unknown dtype %d", dtype);
  return 1;
}

extern "C" int psb_finite_difference(const void* fx, void* lap, void* pdx, void* pdy,
                                     void* pdz, int dtype, int64_t Nx, int64_t Ny,
                                     int64_t Nz, int h, int halo, const double dx[3],
                                     int accumulate, void* stream) {
  return psb_finite_difference_v(fx, lap, pdx, pdy, pdz, dtype, Nx, Ny, Nz, h, halo, dx,
                                 accumulate, 0, stream);
}
