// EXPERIMENTAL (opt-in: PSB_FD_PIPELINE=1): the stand-alone finite-difference kernels
// through the warp-specialised TMA pipeline of the fused stage (fused_stage_ws.cuh): one
// producer lane streams the padded planes of `fx` through an (H+1+P)-slot shared-memory
// ring, consumer warps keep the x-neighbours in a register queue, read the y/z neighbours
// from the centre plane, and store lap / grad straight from registers as 16-byte vectors.
// Same arithmetic (term order, scaling) as fd_kernel, so the strict build stays bit-exact
// against the oracle.  Layouts the pipeline cannot take (halo != h, odd h or extents,
// accumulate mode, single first derivatives) stay on fd_kernel.
// Written after round 1's GPU budget was spent: compiles, not yet run on hardware.
#pragma once
#include "fused_stage_impl.cuh"   // pulls in fused_stage_ws.cuh in the right order

namespace psb {

template <typename T> struct FDParams;   // defined in fd.cu

template <typename T, int H, int BY, int TZ, int P, int VEC> struct FdWsCfg {
  using C = WsCfg<T, H, 1, BY, TZ, P, VEC>;
  static constexpr int RF = C::RF;
  static constexpr int SMEM_BYTES = RF * C::FPLANE + 2 * RF * 8 + 128;
};

// MODE bits: 1 = lap, 2 = pdx, 4 = pdy, 8 = pdz (1, 14 and 15 are instantiated)
template <typename T, int H, int MODE, int BY, int TZ, int P, int MINB, int VEC>
__global__ void __launch_bounds__(BY* TZ + 32, MINB)
    fd_ws_kernel(const __grid_constant__ FDParams<T> p,
                 const __grid_constant__ CUtensorMap map_f) {
  using C = WsCfg<T, H, 1, BY, TZ, P, VEC>;
  constexpr int GV = C::GV, W = C::W, HL = C::HL, RF = C::RF;
  using V = Vec<T, VEC>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  unsigned char* fring = base;
  uint64_t* full_f = reinterpret_cast<uint64_t*>(fring + RF * C::FPLANE);
  uint64_t* empty_f = full_f + RF;

  const int tid = threadIdx.x;
  const int z0 = blockIdx.x * C::BZ, y0 = blockIdx.y * BY;
  const int xs = blockIdx.z * p.xchunk;
  const int xe = min(xs + p.xchunk, p.Nx);
  const int ncomp = xe - xs;
  const int nsteps = ncomp + 2 * H;

  if (tid == 0) {
    for (int s = 0; s < RF; ++s) {
      mbar_init(&full_f[s], 1);
      mbar_init(&empty_f[s], C::NCW);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (tid >= BY * TZ) {
    if (tid == BY * TZ) {   // producer lane
      tma_prefetch_desc(&map_f);
      for (int m = 0; m < nsteps; ++m) {
        const int slot = m % RF;
        if (m >= RF) mbar_wait(&empty_f[slot], ((m / RF) - 1) & 1);
        mbar_expect_tx(&full_f[slot], C::FPLANE_RAW);
        // plane m holds interior x index xs - H + m, i.e. padded index xs + m (halo == H)
        tma_load_4d(fring + slot * C::FPLANE, &map_f, &full_f[slot], z0 + C::ZF, y0, xs + m,
                    0);
      }
    }
    return;
  }

  const int tz = tid % TZ, ty = tid / TZ;
  const int lane = tid & 31;
  const int j = y0 + ty, k = z0 + tz * VEC;
  const bool active = (j < p.Ny) && (k < p.Nz);
  const int own_off = (ty + H) * W + HL + tz * VEC;

  V q[2 * H + 1];
#pragma unroll
  for (int s = 0; s < 2 * H + 1; ++s)
#pragma unroll
    for (int v = 0; v < VEC; ++v) q[s].v[v] = (T)0;

  int64_t o = ((int64_t)xs * p.Ny + j) * p.Nz + k;   // un-padded output index of plane xs
  const int64_t oplane = (int64_t)p.Ny * p.Nz;

  for (int n = 0; n < nsteps; ++n) {
    const int slot = n % RF;
    mbar_wait(&full_f[slot], (n / RF) & 1);
#pragma unroll
    for (int s = 0; s < 2 * H; ++s) q[s] = q[s + 1];
    q[2 * H] = ld_pieces<T, VEC, GV>(
        reinterpret_cast<const T*>(fring + slot * C::FPLANE) + own_off);

    if (n >= 2 * H) {
      if (active) {
        const T* pl = reinterpret_cast<const T*>(fring + ((n - H) % RF) * C::FPLANE);
        constexpr int NVZ = (VEC + 2 * HL) / GV;
        T zrow[NVZ * GV];
#pragma unroll
        for (int w = 0; w < NVZ; ++w) {
          Vec<T, GV> t = ld_vec<T, GV>(pl + own_off - HL + w * GV);
#pragma unroll
          for (int v = 0; v < GV; ++v) zrow[w * GV + v] = t.v[v];
        }
        V yn[2 * H + 1];
#pragma unroll
        for (int s = 1; s <= H; ++s) {
          yn[H + s] = ld_pieces<T, VEC, GV>(pl + own_off + s * W);
          yn[H - s] = ld_pieces<T, VEC, GV>(pl + own_off - s * W);
        }
        if constexpr (MODE & 1) {
          V out;
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const T cc = q[H].v[v];
            T sx = (T)lap_coef<H>(0) * cc, sy = sx, sz = sx;
#pragma unroll
            for (int s = 1; s <= H; ++s) {
              const T cs = (T)lap_coef<H>(s);
              sx = sx + cs * q[H + s].v[v];
              sx = sx + cs * q[H - s].v[v];
              sy = sy + cs * yn[H + s].v[v];
              sy = sy + cs * yn[H - s].v[v];
              sz = sz + cs * zrow[HL + v + s];
              sz = sz + cs * zrow[HL + v - s];
            }
            out.v[v] = sx * p.inv_dx2[0] + sy * p.inv_dx2[1] + sz * p.inv_dx2[2];
          }
          st_pieces<T, VEC, C::GVO, true>(p.lap + o, out);
        }
        if constexpr (MODE & 2) {
          V out;
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            T s1 = (T)0;
#pragma unroll
            for (int s = 1; s <= H; ++s) {
              const T cs = (T)grad_coef<H>(s);
              s1 = (s == 1) ? cs * q[H + s].v[v] : s1 + cs * q[H + s].v[v];
              s1 = s1 + (-cs) * q[H - s].v[v];
            }
            out.v[v] = s1 * p.inv_dx[0];
          }
          st_pieces<T, VEC, C::GVO, true>(p.pdx + o, out);
        }
        if constexpr (MODE & 4) {
          V out;
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            T s1 = (T)0;
#pragma unroll
            for (int s = 1; s <= H; ++s) {
              const T cs = (T)grad_coef<H>(s);
              s1 = (s == 1) ? cs * yn[H + s].v[v] : s1 + cs * yn[H + s].v[v];
              s1 = s1 + (-cs) * yn[H - s].v[v];
            }
            out.v[v] = s1 * p.inv_dx[1];
          }
          st_pieces<T, VEC, C::GVO, true>(p.pdy + o, out);
        }
        if constexpr (MODE & 8) {
          V out;
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            T s1 = (T)0;
#pragma unroll
            for (int s = 1; s <= H; ++s) {
              const T cs = (T)grad_coef<H>(s);
              s1 = (s == 1) ? cs * zrow[HL + v + s] : s1 + cs * zrow[HL + v + s];
              s1 = s1 + (-cs) * zrow[HL + v - s];
            }
            out.v[v] = s1 * p.inv_dx[2];
          }
          st_pieces<T, VEC, C::GVO, true>(p.pdz + o, out);
        }
      }
      o += oplane;
    }
    if (n >= H) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_f[(n - H) % RF]);
    }
  }
}

#ifndef __CUDACC_RTC__
// returns 0 ok, >0 error, -1 = layout not eligible (caller uses fd_kernel)
template <typename T, int H, int MODE>
static int launch_fd_ws(const FDParams<T>& p0, cudaStream_t stream) {
  constexpr int VEC = 2, BY = 4, P = 3, MINB = 2;
  constexpr int TZ = 64;
  using C = WsCfg<T, H, 1, BY, TZ, P, VEC>;
  using F = FdWsCfg<T, H, BY, TZ, P, VEC>;
  if constexpr (H % 2 != 0 || F::SMEM_BYTES > SMEM_LIMIT) {
    return -1;
  } else {
    FDParams<T> p = p0;
    if (p.halo != H || p.Nz % VEC != 0) return -1;
    const int64_t Px = p.Nx + 2 * H, Py = p.Ny + 2 * H, Pz = p.Nz + 2 * H;
    const int es = (int)sizeof(T);
    auto misaligned = [&](const void* q) {
      return q != nullptr && reinterpret_cast<uintptr_t>(q) % (VEC * es) != 0;
    };
    if (misaligned(p.lap) || misaligned(p.pdx) || misaligned(p.pdy) || misaligned(p.pdz))
      return -1;
    CUtensorMap map;
    if (encode_field_tensor_map(&map, es == 8 ? PSB_F64 : PSB_F32, p.fx, 1, Px * Py * Pz, Px,
                                Py, Pz, 1, C::ROWS, C::W) != 0)
      return -1;
    p.xchunk = 16;
    if (p.xchunk > p.Nx) p.xchunk = p.Nx;
    dim3 grid((unsigned)((p.Nz + C::BZ - 1) / C::BZ), (unsigned)((p.Ny + BY - 1) / BY),
              (unsigned)((p.Nx + p.xchunk - 1) / p.xchunk));
    auto kern = fd_ws_kernel<T, H, MODE, BY, TZ, P, MINB, VEC>;
    PSB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        F::SMEM_BYTES));
    kern<<<grid, dim3(C::THREADS, 1, 1), F::SMEM_BYTES, stream>>>(p, map);
    PSB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
}
#endif

}  // namespace psb
