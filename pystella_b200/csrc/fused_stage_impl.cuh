// Fused finite-difference Laplacian + scalar-sector RHS + low-storage RK stage.
//
// One kernel per RK stage (replaces the reference's per-stage pair
// FiniteDifferencer.__call__ -> Stepper.__call__, pystella/derivs.py:339-429 and
// pystella/step.py:473-491, 524-528).  Each CTA owns a (BY x BZ) column of the
// rank-local grid and marches along x:
//
//   * the padded f planes (with y/z halos) stream through a ring of shared-memory
//     slots, filled by TMA (cp.async.bulk.tensor, one elected thread, mbarrier
//     completion) P planes ahead of the compute front;
//   * x-neighbours live in a per-thread register queue (2H+1 planes deep), y/z
//     neighbours are read from the centre plane in shared memory;
//   * dfdt / kf / kdfdt never need neighbours: they are read straight from HBM
//     into registers as 16-byte vectors one plane ahead, and written back in place;
//   * the new f goes to a second padded buffer (f_out) together with its periodic
//     images on the face halos ("write-through halos").
//
// HBM traffic per site and stage: read f, dfdt, kf, kdfdt; write f, dfdt, kf,
// kdfdt = 8*sizeof(T) (6*sizeof(T) when A == 0: the twins are not read).
#pragma once
#ifndef __CUDACC_RTC__
#include <math.h>
#include <string.h>

#include <type_traits>
#endif

#include "common.cuh"
#include "stage_tail.cuh"

namespace psb {

// measured (B200): prefetching planes into L2 ahead of the TMA loads makes every configuration
// slower (fp64 single field 0.946 -> 0.871 of the roofline, tensor stage fp32 0.650 -> 0.492): the
// pipelines are not bound by DRAM latency.  Kept as an opt-in tuning knob (tune_prefetch > 0).
#define PSB_DEFAULT_PREFETCH 0

template <typename T, int NF> struct FusedParams {
  const T* f_in;
  T* f_out;
  T* dfdt;
  T* kf;
  T* kdfdt;
  int64_t field_stride;
  int Nx, Ny, Nz;
  int x_begin, x_end, xchunk;
  int x2_begin, x2_end, nchunks1;
  int nmain, split_len;   // x-chunk schedule of range 1 (chunk_range())
  int pf_dist;            // L2 prefetch distance of the pipeline's producer (planes; 0 = off)
  int pws_rows;           // persistent kernel: chunks (= published rows) of the launch
  int quadratic;          // one field, V = c f^2 (dV/df = c' f): loop-free evaluation
  int wrap_x, wrap_y, wrap_z;
  T* peer_lo;   // peer-mapped f_out of rank rx-1 / rx+1 (or null)
  T* peer_hi;
  int peer_lo_Nx;
  int64_t peer_lo_stride, peer_hi_stride;
  T inv_dx2[3];
  T dt, A, B, a, hubble;
  const double* a_dev;
  const double* hubble_dev;
  int nterms[NF];
  T coef[NF][PSB_MAX_TERMS];
  int8_t power[NF][PSB_MAX_TERMS][NF];
  double* energy_partials;
  int64_t energy_slot_offset;
  int nvterms;
  T vcoef[PSB_MAX_TERMS];
  int8_t vpower[PSB_MAX_TERMS][NF];
  TailParams tail;     // in-kernel stage tail (stage_tail.cuh); mode 0 = none
};

// sum_t coef[t] * prod_m f_m^power[t][m] for the VEC sites of a thread at once: the loops
// over terms and powers (run-time, uniform) are outermost, the sites innermost, so their
// overhead is paid once per thread and not once per site.  Per site the association is
// term = ((coef * f) * f) ..., out = out + term in term order -- identical to the oracle's.
template <typename T, int NF, int VEC>
__device__ __forceinline__ Vec<T, VEC> poly_eval(int nterms, const T* coef,
                                                 const int8_t (*power)[NF],
                                                 const Vec<T, VEC>* fc) {
  Vec<T, VEC> out;
#ifdef PSB_POLY_PER_SITE
  // development switch: the round-1 formulation (all loops per site)
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    T acc = (T)0;
    for (int t = 0; t < nterms; ++t) {
      T term = coef[t];
#pragma unroll
      for (int m = 0; m < NF; ++m) {
        const T fm = fc[m].v[v];
        for (int e = 0; e < power[t][m]; ++e) term = term * fm;
      }
      acc = acc + term;
    }
    out.v[v] = acc;
  }
#else
#pragma unroll
  for (int v = 0; v < VEC; ++v) out.v[v] = (T)0;
  for (int t = 0; t < nterms; ++t) {
    Vec<T, VEC> term;
    const T c = coef[t];
#pragma unroll
    for (int v = 0; v < VEC; ++v) term.v[v] = c;
#pragma unroll
    for (int m = 0; m < NF; ++m) {
      const int pw = power[t][m];
      for (int e = 0; e < pw; ++e) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) term.v[v] = term.v[v] * fc[m].v[v];
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) out.v[v] = out.v[v] + term.v[v];
  }
#endif
  return out;
}

template <typename T, int VEC>
__device__ __forceinline__ Vec<T, VEC> ld_stream(const T* p) {
  Vec<T, VEC> r;
  if constexpr (sizeof(T) * VEC == 16) {
    float4 x = __ldcs(reinterpret_cast<const float4*>(p));
    r = *reinterpret_cast<Vec<T, VEC>*>(&x);
  } else if constexpr (sizeof(T) * VEC == 8) {
    float2 x = __ldcs(reinterpret_cast<const float2*>(p));
    r = *reinterpret_cast<Vec<T, VEC>*>(&x);
  } else {
    float x = __ldcs(reinterpret_cast<const float*>(p));
    r = *reinterpret_cast<Vec<T, VEC>*>(&x);
  }
  return r;
}

template <typename T, int VEC>
__device__ __forceinline__ void st_stream(T* p, const Vec<T, VEC>& v) {
  if constexpr (sizeof(T) * VEC == 16) {
    __stcs(reinterpret_cast<float4*>(p), *reinterpret_cast<const float4*>(&v));
  } else if constexpr (sizeof(T) * VEC == 8) {
    __stcs(reinterpret_cast<float2*>(p), *reinterpret_cast<const float2*>(&v));
  } else {
    __stcs(reinterpret_cast<float*>(p), *reinterpret_cast<const float*>(&v));
  }
}

// plain (L2-allocating) store: f_out is re-read by the next stage
template <typename T, int VEC>
__device__ __forceinline__ void st_plain(T* p, const Vec<T, VEC>& v) {
  *reinterpret_cast<Vec<T, VEC>*>(p) = v;
}

// x-chunk schedule of a launch (gridDim.z = chunk index, the slowest dispatch dimension):
// range 1 = [x_begin, x_end) is marched in chunks of `xchunk` planes, except that its LAST
// `split_len` planes are cut into split_len/2, /4, /4: the CTAs dispatched last are short, so
// the tail in which SMs run out of work is a quarter as long (measured: ~10 % of a 64-plane
// launch was idle SM time with uniform 16-plane chunks).  Range 2 (second boundary slab of
// the NCCL fallback) is uniform.  nmain = ordinary chunks of range 1 before the split.
struct ChunkRange {
  int xs, xe;
};
__device__ __forceinline__ ChunkRange chunk_range(int zc, int x_begin, int x_end, int x2_begin,
                                                  int x2_end, int xchunk, int nchunks1,
                                                  int nmain, int split_len,
                                                  bool top_first = false) {
  ChunkRange r;
  // top_first (launches that store halo planes into a neighbour's memory): the chunk holding
  // the HIGHEST planes is dispatched second instead of last, so that both boundary chunks'
  // NVLink stores and system-scope fences overlap with the bulk of the launch instead of
  // sitting on its critical path
  if (top_first && zc < nchunks1 && nchunks1 > 2)
    zc = zc == 0 ? 0 : (zc == 1 ? nchunks1 - 1 : zc - 1);
  if (zc >= nchunks1) {   // second x range of the launch
    r.xs = x2_begin + (zc - nchunks1) * xchunk;
    r.xe = min(r.xs + xchunk, x2_end);
  } else if (zc < nmain) {
    r.xs = x_begin + zc * xchunk;
    r.xe = min(r.xs + xchunk, x_end - split_len);
  } else {
    const int t = zc - nmain, base = x_end - split_len;
    const int h2 = split_len / 2, q = split_len / 4;
    r.xs = base + (t == 0 ? 0 : (t == 1 ? h2 : h2 + q));
    r.xe = base + (t == 0 ? h2 : (t == 1 ? h2 + q : split_len));
  }
  return r;
}

template <typename T, int H, int NF, int VEC, int BY, int TZ, int P>
struct FusedCfg {
  static constexpr int BZ = TZ * VEC;
  static constexpr int HL = ((H + VEC - 1) / VEC) * VEC;  // aligned left pad
  static constexpr int WQ = 16 / (int)sizeof(T);          // TMA rows: multiple of 16 B
  static constexpr int W = ((BZ + 2 * HL + WQ - 1) / WQ) * WQ;  // smem row length
  static constexpr int ROWS = BY + 2 * H;
  static constexpr int R = H + 1 + P;                     // ring slots
  static constexpr int PLANE_BYTES_RAW = ROWS * W * (int)sizeof(T);
  static constexpr int PLANE_BYTES = ((PLANE_BYTES_RAW + 127) / 128) * 128;
  static constexpr int PLANE = PLANE_BYTES / (int)sizeof(T);
  static constexpr int THREADS = BY * TZ;
  static constexpr int SMEM_BYTES = R * NF * PLANE_BYTES + R * 8 + 128;
  static_assert((W * sizeof(T)) % 16 == 0, "TMA box row must be a multiple of 16 B");
};

template <typename T, int H, int NF, int VEC, int BY, int TZ, int P, bool USE_TMA,
          bool ENERGY, int MINB>
__global__ void __launch_bounds__(BY* TZ, MINB)
    fused_stage_kernel(const __grid_constant__ FusedParams<T, NF> p,
                       const __grid_constant__ CUtensorMap tmap) {
  using C = FusedCfg<T, H, NF, VEC, BY, TZ, P>;
  using V = Vec<T, VEC>;
  constexpr int R = C::R, W = C::W, HL = C::HL;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  // align dynamic smem base to 128 B for TMA destinations
  unsigned char* base = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  T* ring = reinterpret_cast<T*>(base);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + R * NF * C::PLANE_BYTES);

  const int tz = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * TZ + tz;
  const int z0 = blockIdx.x * C::BZ, y0 = blockIdx.y * BY;
  const ChunkRange cr = chunk_range(blockIdx.z, p.x_begin, p.x_end, p.x2_begin, p.x2_end,
                                    p.xchunk, p.nchunks1, p.nmain, p.split_len,
                                    p.peer_hi != nullptr);
  const int xs = cr.xs, xe = cr.xe;
  const int nsteps = (xe - xs) + 2 * H;

  const int j = y0 + ty, k = z0 + tz * VEC;
  const bool active = (j < p.Ny) && (k < p.Nz);
  const int64_t Pz = p.Nz + 2 * H, Py = p.Ny + 2 * H;
  const int64_t plane_stride = Py * Pz;
  // offset of site (i=0, j, k) in a padded array
  const int64_t site0 = ((int64_t)H * Py + (j + H)) * Pz + (k + H);

  T a = p.a, hubble = p.hubble;
  if (p.a_dev) a = (T)(*p.a_dev);
  if (p.hubble_dev) hubble = (T)(*p.hubble_dev);
  const T two_h = (T)2 * hubble;
  const T a_sq = a * a;
  const bool read_k = (p.A != (T)0);

  auto issue_plane = [&](int n) {
    // plane arriving at step n: interior x index xs - H + n -> padded index xs + n
    const int slot = n % R;
    if constexpr (USE_TMA) {
      mbar_expect_tx(&bars[slot], NF * C::PLANE_BYTES_RAW);
#pragma unroll
      for (int f = 0; f < NF; ++f)
        tma_load_4d(ring + (slot * NF + f) * C::PLANE, &tmap, &bars[slot],
                    z0 + H - HL, y0, xs + n, f);
    }
  };

  auto coop_load_plane = [&](int n) {
    // generic loader (layouts TMA cannot describe): all threads, zero fill OOB
    const int slot = n % R;
    const int px = xs + n;
    for (int f = 0; f < NF; ++f) {
      T* dst = ring + (slot * NF + f) * C::PLANE;
      const T* src = p.f_in + f * p.field_stride + (int64_t)px * plane_stride;
      for (int e = tid; e < C::ROWS * W; e += C::THREADS) {
        const int r = e / W, c = e % W;
        const int gy = y0 + r, gz = z0 + H - HL + c;
        T val = (T)0;
        if (gy < Py && gz >= 0 && gz < Pz) val = src[(int64_t)gy * Pz + gz];
        dst[e] = val;
      }
    }
  };

  if constexpr (USE_TMA) {
    if (tid == 0) {
      tma_prefetch_desc(&tmap);
      for (int s = 0; s < R; ++s) mbar_init(&bars[s], 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      for (int n = 0; n <= P && n < nsteps; ++n) issue_plane(n);
    }
  } else {
    for (int n = 0; n <= P && n < nsteps; ++n) coop_load_plane(n);
    __syncthreads();
  }

  V q[NF][2 * H + 1];
  V cur_d[NF], cur_kf[NF], cur_kd[NF];
  V nxt_d[NF], nxt_kf[NF], nxt_kd[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) {
#pragma unroll
    for (int s = 0; s < 2 * H + 1; ++s)
#pragma unroll
      for (int v = 0; v < VEC; ++v) q[f][s].v[v] = (T)0;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      cur_d[f].v[v] = cur_kf[f].v[v] = cur_kd[f].v[v] = (T)0;
      nxt_d[f].v[v] = nxt_kf[f].v[v] = nxt_kd[f].v[v] = (T)0;
    }
  }

  const int own_off = (ty + H) * W + HL + tz * VEC;

  // fused energy sums of the input state: {dfdt_n^2}, V(f), {-f_n lap f_n}
  double e_kin[NF], e_grad[NF], e_pot = 0.0;
#pragma unroll
  for (int f = 0; f < NF; ++f) e_kin[f] = e_grad[f] = 0.0;

  for (int n = 0; n < nsteps; ++n) {
    const int i = xs - 2 * H + n;  // site plane computed at this step (if n >= 2H)

    // (1) prefetch next plane's pointwise arrays
    if (active && (n + 1 >= 2 * H) && (i + 1 < xe)) {
      const int64_t off = site0 + (int64_t)(i + 1) * plane_stride;
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const int64_t o = off + f * p.field_stride;
        nxt_d[f] = ld_stream<T, VEC>(p.dfdt + o);
        if (read_k) {
          nxt_kf[f] = ld_stream<T, VEC>(p.kf + o);
          nxt_kd[f] = ld_stream<T, VEC>(p.kdfdt + o);
        }
      }
    }

    // (2) wait for the f plane of this step, push own values into the x queue
    const int slot = n % R;
    if constexpr (USE_TMA) mbar_wait(&bars[slot], (n / R) & 1);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
#pragma unroll
      for (int s = 0; s < 2 * H; ++s) q[f][s] = q[f][s + 1];
      q[f][2 * H] = ld_vec<T, VEC>(ring + (slot * NF + f) * C::PLANE + own_off);
    }

    // (3) compute
    if (n >= 2 * H && active) {
      const int cslot = (n - H) % R;
      V lap[NF];
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const T* pl = ring + (cslot * NF + f) * C::PLANE;
        constexpr int NV = 1 + 2 * (HL / VEC);  // vectors covering the z neighbourhood
        T zrow[NV * VEC];
#pragma unroll
        for (int w = 0; w < NV; ++w) {
          V t = ld_vec<T, VEC>(pl + own_off - HL + w * VEC);
#pragma unroll
          for (int v = 0; v < VEC; ++v) zrow[w * VEC + v] = t.v[v];
        }
        V yn[2 * H + 1];
#pragma unroll
        for (int s = 1; s <= H; ++s) {
          yn[H + s] = ld_vec<T, VEC>(pl + own_off + s * W);
          yn[H - s] = ld_vec<T, VEC>(pl + own_off - s * W);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const T c = q[f][H].v[v];
          // term order of the reference: c0*f0, c1*f[+1], c1*f[-1], c2*f[+2], ...
          T sx = (T)lap_coef<H>(0) * c;
          T sy = (T)lap_coef<H>(0) * c;
          T sz = (T)lap_coef<H>(0) * c;
#pragma unroll
          for (int s = 1; s <= H; ++s) {
            const T cs = (T)lap_coef<H>(s);
            sx = sx + cs * q[f][H + s].v[v];
            sx = sx + cs * q[f][H - s].v[v];
            sy = sy + cs * yn[H + s].v[v];
            sy = sy + cs * yn[H - s].v[v];
            sz = sz + cs * zrow[HL + v + s];
            sz = sz + cs * zrow[HL + v - s];
          }
          lap[f].v[v] = sx * p.inv_dx2[0] + sy * p.inv_dx2[1] + sz * p.inv_dx2[2];
        }
      }

      V fc[NF];
#pragma unroll
      for (int f = 0; f < NF; ++f) fc[f] = q[f][H];

      if constexpr (ENERGY) {
        // per-plane partial sums in T over the thread's VEC sites, one double add each
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          T sk = (T)0, sg = (T)0;
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const T dv = cur_d[f].v[v];
            sk = sk + dv * dv;
            sg = sg + (-(fc[f].v[v] * lap[f].v[v]));
          }
          e_kin[f] += (double)sk;
          e_grad[f] += (double)sg;
        }
        const V Vv = poly_eval<T, NF, VEC>(p.nvterms, p.vcoef, p.vpower, fc);
        T sp = (T)0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) sp = sp + Vv.v[v];
        e_pot += (double)sp;
      }

      const int64_t off = site0 + (int64_t)i * plane_stride;
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        V fo, dn, kfn, kdn;
        // dV/df_f = sum_t coef * prod_m f_m^power
        const V dVv = poly_eval<T, NF, VEC>(p.nterms[f], p.coef[f], p.power[f], fc);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const T dV = dVv.v[v];
          const T d_old = cur_d[f].v[v];
          const T rhs_f = d_old;
          const T rhs_d = (lap[f].v[v] - two_h * d_old) - a_sq * dV;
          T kfv, kdv;
          if (read_k) {
            kfv = p.A * cur_kf[f].v[v] + p.dt * rhs_f;
            kdv = p.A * cur_kd[f].v[v] + p.dt * rhs_d;
          } else {
            kfv = p.dt * rhs_f;
            kdv = p.dt * rhs_d;
          }
          kfn.v[v] = kfv;
          kdn.v[v] = kdv;
          fo.v[v] = q[f][H].v[v] + p.B * kfv;
          dn.v[v] = d_old + p.B * kdv;
        }
        const int64_t o = off + f * p.field_stride;
        st_stream<T, VEC>(p.kf + o, kfn);
        st_stream<T, VEC>(p.kdfdt + o, kdn);
        st_stream<T, VEC>(p.dfdt + o, dn);
        T* fout = p.f_out + o;
        st_plain<T, VEC>(fout, fo);
        // periodic images on the face halos (faces only: star stencils)
        if (p.wrap_x) {
          if (i < H) st_plain<T, VEC>(fout + (int64_t)p.Nx * plane_stride, fo);
          if (i >= p.Nx - H) st_plain<T, VEC>(fout - (int64_t)p.Nx * plane_stride, fo);
        }
        // NVLink: my boundary planes -> the neighbours' x halos (peer memory)
        if (p.peer_lo != nullptr && i < H)
          st_plain<T, VEC>(p.peer_lo + f * p.peer_lo_stride +
                               (int64_t)(p.peer_lo_Nx + H + i) * plane_stride +
                               (int64_t)(j + H) * Pz + (k + H),
                           fo);
        if (p.peer_hi != nullptr && i >= p.Nx - H)
          st_plain<T, VEC>(p.peer_hi + f * p.peer_hi_stride +
                               (int64_t)(i - p.Nx + H) * plane_stride +
                               (int64_t)(j + H) * Pz + (k + H),
                           fo);
        if (p.wrap_y) {
          if (j < H) st_plain<T, VEC>(fout + (int64_t)p.Ny * Pz, fo);
          if (j >= p.Ny - H) st_plain<T, VEC>(fout - (int64_t)p.Ny * Pz, fo);
        }
        if (p.wrap_z) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if (k + v < H) fout[v + p.Nz] = fo.v[v];
            if (k + v >= p.Nz - H) fout[v - p.Nz] = fo.v[v];
          }
        }
      }
    }

    // (4) rotate prefetch registers
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      cur_d[f] = nxt_d[f];
      cur_kf[f] = nxt_kf[f];
      cur_kd[f] = nxt_kd[f];
    }

    // (5) everyone is done with the centre plane of this step: refill its slot
    __syncthreads();
    const int nn = n + P + 1;
    if (nn < nsteps) {
      if constexpr (USE_TMA) {
        if (tid == 0) issue_plane(nn);
      } else {
        coop_load_plane(nn);
        __syncthreads();
      }
    }
  }

  double xrow = 0.0;
  if constexpr (ENERGY) {
    // block reduction of the 2*NF+1 sums -> one partials row per CTA
    constexpr int NV = 2 * NF + 1;
    constexpr int NW = C::THREADS / 32;
    double vals[NV];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      vals[f] = e_kin[f];
      vals[NF + 1 + f] = e_grad[f];
    }
    vals[NF] = e_pot;
    double* red = reinterpret_cast<double*>(base);  // ring is dead: reuse it
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      double x = vals[r];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
      if ((tid & 31) == 0) red[(tid >> 5) * NV + r] = x;
    }
    __syncthreads();
    if (tid < NV) {
      double x = 0.0;
      for (int w = 0; w < NW; ++w) x += red[w * NV + tid];
      if (p.tail.mode == 0) {
        const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x +
                            blockIdx.x;
        p.energy_partials[(p.energy_slot_offset + blk) * PSB_ENERGY_STRIDE + tid] = x;
      }
      xrow = x;
    }
  }
  if (p.tail.mode != 0) {
    constexpr int NVROW = ENERGY ? 2 * NF + 1 : 0;
    const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const bool peer_store = (p.peer_lo != nullptr && xs < H) ||
                            (p.peer_hi != nullptr && xe > p.Nx - H);
    tail_publish_row<0, C::THREADS>(p.tail, blk, tid, NVROW, xrow, peer_store);
    if (blk == 0)
      tail_watch<0, C::THREADS>(p.tail, (int64_t)gridDim.x * gridDim.y * gridDim.z, NVROW,
                                tid);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------

// chunking of the (up to two) x ranges of one launch along gridDim.z
#ifndef __CUDACC_RTC__   // host side: launch geometry, parameter marshalling, dispatch
// set (per thread) by psb_scalar_stage_fused_plan: the dispatch then fills the plan of
// the instance it would launch instead of launching it
extern thread_local psb_fused_plan* t_fused_plan;

struct XLayout {
  int xchunk, n1, n2, nmain, split_len;
};
// tune_xchunk: 0 = default chunk with the split tail, > 0 = that chunk with the split tail,
// < 0 = uniform chunks of -tune_xchunk planes (no split)
static inline XLayout x_layout(const psb_scalar_stage_args* a, int default_chunk,
                               int64_t tiles = 0) {
  const int64_t nx1 = a->x_end - a->x_begin;
  const int64_t nx2 = a->x2_end > a->x2_begin ? a->x2_end - a->x2_begin : 0;
  const bool uniform = a->tune_xchunk < 0;
  int64_t xc = a->tune_xchunk > 0 ? a->tune_xchunk
                                  : (a->tune_xchunk < 0 ? -a->tune_xchunk : default_chunk);
  const int64_t nmax = nx1 > nx2 ? nx1 : nx2;
  if (xc > nmax) xc = nmax;
  if (xc < 1) xc = 1;
  XLayout L;
  L.xchunk = (int)xc;
  // the short chunks re-read their 2h prologue planes for little work: worth it only when a
  // chunk index holds few waves of CTAs (measured: 512^2 planes, 512 tiles: +2..3 %;
  // 1024^2 planes, 2048 tiles: -5 %)
  const bool few_tiles = tiles <= 0 || tiles <= 1024;
  L.split_len =
      (!uniform && few_tiles && xc >= 8 && xc % 4 == 0 && nx1 >= 2 * xc) ? (int)xc : 0;
  L.nmain = (int)((nx1 - L.split_len + xc - 1) / xc);
  L.n1 = L.nmain + (L.split_len ? 3 : 0);
  L.n2 = (int)((nx2 + xc - 1) / xc);
  return L;
}

template <typename T, int NF>
static void fill_params(FusedParams<T, NF>& p, const psb_scalar_stage_args* a) {
  p.f_in = static_cast<const T*>(a->f_in);
  p.f_out = static_cast<T*>(a->f_out);
  p.dfdt = static_cast<T*>(a->dfdt);
  p.kf = static_cast<T*>(a->kf);
  p.kdfdt = static_cast<T*>(a->kdfdt);
  p.field_stride = a->field_stride;
  p.Nx = (int)a->Nx;
  p.Ny = (int)a->Ny;
  p.Nz = (int)a->Nz;
  p.x_begin = (int)a->x_begin;
  p.x_end = (int)a->x_end;
  p.x2_begin = (int)a->x2_begin;
  p.x2_end = (int)a->x2_end;
  p.peer_lo = static_cast<T*>(a->peer_lo_f_out);
  p.peer_hi = static_cast<T*>(a->peer_hi_f_out);
  p.peer_lo_Nx = (int)a->peer_lo_Nx;
  p.peer_lo_stride = a->peer_lo_stride;
  p.peer_hi_stride = a->peer_hi_stride;
  p.wrap_x = a->wrap_x;
  p.wrap_y = a->wrap_y;
  p.wrap_z = a->wrap_z;
  for (int d = 0; d < 3; ++d) {
    // dx**-2 as the reference evaluates it in Python (derivs.py:295-296)
    p.inv_dx2[d] = (T)pow(a->dx[d], -2.0);
  }
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
  p.energy_partials = a->energy_partials;
  p.energy_slot_offset = a->energy_slot_offset;
  p.nvterms = a->nvterms;
  for (int t = 0; t < PSB_MAX_TERMS; ++t) {
    p.vcoef[t] = (T)a->vcoef[t];
    for (int m = 0; m < NF; ++m) p.vpower[t][m] = a->vpower[t][m];
  }
  // the most common potential (m^2 f^2 / 2, one field) is evaluated without the run-time
  // loops over terms and powers; same association as poly_eval, hence the same bits
  p.quadratic = (NF == 1 && a->nterms[0] == 1 && a->power[0][0][0] == 1 &&
                 (a->nvterms == 0 || (a->nvterms == 1 && a->vpower[0][0] == 2)))
                    ? 1
                    : 0;
  p.pf_dist = a->tune_prefetch == 0 ? PSB_DEFAULT_PREFETCH
                                    : (a->tune_prefetch < 0 ? 0 : a->tune_prefetch);
  fill_tail(p.tail, a->tail,
            a->energy_partials ? a->energy_partials + a->energy_slot_offset * PSB_ENERGY_STRIDE
                               : nullptr);
}

template <typename T, int H, int NF, int VEC, int BY, int TZ, int P, int MINB>
static int launch_fused(const psb_scalar_stage_args* a, cudaStream_t stream) {
  using C = FusedCfg<T, H, NF, VEC, BY, TZ, P>;
  FusedParams<T, NF> p;
  fill_params<T, NF>(p, a);
  const XLayout L = x_layout(a, 64, ((a->Nz + C::BZ - 1) / C::BZ) * ((a->Ny + BY - 1) / BY));
  p.xchunk = L.xchunk;
  p.nchunks1 = L.n1;
  p.nmain = L.nmain;
  p.split_len = L.split_len;
  dim3 grid((unsigned)((a->Nz + C::BZ - 1) / C::BZ), (unsigned)((a->Ny + BY - 1) / BY),
            (unsigned)(L.n1 + L.n2));
  dim3 block(TZ, BY, 1);

  const int64_t Px = a->Nx + 2 * H, Py = a->Ny + 2 * H, Pz = a->Nz + 2 * H;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  bool use_tma = (a->tune_variant != 2);
  if (use_tma) {
    int rc = encode_field_tensor_map(&tmap, a->dtype, a->f_in, NF, a->field_stride, Px,
                                     Py, Pz, 1, C::ROWS, C::W);
    if (rc != 0) use_tma = false;
  }

  const bool energy = a->energy_partials != nullptr;
#define PSB_LAUNCH_FUSED(TMA_, EN_)                                                 \
  do {                                                                              \
    auto kern = fused_stage_kernel<T, H, NF, VEC, BY, TZ, P, TMA_, EN_, MINB>;      \
    PSB_CHECK_CUDA(cudaFuncSetAttribute(                                            \
        kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));         \
    kern<<<grid, block, C::SMEM_BYTES, stream>>>(p, tmap);                          \
  } while (0)
  if (use_tma) {
    if (energy) PSB_LAUNCH_FUSED(true, true); else PSB_LAUNCH_FUSED(true, false);
  } else {
    if (energy) PSB_LAUNCH_FUSED(false, true); else PSB_LAUNCH_FUSED(false, false);
  }
#undef PSB_LAUNCH_FUSED
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// grid of one launch (also the number of energy-partial rows it writes)
template <typename T, int H, int NF, int VEC, int BY, int TZ, int P>
static int64_t fused_grid_blocks(const psb_scalar_stage_args* a) {
  using C = FusedCfg<T, H, NF, VEC, BY, TZ, P>;
  const XLayout L = x_layout(a, 64, ((a->Nz + C::BZ - 1) / C::BZ) * ((a->Ny + BY - 1) / BY));
  return ((a->Nz + C::BZ - 1) / C::BZ) * ((a->Ny + BY - 1) / BY) * (int64_t)(L.n1 + L.n2);
}

#endif  // !__CUDACC_RTC__
}  // namespace psb
#include "fused_stage_ws.cuh"
#include "fused_stage_pws.cuh"
#ifndef __CUDACC_RTC__
namespace psb {

constexpr int v1_smem_bytes(int es, int H, int NF, int VEC, int BY, int TZ, int P) {
  const int BZ = TZ * VEC;
  const int HL = ((H + VEC - 1) / VEC) * VEC;
  const int WQ = 16 / es;
  const int W = ((BZ + 2 * HL + WQ - 1) / WQ) * WQ;
  const int plane = (((BY + 2 * H) * W * es + 127) / 128) * 128;
  return (H + 1 + P) * NF * plane + (H + 1 + P) * 8 + 128;
}
constexpr int v1_pick_tz(int es, int H, int NF, int VEC, int BY, int P) {
  for (int tz = 64; tz >= 16; tz /= 2)
    if (v1_smem_bytes(es, H, NF, VEC, BY, tz, P) <= 100 * 1024) return tz;
  for (int tz = 64; tz >= 16; tz /= 2)
    if (v1_smem_bytes(es, H, NF, VEC, BY, tz, P) <= SMEM_LIMIT) return tz;
  return 16;
}

template <typename T, int H, int NF, int VEC, int BY, int TZ, int P, int MINB>
static int fused_op(const psb_scalar_stage_args* a, cudaStream_t stream, int64_t* count) {
  if (t_fused_plan) {
    set_error("psb_scalar_stage_fused_plan: this layout is served by the cooperative "
              "loader, not by the TMA pipeline (needs even Nz, 16-byte aligned rows and "
              "pointers)");
    return 3;
  }
  if (count) {
    *count = fused_grid_blocks<T, H, NF, VEC, BY, TZ, P>(a);
    return 0;
  }
  return launch_fused<T, H, NF, VEC, BY, TZ, P, MINB>(a, stream);
}

template <typename T, int H, int NF, int BY, int TZ, int P, int MINB, int VEC = 2,
          bool AL = false>
static int ws_op(const psb_scalar_stage_args* a, cudaStream_t stream, int64_t* count) {
  if (t_fused_plan) return plan_ws<T, H, NF, BY, TZ, P, MINB, VEC, AL>(a, t_fused_plan);
  if (count) {
    *count = ws_grid_blocks<T, H, NF, BY, TZ, P, VEC, AL>(a);
    return 0;
  }
  return launch_ws<T, H, NF, BY, TZ, P, MINB, VEC, AL>(a, stream);
}

// persistent variant (fused_stage_pws.cuh): MINB = resident CTAs per SM the grid is sized for
template <typename T, int H, int NF, int BY, int TZ, int P, int MINB, int VEC = 2>
static int pws_op(const psb_scalar_stage_args* a, cudaStream_t stream, int64_t* count) {
  if (t_fused_plan) {
    set_error("psb_scalar_stage_fused_plan: the persistent variant has no plan");
    return 3;
  }
  if (count) {
    *count = pws_rows<T, H, NF, BY, TZ, P, MINB, VEC>(a);
    return 0;
  }
  return launch_pws<T, H, NF, BY, TZ, P, MINB, VEC>(a, stream);
}

// can TMA describe all four arrays?  (mirrors encode_field_tensor_map's checks)
template <typename T>
static bool tma_layout_ok(const psb_scalar_stage_args* a, int H) {
  const int64_t es = sizeof(T);
  const int64_t Pz = a->Nz + 2 * H, Py = a->Ny + 2 * H;
  if ((Pz * es) % 16 || (Py * Pz * es) % 16 || (a->field_stride * es) % 16) return false;
  const void* ptrs[5] = {a->f_in, a->f_out, a->dfdt, a->kf, a->kdfdt};
  for (const void* q : ptrs)
    if (reinterpret_cast<uintptr_t>(q) % 16) return false;
  return true;
}

template <typename T, int H, int NF>
static int dispatch_vec(const psb_scalar_stage_args* a, cudaStream_t stream,
                        int64_t* count) {
  // 16-byte (fp64) / 8-byte (fp32) vector path needs even alignment everywhere
  const int64_t Pz = a->Nz + 2 * H;
  const int64_t Py = a->Ny + 2 * H;
  auto aligned = [&](const void* ptr) {
    return reinterpret_cast<uintptr_t>(ptr) % (2 * sizeof(T)) == 0;
  };
  const bool vec2 = (H % 2 == 0) && (Pz % 2 == 0) && (a->Nz % 2 == 0) &&
                    ((Py * Pz) % 2 == 0) && (a->field_stride % 2 == 0) &&
                    aligned(a->f_in) && aligned(a->f_out) && aligned(a->dfdt) &&
                    aligned(a->kf) && aligned(a->kdfdt) && a->tune_variant != 3;
  constexpr int ES = (int)sizeof(T);
  if (a->tune_variant == 16) {
    // EXPERIMENTAL aligned mapping: full 16-byte vectors for any radius (fp32: 4 sites per
    // thread, fp64: 2), masked row ends.  Not yet run on hardware.
    constexpr int VA = 16 / ES;
    const bool ok16 = (Pz % VA == 0) && ((Py * Pz) % VA == 0) && (a->field_stride % VA == 0) &&
                      tma_layout_ok<T>(a, H);
    if (!ok16) {
      set_error("psb_scalar_stage_fused: tune_variant 16 needs rows that are multiples of "
                "16 bytes and 16-byte aligned pointers");
      return 1;
    }
    if constexpr (WsCfg<T, H, NF, 4, 32, 2, VA, true>::SMEM_BYTES <= SMEM_LIMIT)
      return ws_op<T, H, NF, 4, 32, 2, 1, VA, true>(a, stream, count);
    else
      return ws_op<T, H, NF, 4, 16, 2, 1, VA, true>(a, stream, count);
  }
  if (vec2) {
    // warp-specialised TMA pipeline (default)
    const int var = a->tune_variant;
    if ((var == 0 || var >= 10) && tma_layout_ok<T>(a, H)) {
      if constexpr (H == 2 && NF <= 2) {
        // persistent CTAs with a continuous pipeline and guided self-scheduling
        // (fused_stage_pws.cuh); tune_variant 30 while it is being measured
        if (var == 30 && pws_eligible(a)) {
          if constexpr (std::is_same<T, double>::value && NF == 1) {
            return pws_op<T, H, NF, 4, 64, 3, 2>(a, stream, count);
          } else if constexpr (std::is_same<T, double>::value) {
            constexpr int TZp = ws_pick_tz(ES, H, NF, 4, 1, 2, SMEM_LIMIT_3);
            return pws_op<T, H, NF, 4, TZp, 1, 3>(a, stream, count);
          } else if constexpr (NF == 1) {
            if (a->Nz % 4 == 0) {
              constexpr int TZp = ws_pick_tz(ES, H, NF, 4, 2, 4);
              return pws_op<T, H, NF, 4, TZp, 2, 3, 4>(a, stream, count);
            }
          } else {
            if (a->Nz % 4 == 0) {
              constexpr int TZp = ws_pick_tz(ES, H, NF, 4, 1, 4, SMEM_LIMIT_3);
              return pws_op<T, H, NF, 4, TZp, 1, 3, 4>(a, stream, count);
            }
          }
        }
      }
      if constexpr (std::is_same<T, double>::value && H == 2 && NF == 1) {
        // tuned configurations of the headline case
        if (var == 11) return ws_op<T, H, NF, 8, 64, 3, 1>(a, stream, count);
        if (var == 12) return ws_op<T, H, NF, 8, 64, 2, 1>(a, stream, count);
        if (var == 13) return ws_op<T, H, NF, 4, 64, 2, 2>(a, stream, count);
        if (var == 14) return ws_op<T, H, NF, 2, 64, 3, 3>(a, stream, count);
        return ws_op<T, H, NF, 4, 64, 3, 2>(a, stream, count);
      } else {
        if constexpr (NF >= 2) {
          // several fields: the rings of the default depth leave room for 8 consumer warps
          // per SM only (latency-bound, ncu: 58 % of the cycles no warp is eligible); with a
          // one-plane-shallower pipeline three CTAs fit: measured 0.56 -> 0.75 of the HBM
          // roofline (2 fields fp32 512^3), 0.62 -> 0.79 (2 fields fp64), 0.40 -> 0.66
          // (3 fields fp64).  tune_variant 21 = the deeper two-CTA configuration.
          constexpr int VV = std::is_same<T, float>::value ? 4 : 2;
          constexpr int TZ3 = ws_pick_tz(ES, H, NF, 4, 1, VV, SMEM_LIMIT_3);
          constexpr int TZ2 = ws_pick_tz(ES, H, NF, 4, 3, VV, 113 * 1024);
          if constexpr (TZ2 > 0) {   // two CTAs with a deeper (P = 3) pipeline
            if (var == 22 && a->Nz % VV == 0)
              return ws_op<T, H, NF, 4, TZ2, 3, 2, VV>(a, stream, count);
          }
          if constexpr (TZ3 > 0) {
            if (var != 21 && var != 15 && a->Nz % VV == 0)
              return ws_op<T, H, NF, 4, TZ3, 1, 3, VV>(a, stream, count);
          }
        }
        if constexpr (std::is_same<T, float>::value) {
          // fp32: 4 sites per thread (16-byte shared-memory vectors) when rows allow
          constexpr int TZ4 = ws_pick_tz(ES, H, NF, 4, 2, 4);
          if constexpr (TZ4 > 0) {
            if (a->Nz % 4 == 0 && var != 15)
              return ws_op<T, H, NF, 4, TZ4, 2, 1, 4>(a, stream, count);
          }
        }
        constexpr int TZW = ws_pick_tz(ES, H, NF, 4, 2);
        if constexpr (TZW > 0) return ws_op<T, H, NF, 4, TZW, 2, 1>(a, stream, count);
      }
    }
    constexpr int TZ2 = v1_pick_tz(ES, H, NF, 2, (NF == 1 ? 8 : 4), 2);
    if constexpr (NF == 1)
      return fused_op<T, H, NF, 2, 8, TZ2, 2, (TZ2 == 64 ? 2 : 1)>(a, stream, count);
    else
      return fused_op<T, H, NF, 2, 4, TZ2, 2, 1>(a, stream, count);
  }
  constexpr int TZ1 = v1_pick_tz(ES, H, NF, 1, (NF == 1 ? 8 : 4), 2);
  if constexpr (NF == 1)
    return fused_op<T, H, NF, 1, 8, TZ1, 2, (TZ1 == 64 ? 2 : 1)>(a, stream, count);
  else
    return fused_op<T, H, NF, 1, 4, TZ1, 2, 1>(a, stream, count);
}

template <typename T, int H>
static int dispatch_nf(const psb_scalar_stage_args* a, cudaStream_t stream,
                       int64_t* count) {
  switch (a->nscalars) {
    case 1: return dispatch_vec<T, H, 1>(a, stream, count);
    case 2: return dispatch_vec<T, H, 2>(a, stream, count);
    case 3: return dispatch_vec<T, H, 3>(a, stream, count);
    case 4: return dispatch_vec<T, H, 4>(a, stream, count);
  }
  set_error("psb_scalar_stage_fused: nscalars=%d out of range", a->nscalars);
  return 1;
}

template <typename T>
static int dispatch_h(const psb_scalar_stage_args* a, cudaStream_t stream,
                      int64_t* count) {
  switch (a->h) {
    case 1: return dispatch_nf<T, 1>(a, stream, count);
    case 2: return dispatch_nf<T, 2>(a, stream, count);
    case 3: return dispatch_nf<T, 3>(a, stream, count);
    case 4: return dispatch_nf<T, 4>(a, stream, count);
  }
  set_error("psb_scalar_stage_fused: h=%d out of range 1..4", a->h);
  return 1;
}

}  // namespace psb
#endif  // !__CUDACC_RTC__
