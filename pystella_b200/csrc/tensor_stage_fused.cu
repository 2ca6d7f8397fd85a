// Fused stage of the tensor-perturbation sector in the warp-specialised TMA pipeline
// of fused_stage_ws.cuh: lap(hij_c) and the gradient of the source fields are
// evaluated in-kernel, so lap_hij / dfdx never touch HBM
// (TensorPerturbationSector.rhs_dict, pystella/sectors.py:187-204; scalar stress tensor
// with drop_trace=True, sectors.py:146-152; low-storage stage, step.py:473-491).
//
// One CTA column handles ONE packed component c (the component is the fastest grid
// dimension, so the six CTAs of a tile run together and share the source tiles in L2):
//   h ring  : R_f slots of the padded hij_c plane           (stencil: Laplacian)
//   s ring  : R_f slots of the padded planes of the NS source fields (gradient)
//   pw ring : R_p slots of {dhijdt_c, khij_c, kdhijdt_c} interior tiles
// x neighbours of h and of the sources live in per-thread register queues.
#include <math.h>
#include <string.h>

#include "fused_stage_impl.cuh"  // also brings in fused_stage_ws.cuh (the pipeline helpers)

namespace psb {

template <typename T> struct TensorFusedParams {
  const T* h_in;
  T* h_out;
  T* dh;
  T* kh;
  T* kdh;
  int64_t comp_stride;
  int Nx, Ny, Nz;
  int x_begin, x_end, xchunk;
  int x2_begin, x2_end, nchunks1;
  int nmain, split_len;
  int pf_dist;
  int wrap_x, wrap_y, wrap_z;
  T inv_dx2[3];
  T inv_dx[3];
  T dt, A, B, hubble, coef;
  const double* hubble_dev;
  T* peer_lo;   // peer-mapped hij_out of rank rx-1 / rx+1 (component 0) or null
  T* peer_hi;
  int peer_lo_Nx;
  int64_t peer_lo_stride, peer_hi_stride;
  TailParams tail;
};

// SMEMX: the x neighbours are read from the ring in shared memory (which then keeps the 2H
// planes behind the newest one: R_f = 2H+1+P) instead of living in per-thread register
// queues that are shifted once per plane -- ncu showed a third of the executed instructions
// of the register-queue formulation to be those moves (IMAD.MOV) and integer bookkeeping.
template <typename T, int H, int NS, int BY, int TZ, int P, int VEC_, bool SMEMX = false>
struct TwsCfg {
  using B = WsCfg<T, H, 1, BY, TZ, P, VEC_>;
  static constexpr int RF = SMEMX ? 2 * H + 1 + P : B::RF, RP = B::RP;
  static constexpr int FSLOT = (1 + NS) * B::FPLANE;   // h plane + NS source planes
  static constexpr int PWSLOT = 3 * B::PWTILE;
  static constexpr int NBAR = 2 * RF + 2 * RP;
  static constexpr int SMEM_BYTES = RF * FSLOT + RP * PWSLOT + NBAR * 8 + 128;
  static constexpr int THREADS = BY * TZ + 32;
  static constexpr int NCW = BY * TZ / 32;
};

// Body of one CTA for the packed component (CI, CJ), CI <= CJ, known at compile time: only
// the gradient axes the component needs are evaluated -- xx needs d_x only (no y/z neighbour
// loads of the source planes), yy / yz / zz need no x register queue of the sources at all --
// which removes a third of the shared-memory loads and of the arithmetic of the generic
// formulation and most of its selects.
template <typename T, int H, int NS, int BY, int TZ, int P, int VEC_, bool SMEMX, int CI,
          int CJ>
__device__ __forceinline__ void tensor_stage_body(const TensorFusedParams<T>& p,
                                                  const CUtensorMap& map_h,
                                                  const CUtensorMap& map_s,
                                                  const CUtensorMap& map_d,
                                                  const CUtensorMap& map_kh,
                                                  const CUtensorMap& map_kd) {
  constexpr bool NX = (CI == 0);                 // d_x of the sources needed
  constexpr bool NY = (CI == 1 || CJ == 1);      // d_y
  constexpr bool NZ = (CJ == 2);                 // d_z
  constexpr int comp = CI == 0 ? CJ : (CI == 1 ? 2 + CJ : 5);
  using C = TwsCfg<T, H, NS, BY, TZ, P, VEC_, SMEMX>;
  using B = typename C::B;
  constexpr int VEC = B::VEC, GV = B::GV, W = B::W, HL = B::HL, RF = C::RF, RP = C::RP;
  using V = Vec<T, VEC>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  unsigned char* fring = base;
  unsigned char* pring = base + RF * C::FSLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(pring + RP * C::PWSLOT);
  uint64_t* full_f = bars;
  uint64_t* empty_f = bars + RF;
  uint64_t* full_p = bars + 2 * RF;
  uint64_t* empty_p = bars + 2 * RF + RP;

  const int tid = threadIdx.x;
  const int z0 = (blockIdx.x / 6) * B::BZ, y0 = blockIdx.y * BY;
  const ChunkRange cr = chunk_range(blockIdx.z, p.x_begin, p.x_end, p.x2_begin, p.x2_end,
                                    p.xchunk, p.nchunks1, p.nmain, p.split_len,
                                    p.peer_hi != nullptr);
  const int xs = cr.xs, xe = cr.xe;
  const int nsteps = (xe - xs) + 2 * H;
  const int ncomp = xe - xs;
  const bool read_k = (p.A != (T)0);

  if (tid == 0) {
    for (int s = 0; s < RF; ++s) {
      mbar_init(&full_f[s], 1);
      mbar_init(&empty_f[s], C::NCW);
    }
    for (int s = 0; s < RP; ++s) {
      mbar_init(&full_p[s], 1);
      mbar_init(&empty_p[s], C::NCW);
    }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_launch_dependents();   // PSB_LAUNCH_PDL, see fused_stage_ws.cuh
  pdl_wait();

  if (tid >= BY * TZ) {
    if (tid == BY * TZ) {
      tma_prefetch_desc(&map_h);
      tma_prefetch_desc(&map_s);
      tma_prefetch_desc(&map_d);
      tma_prefetch_desc(&map_kh);
      tma_prefetch_desc(&map_kd);
      const uint32_t pw_bytes = (read_k ? 3u : 1u) * B::PWTILE_RAW;
      const int pf = p.pf_dist;
      auto prefetch_f = [&](int m) {
        tma_prefetch_4d(&map_h, z0 + H - HL, y0, xs + m, comp);
        // the source planes are shared by the six components of a tile: one of them asks
#pragma unroll
        for (int n = 0; n < NS; ++n)
          if (comp == n) tma_prefetch_4d(&map_s, z0 + H - HL, y0, xs + m, n);
      };
      auto prefetch_pw = [&](int c) {
        const int px = xs + c + H;
        tma_prefetch_4d(&map_d, z0 + H - B::PWL, y0 + H, px, comp);
        if (read_k) {
          tma_prefetch_4d(&map_kh, z0 + H - B::PWL, y0 + H, px, comp);
          tma_prefetch_4d(&map_kd, z0 + H - B::PWL, y0 + H, px, comp);
        }
      };
      for (int m = 0; m < nsteps; ++m) {
        if (pf > 0) {
          // pf planes ahead of the loads (the first RF / RP planes are loaded right away)
          if (m == 0)
            for (int q = RF; q < pf && q < nsteps; ++q) prefetch_f(q);
          if (m + pf >= RF && m + pf < nsteps) prefetch_f(m + pf);
          const int c0 = m - 2 * H;
          if (c0 == 0)
            for (int q = RP; q < pf && q < ncomp; ++q) prefetch_pw(q);
          if (c0 >= 0 && c0 + pf >= RP && c0 + pf < ncomp) prefetch_pw(c0 + pf);
        }
        {
          const int slot = m % RF;
          if (m >= RF) mbar_wait(&empty_f[slot], ((m / RF) - 1) & 1);
          mbar_expect_tx(&full_f[slot], (1 + NS) * B::FPLANE_RAW);
          unsigned char* dst = fring + slot * C::FSLOT;
          tma_load_4d(dst, &map_h, &full_f[slot], z0 + H - HL, y0, xs + m, comp);
#pragma unroll
          for (int n = 0; n < NS; ++n)
            tma_load_4d(dst + (1 + n) * B::FPLANE, &map_s, &full_f[slot], z0 + H - HL, y0,
                        xs + m, n);
        }
        const int c = m - 2 * H;
        if (c >= 0 && c < ncomp) {
          const int slot = c % RP;
          if (c >= RP) mbar_wait(&empty_p[slot], ((c / RP) - 1) & 1);
          mbar_expect_tx(&full_p[slot], pw_bytes);
          unsigned char* dst = pring + slot * C::PWSLOT;
          const int px = xs + c + H;
          tma_load_4d(dst, &map_d, &full_p[slot], z0 + H - B::PWL, y0 + H, px, comp);
          if (read_k) {
            tma_load_4d(dst + B::PWTILE, &map_kh, &full_p[slot], z0 + H - B::PWL, y0 + H, px,
                        comp);
            tma_load_4d(dst + 2 * B::PWTILE, &map_kd, &full_p[slot], z0 + H - B::PWL, y0 + H,
                        px, comp);
          }
        }
      }
    }
    return;
  }

  // split vectors (fp32, h = 2): pieces of GV sites, GS apart; see ld_pieces_s
  constexpr bool SPLIT = PSB_SPLIT_VECTORS && (GV < VEC);
  constexpr int NG = VEC / GV;
  constexpr int TS = SPLIT ? GV : VEC;        // z stride between threads (sites)
  constexpr int GS = SPLIT ? TZ * GV : GV;    // z stride between the pieces of a vector
  constexpr int ZG = SPLIT ? NG : 1;          // row segments a thread differentiates along z
  constexpr int ZW = SPLIT ? GV : VEC;        // sites per segment
  auto voff = [](int v) { return (v / GV) * GS + (v % GV); };   // element offset of site v
  const int tz = tid % TZ, ty = tid / TZ;
  const int lane = tid & 31;
  const int j = y0 + ty, k = z0 + tz * TS;
  bool pact[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) pact[g] = (j < p.Ny) && (k + g * GS < p.Nz);
  const bool active = pact[0];
  const int64_t Pz = p.Nz + 2 * H, Py = p.Ny + 2 * H;
  const int64_t plane_stride = Py * Pz;
  const int64_t site0 = ((int64_t)H * Py + (j + H)) * Pz + (k + H);

  T hubble = p.hubble;
  if (p.hubble_dev) hubble = (T)(*p.hubble_dev);
  const T two_h = (T)2 * hubble;

  const int own_off = (ty + H) * W + HL + tz * TS;
  const int pw_off = ty * B::PWW + B::PWL + tz * TS;

  V qh[2 * H + 1];
  V qs[NX ? NS : 1][2 * H + 1];   // x register queue of the sources: only for d_x
#pragma unroll
  for (int s = 0; s < 2 * H + 1; ++s) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) qh[s].v[v] = (T)0;
#pragma unroll
    for (int n = 0; n < (NX ? NS : 1); ++n)
#pragma unroll
      for (int v = 0; v < VEC; ++v) qs[n][s].v[v] = (T)0;
  }

  const int64_t cofs = (int64_t)comp * p.comp_stride;
  T* o_h = p.h_out + cofs + site0 + (int64_t)xs * plane_stride;
  T* o_d = p.dh + cofs + site0 + (int64_t)xs * plane_stride;
  T* o_kh = p.kh + cofs + site0 + (int64_t)xs * plane_stride;
  T* o_kd = p.kdh + cofs + site0 + (int64_t)xs * plane_stride;

  for (int n = 0; n < nsteps; ++n) {
    const int slot = n % RF;
    mbar_wait(&full_f[slot], (n / RF) & 1);
    if constexpr (!SMEMX) {
      const unsigned char* sl = fring + slot * C::FSLOT;
#pragma unroll
      for (int s = 0; s < 2 * H; ++s) qh[s] = qh[s + 1];
      qh[2 * H] = ld_pieces_s<T, VEC, GV, GS>(reinterpret_cast<const T*>(sl) + own_off);
      if constexpr (NX) {
#pragma unroll
        for (int m = 0; m < NS; ++m) {
#pragma unroll
          for (int s = 0; s < 2 * H; ++s) qs[m][s] = qs[m][s + 1];
          qs[m][2 * H] = ld_pieces_s<T, VEC, GV, GS>(
              reinterpret_cast<const T*>(sl + (1 + m) * B::FPLANE) + own_off);
        }
      }
    }

    if (n >= 2 * H) {
      const int c = n - 2 * H;
      const int i = xs + c;
      const int cslot = (n - H) % RF;
      const int pslot = c % RP;
      mbar_wait(&full_p[pslot], (c / RP) & 1);
      if (active) {
        const unsigned char* pw = pring + pslot * C::PWSLOT;
        const unsigned char* cen = fring + cslot * C::FSLOT;
        V dold = ld_pieces_s<T, VEC, GV, GS>(reinterpret_cast<const T*>(pw) + pw_off);
        V kho, kdo;
        if (read_k) {
          kho = ld_pieces_s<T, VEC, GV, GS>(reinterpret_cast<const T*>(pw + B::PWTILE) + pw_off);
          kdo = ld_pieces_s<T, VEC, GV, GS>(reinterpret_cast<const T*>(pw + 2 * B::PWTILE) + pw_off);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) kho.v[v] = kdo.v[v] = (T)0;
        }
        constexpr int NVZ = (ZW + 2 * HL) / GV;   // aligned pieces covering one row segment
        if constexpr (SMEMX) {
          // x neighbours of the centre plane n-H: planes n-2H .. n, all still in the ring
          int xslot = (n - 2 * H) % RF;
#pragma unroll
          for (int s = 0; s < 2 * H + 1; ++s) {
            const unsigned char* sl = fring + xslot * C::FSLOT;
            qh[s] = ld_pieces_s<T, VEC, GV, GS>(reinterpret_cast<const T*>(sl) + own_off);
            if constexpr (NX) {
              if (s != H) {
#pragma unroll
                for (int m = 0; m < NS; ++m)
                  qs[m][s] = ld_pieces_s<T, VEC, GV, GS>(
                      reinterpret_cast<const T*>(sl + (1 + m) * B::FPLANE) + own_off);
              }
            }
            xslot = (xslot + 1 == RF) ? 0 : xslot + 1;
          }
        }

        // ---- Laplacian of hij_c
        V lap;
        {
          const T* pl = reinterpret_cast<const T*>(cen);
          T zrow[ZG][ZW + 2 * HL];
#pragma unroll
          for (int zg = 0; zg < ZG; ++zg)
#pragma unroll
            for (int w = 0; w < NVZ; ++w) {
              Vec<T, GV> t = ld_vec<T, GV>(pl + own_off + zg * GS - HL + w * GV);
#pragma unroll
              for (int v = 0; v < GV; ++v) zrow[zg][w * GV + v] = t.v[v];
            }
          V yn[2 * H + 1];
#pragma unroll
          for (int s = 1; s <= H; ++s) {
            yn[H + s] = ld_pieces_s<T, VEC, GV, GS>(pl + own_off + s * W);
            yn[H - s] = ld_pieces_s<T, VEC, GV, GS>(pl + own_off - s * W);
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const T cc = qh[H].v[v];
            T sx = (T)lap_coef<H>(0) * cc, sy = sx, sz = sx;
#pragma unroll
            for (int s = 1; s <= H; ++s) {
              const T cs = (T)lap_coef<H>(s);
              sx = sx + cs * qh[H + s].v[v];
              sx = sx + cs * qh[H - s].v[v];
              sy = sy + cs * yn[H + s].v[v];
              sy = sy + cs * yn[H - s].v[v];
              sz = sz + cs * zrow[SPLIT ? v / GV : 0][HL + (SPLIT ? v % GV : v) + s];
              sz = sz + cs * zrow[SPLIT ? v / GV : 0][HL + (SPLIT ? v % GV : v) - s];
            }
            lap.v[v] = sx * p.inv_dx2[0] + sy * p.inv_dx2[1] + sz * p.inv_dx2[2];
          }
        }

        // ---- source term S_c = sum_n d_i f_n d_j f_n  (first-derivative stencils; only
        // the axes this component needs)
        V S;
#pragma unroll
        for (int v = 0; v < VEC; ++v) S.v[v] = (T)0;
#pragma unroll
        for (int m = 0; m < NS; ++m) {
          const T* pl = reinterpret_cast<const T*>(cen + (1 + m) * B::FPLANE);
          V gx, gy, gz;
          if constexpr (NX) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              T g = (T)0;
#pragma unroll
              for (int s = 1; s <= H; ++s) {
                const T cs = (T)grad_coef<H>(s);
                g = (s == 1) ? cs * qs[m][H + s].v[v] : g + cs * qs[m][H + s].v[v];
                g = g + (-cs) * qs[m][H - s].v[v];
              }
              gx.v[v] = g * p.inv_dx[0];
            }
          }
          if constexpr (NY) {
            V yn[2 * H + 1];
#pragma unroll
            for (int s = 1; s <= H; ++s) {
              yn[H + s] = ld_pieces_s<T, VEC, GV, GS>(pl + own_off + s * W);
              yn[H - s] = ld_pieces_s<T, VEC, GV, GS>(pl + own_off - s * W);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              T g = (T)0;
#pragma unroll
              for (int s = 1; s <= H; ++s) {
                const T cs = (T)grad_coef<H>(s);
                g = (s == 1) ? cs * yn[H + s].v[v] : g + cs * yn[H + s].v[v];
                g = g + (-cs) * yn[H - s].v[v];
              }
              gy.v[v] = g * p.inv_dx[1];
            }
          }
          if constexpr (NZ) {
            T zrow[ZG][ZW + 2 * HL];
#pragma unroll
            for (int zg = 0; zg < ZG; ++zg)
#pragma unroll
              for (int w = 0; w < NVZ; ++w) {
                Vec<T, GV> t = ld_vec<T, GV>(pl + own_off + zg * GS - HL + w * GV);
#pragma unroll
                for (int v = 0; v < GV; ++v) zrow[zg][w * GV + v] = t.v[v];
              }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              T g = (T)0;
#pragma unroll
              for (int s = 1; s <= H; ++s) {
                const T cs = (T)grad_coef<H>(s);
                const int zg = SPLIT ? v / GV : 0, ze = SPLIT ? v % GV : v;
                g = (s == 1) ? cs * zrow[zg][HL + ze + s] : g + cs * zrow[zg][HL + ze + s];
                g = g + (-cs) * zrow[zg][HL + ze - s];
              }
              gz.v[v] = g * p.inv_dx[2];
            }
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const T gi = CI == 0 ? gx.v[v] : (CI == 1 ? gy.v[v] : gz.v[v]);
            const T gj = CJ == 0 ? gx.v[v] : (CJ == 1 ? gy.v[v] : gz.v[v]);
            S.v[v] = (m == 0) ? gi * gj : S.v[v] + gi * gj;
          }
        }

        V ho, dn, khn, kdn;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const T d_old = dold.v[v];
          const T rhs_d = (lap.v[v] - two_h * d_old) + p.coef * S.v[v];
          T khv, kdv;
          if (read_k) {
            khv = p.A * kho.v[v] + p.dt * d_old;
            kdv = p.A * kdo.v[v] + p.dt * rhs_d;
          } else {
            khv = p.dt * d_old;
            kdv = p.dt * rhs_d;
          }
          khn.v[v] = khv;
          kdn.v[v] = kdv;
          ho.v[v] = qh[H].v[v] + p.B * khv;
          dn.v[v] = d_old + p.B * kdv;
        }
        st_pieces_s<T, VEC, GV, GS>(o_kh, khn, pact);
        st_pieces_s<T, VEC, GV, GS>(o_kd, kdn, pact);
        st_pieces_s<T, VEC, GV, GS>(o_d, dn, pact);
        st_pieces_s<T, VEC, GV, GS>(o_h, ho, pact);
        if (p.wrap_x) {
          if (i < H) st_pieces_s<T, VEC, GV, GS>(o_h + (int64_t)p.Nx * plane_stride, ho, pact);
          if (i >= p.Nx - H)
            st_pieces_s<T, VEC, GV, GS>(o_h - (int64_t)p.Nx * plane_stride, ho, pact);
        }
        // NVLink: my boundary planes -> the neighbours' x halos (peer memory)
        if (p.peer_lo != nullptr && i < H)
          st_pieces_s<T, VEC, GV, GS>(
              p.peer_lo + comp * p.peer_lo_stride +
                  (int64_t)(p.peer_lo_Nx + H + i) * plane_stride + (int64_t)(j + H) * Pz +
                  (k + H),
              ho, pact);
        if (p.peer_hi != nullptr && i >= p.Nx - H)
          st_pieces_s<T, VEC, GV, GS>(
              p.peer_hi + comp * p.peer_hi_stride + (int64_t)(i - p.Nx + H) * plane_stride +
                  (int64_t)(j + H) * Pz + (k + H),
              ho, pact);
        if (p.wrap_y) {
          if (j < H) st_pieces_s<T, VEC, GV, GS>(o_h + (int64_t)p.Ny * Pz, ho, pact);
          if (j >= p.Ny - H) st_pieces_s<T, VEC, GV, GS>(o_h - (int64_t)p.Ny * Pz, ho, pact);
        }
        if (p.wrap_z) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if (!pact[v / GV]) continue;
            if (k + voff(v) < H) o_h[voff(v) + p.Nz] = ho.v[v];
            if (k + voff(v) >= p.Nz - H) o_h[voff(v) - p.Nz] = ho.v[v];
          }
        }
      }
      o_h += plane_stride;
      o_d += plane_stride;
      o_kh += plane_stride;
      o_kd += plane_stride;
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_p[pslot]);
    }
    if constexpr (SMEMX) {
      // the oldest plane of the x stencil is no longer needed
      if (n >= 2 * H) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_f[(n - 2 * H) % RF]);
      }
    } else if (n >= H) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_f[(n - H) % RF]);
    }
  }
  if (p.tail.mode != 0) {
    const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const bool peer_store = (p.peer_lo != nullptr && xs < H) ||
                            (p.peer_hi != nullptr && xe > p.Nx - H);
    tail_publish_row<1, BY * TZ>(p.tail, blk, tid, 0, 0.0, peer_store);
    if (blk == 0)
      tail_watch<1, BY * TZ>(p.tail, (int64_t)gridDim.x * gridDim.y * gridDim.z, 0, tid);
  }
}

template <typename T, int H, int NS, int BY, int TZ, int P, int MINB, int VEC_, bool SMEMX>
__global__ void __launch_bounds__(BY* TZ + 32, MINB)
    tensor_stage_ws_kernel(const __grid_constant__ TensorFusedParams<T> p,
                           const __grid_constant__ CUtensorMap map_h,
                           const __grid_constant__ CUtensorMap map_s,
                           const __grid_constant__ CUtensorMap map_d,
                           const __grid_constant__ CUtensorMap map_kh,
                           const __grid_constant__ CUtensorMap map_kd) {
  // packed component (1,1),(1,2),(1,3),(2,2),(2,3),(3,3) = fastest grid dimension
#define PSB_TENSOR_BODY(CI_, CJ_)                                                        \
  tensor_stage_body<T, H, NS, BY, TZ, P, VEC_, SMEMX, CI_, CJ_>(p, map_h, map_s, map_d,    \
                                                                 map_kh, map_kd)
  switch (blockIdx.x % 6) {
    case 0: PSB_TENSOR_BODY(0, 0); break;
    case 1: PSB_TENSOR_BODY(0, 1); break;
    case 2: PSB_TENSOR_BODY(0, 2); break;
    case 3: PSB_TENSOR_BODY(1, 1); break;
    case 4: PSB_TENSOR_BODY(1, 2); break;
    default: PSB_TENSOR_BODY(2, 2); break;
  }
#undef PSB_TENSOR_BODY
}

constexpr int tws_smem_bytes(int es, int H, int NS, int BY, int TZ, int P, int VEC,
                             bool smemx) {
  const int BZ = TZ * VEC;
  const int GV = ws_gcd(VEC, H);
  const int HL = ((H + GV - 1) / GV) * GV;
  const int WQ = 16 / es;
  const int W = ((BZ + 2 * HL + WQ - 1) / WQ) * WQ;
  const int fplane = (((BY + 2 * H) * W * es + 127) / 128) * 128;
  const int pww = ((BZ + H % WQ + WQ - 1) / WQ) * WQ;
  const int pwtile = ((BY * pww * es + 127) / 128) * 128;
  const int rf = smemx ? 2 * H + 1 + P : H + 1 + P;
  return rf * (1 + NS) * fplane + (P + 1) * 3 * pwtile + (2 * rf + 2 * (P + 1)) * 8 + 128;
}
// two CTAs per SM where possible (113 KB each), else whatever fits
constexpr int tws_pick_tz(int es, int H, int NS, int BY, int P, int VEC, bool smemx) {
  for (int tz = 128 / VEC; tz >= 16; tz /= 2)
    if (tws_smem_bytes(es, H, NS, BY, tz, P, VEC, smemx) <= 113 * 1024) return tz;
  for (int tz = 128 / VEC; tz >= 16; tz /= 2)
    if (tws_smem_bytes(es, H, NS, BY, tz, P, VEC, smemx) <= SMEM_LIMIT) return tz;
  return 16;
}

// Default configuration: BY = 4 rows, pipeline depth P = 3, two CTAs per SM (measured, 512^3
// fp32: P = 2 0.650 of the roofline, P = 3 0.665, with 32-plane x-chunks 0.681; fp64 256^3:
// 0.710 -> 0.783).
template <typename T, int H, int NS, int VEC, int BY = 4, int P = 3, int MINB = 2,
          bool SMEMX = false>
static int launch_tensor_ws(const psb_tensor_fused_args* a, cudaStream_t stream,
                            int64_t* count) {
  constexpr int TZ = tws_pick_tz((int)sizeof(T), H, NS, BY, P, VEC, SMEMX);
  using C = TwsCfg<T, H, NS, BY, TZ, P, VEC, SMEMX>;
  using B = typename C::B;
  // the scalar stage's x-chunk schedule (split tail included)
  psb_scalar_stage_args xa;
  memset(&xa, 0, sizeof(xa));
  xa.x_begin = 0;
  xa.x_end = a->Nx;
  xa.tune_xchunk = a->tune_xchunk;
  const XLayout L = x_layout(&xa, a->Nx >= 64 ? 32 : WS_DEFAULT_XCHUNK,
                             6 * ((a->Nz + B::BZ - 1) / B::BZ) * ((a->Ny + BY - 1) / BY));
  if (count != nullptr) {
    *count = 6 * ((a->Nz + B::BZ - 1) / B::BZ) * ((a->Ny + BY - 1) / BY) * (int64_t)L.n1;
    return 0;
  }
  static_assert(C::SMEM_BYTES <= 227 * 1024, "tensor pipeline does not fit shared memory");
  const int64_t Px = a->Nx + 2 * H, Py = a->Ny + 2 * H, Pz = a->Nz + 2 * H;
  const int64_t vol = Px * Py * Pz;
  // argument checks first (they need no driver), then the TMA descriptors
  if (int rc = validate_tail(a->tail,
                             6 * ((a->Nz + B::BZ - 1) / B::BZ) * ((a->Ny + BY - 1) / BY) *
                                 (int64_t)L.n1,
                             false, "psb_tensor_stage_fused"))
    return rc;
  CUtensorMap mh, ms, md, mkh, mkd;
  if (encode_field_tensor_map(&mh, a->dtype, a->hij_in, 6, vol, Px, Py, Pz, 1, B::ROWS,
                              B::W) != 0 ||
      encode_field_tensor_map(&ms, a->dtype, a->src, NS, a->src_stride, Px, Py, Pz, 1,
                              B::ROWS, B::W) != 0 ||
      encode_field_tensor_map(&md, a->dtype, a->dhijdt, 6, vol, Px, Py, Pz, 1, BY,
                              B::PWW) != 0 ||
      encode_field_tensor_map(&mkh, a->dtype, a->khij, 6, vol, Px, Py, Pz, 1, BY, B::PWW) !=
          0 ||
      encode_field_tensor_map(&mkd, a->dtype, a->kdhijdt, 6, vol, Px, Py, Pz, 1, BY,
                              B::PWW) != 0) {
    set_error("psb_tensor_stage_fused: layout not describable by TMA (use the element-wise "
              "tensor stage)");
    return 3;
  }
  TensorFusedParams<T> p;
  p.h_in = static_cast<const T*>(a->hij_in);
  p.h_out = static_cast<T*>(a->hij_out);
  p.dh = static_cast<T*>(a->dhijdt);
  p.kh = static_cast<T*>(a->khij);
  p.kdh = static_cast<T*>(a->kdhijdt);
  p.comp_stride = vol;
  p.Nx = (int)a->Nx;
  p.Ny = (int)a->Ny;
  p.Nz = (int)a->Nz;
  p.x_begin = 0;
  p.x_end = (int)a->Nx;
  p.x2_begin = p.x2_end = 0;
  p.xchunk = L.xchunk;
  p.nchunks1 = L.n1;
  p.nmain = L.nmain;
  p.split_len = L.split_len;
  p.pf_dist = a->tune_prefetch == 0 ? PSB_DEFAULT_PREFETCH
                                    : (a->tune_prefetch < 0 ? 0 : a->tune_prefetch);
  p.wrap_x = a->wrap_x;
  p.wrap_y = a->wrap_y;
  p.wrap_z = a->wrap_z;
  for (int d = 0; d < 3; ++d) {
    p.inv_dx2[d] = (T)pow(a->dx[d], -2.0);
    p.inv_dx[d] = (T)(1.0 / a->dx[d]);
  }
  p.dt = (T)a->dt;
  p.A = (T)a->A;
  p.B = (T)a->B;
  p.hubble = (T)a->hubble;
  p.coef = (T)a->source_coef;
  p.hubble_dev = a->hubble_dev;
  p.peer_lo = static_cast<T*>(a->peer_lo_hij_out);
  p.peer_hi = static_cast<T*>(a->peer_hi_hij_out);
  p.peer_lo_Nx = (int)a->peer_lo_Nx;
  p.peer_lo_stride = a->peer_lo_stride;
  p.peer_hi_stride = a->peer_hi_stride;
  dim3 grid((unsigned)(6 * ((a->Nz + B::BZ - 1) / B::BZ)), (unsigned)((a->Ny + BY - 1) / BY),
            (unsigned)p.nchunks1);
  fill_tail(p.tail, a->tail, nullptr);
  dim3 block(C::THREADS, 1, 1);
  auto kern = tensor_stage_ws_kernel<T, H, NS, BY, TZ, P, MINB, VEC, SMEMX>;
  PSB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      C::SMEM_BYTES));
  PSB_CHECK_CUDA(launch_kernel(kern, grid, block, C::SMEM_BYTES, stream, a->launch_flags, p, mh,
                               ms, md, mkh, mkd));
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T, int H>
static int tensor_fused_ns(const psb_tensor_fused_args* a, cudaStream_t s, int64_t* count) {
  // 16-byte shared-memory vectors: 2 doubles, or 4 floats when the rows allow it
  if constexpr (H == 2) {
    // development variants of the headline radius (tile height, pipeline depth, CTAs/SM)
    constexpr int VV = sizeof(T) == 4 ? 4 : 2;
    if (a->nscalars == 2 && (sizeof(T) == 8 || a->Nz % 4 == 0)) {
      if (a->tune_variant == 2) return launch_tensor_ws<T, H, 2, VV, 4, 1, 3>(a, s, count);
      if (a->tune_variant == 4) return launch_tensor_ws<T, H, 2, VV, 8, 2, 1>(a, s, count);
      if (a->tune_variant == 10) return launch_tensor_ws<T, H, 2, VV, 4, 2, 1>(a, s, count);
      if (a->tune_variant == 6)
        return launch_tensor_ws<T, H, 2, VV, 4, 2, 2, true>(a, s, count);
      if (a->tune_variant == 7)
        return launch_tensor_ws<T, H, 2, VV, 4, 1, 2, true>(a, s, count);
      if (a->tune_variant == 8)
        return launch_tensor_ws<T, H, 2, VV, 4, 1, 3, true>(a, s, count);
    }
  }
  if constexpr (sizeof(T) == 4) {
    if (a->Nz % 4 == 0 && a->tune_variant != 1) {
      switch (a->nscalars) {
        case 1: return launch_tensor_ws<T, H, 1, 4>(a, s, count);
        case 2: return launch_tensor_ws<T, H, 2, 4>(a, s, count);
      }
    }
  }
  constexpr int VEC = 2;
  switch (a->nscalars) {
    case 1: return launch_tensor_ws<T, H, 1, VEC>(a, s, count);
    case 2: return launch_tensor_ws<T, H, 2, VEC>(a, s, count);
  }
  set_error("psb_tensor_stage_fused: nscalars=%d (1 or 2 source fields)", a->nscalars);
  return 1;
}

}  // namespace psb

static int tensor_fused_entry(const psb_tensor_fused_args* a, void* stream, int64_t* count) {
  using namespace psb;
  PSB_REQUIRE(a != nullptr, "psb_tensor_stage_fused: null args");
  PSB_REQUIRE(a->hij_in && a->hij_out && a->dhijdt && a->khij && a->kdhijdt && a->src,
              "psb_tensor_stage_fused: null array pointer");
  PSB_REQUIRE(a->hij_in != a->hij_out, "psb_tensor_stage_fused: hij_out must not alias hij_in");
  PSB_REQUIRE(a->Nx > 0 && a->Ny > 0 && a->Nz > 0, "psb_tensor_stage_fused: empty grid");
  PSB_REQUIRE(a->h == 2 || a->h == 4,
              "psb_tensor_stage_fused: h=%d (even halos only; use the element-wise stage)",
              a->h);
  PSB_REQUIRE(a->Nz % 2 == 0, "psb_tensor_stage_fused: odd Nz (use the element-wise stage)");
  PSB_REQUIRE(!(a->wrap_x && (a->peer_lo_hij_out || a->peer_hi_hij_out)),
              "psb_tensor_stage_fused: wrap_x and peer halos are mutually exclusive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dtype == PSB_F64)
    return a->h == 2 ? tensor_fused_ns<double, 2>(a, s, count)
                     : tensor_fused_ns<double, 4>(a, s, count);
  if (a->dtype == PSB_F32)
    return a->h == 2 ? tensor_fused_ns<float, 2>(a, s, count)
                     : tensor_fused_ns<float, 4>(a, s, count);
  set_error("psb_tensor_stage_fused: unknown dtype %d", a->dtype);
  return 1;
}

extern "C" int psb_tensor_stage_fused(const psb_tensor_fused_args* a, void* stream) {
  return tensor_fused_entry(a, stream, nullptr);
}

extern "C" int64_t psb_tensor_stage_fused_rows(const psb_tensor_fused_args* a) {
  int64_t count = 0;
  return tensor_fused_entry(a, nullptr, &count) == 0 ? count : -1;
}
