// Persistent variant of the warp-specialised fused stage (fused_stage_ws.cuh): one CTA per
// SM slot for the whole launch, a CONTINUOUS TMA pipeline across tiles, and guided
// self-scheduling of the work.
//
// Why: ncu shows the SMs idle ~20 us at the end of every launch of the one-chunk-per-CTA
// kernel (the CTAs dispatched last finish spread over one CTA duration), and every CTA pays
// a start-up bubble (barrier init, descriptor fetch, first loads) plus 2h prologue planes per
// 16-plane chunk.  That fixed cost is 1.5 % of a 512^3 launch but 10 % of a 64-plane slab
// (config 3 per GPU at N = 8) and of 256^3 (config 2).
//
// Work decomposition: the lattice is cut into columns (BY x BZ tiles marched along x) and
// every column into chunks of 128 ... 4 site planes (long chunks for the bulk -- few
// prologues --, short ones at the end of the launch -- a short tail; see PwsSched).  CTAs
// claim chunks by number (one atomicAdd); the partition is static, only which CTA executes a
// chunk varies from run to run.  Each chunk
// publishes one row of energy partial sums (row index = chunk index): the watcher CTA's sum
// over the rows is deterministic, as in the one-chunk-per-CTA kernel.
//
// Pipeline: exactly the rings of fused_stage_ws.cuh; planes of consecutive chunks follow
// each other through the rings without draining them.  Every f slot carries a small header
// (column, padded x index, output range of the segment, row, flags) written by the producer
// before the TMA is issued; consumers learn from it where the plane belongs.
#pragma once
#include "fused_stage_ws.cuh"

namespace psb {

constexpr int PWS_MAX_CHUNK = 128;   // site planes of the longest chunks
constexpr int PWS_PHASES = 6;        // chunk lengths 128, 64, 32, 16, 8, 4 planes

struct PwsHeader {
  int col;        // column (tile) index
  int xpad;       // padded x index of this f plane
  int out_lo;     // segment: site planes [out_lo, out_hi) of this column are produced
  int out_hi;
  int row;        // energy / marker row of the chunk
  int flags;      // PWS_* below
  int pad0, pad1;
};
constexpr int PWS_TERMINATE = 1;   // no more work: consumers leave
constexpr int PWS_SEG_END = 2;     // last plane of a segment: release the trailing planes
constexpr int PWS_FLUSH = 4;       // last plane of a chunk: publish the chunk's row

// The schedule.  A chunk is an x range of ONE column; the x axis is cut into six phases with
// chunk lengths 128, 64, ... 4 planes (every phase leaves at least one chunk length for the
// shorter ones: few prologues for the bulk, short chunks at the end of the launch).  Chunks
// are numbered phase by phase, inside a phase x-range by x-range, inside an x-range column by
// column -- so the CTAs that run together work on NEIGHBOURING tiles of the SAME x range and
// share their y/z halo rows through L2, like the one-chunk-per-CTA kernel (a column-major
// numbering was measured first: 12 % slower at 512^3, the halo rows then come from DRAM).
// The numbering is a pure function of (Nx, columns): which CTA runs a chunk varies from run
// to run, the partition (and with it the order of the energy sums) does not.  Chunk i is
// claimed with ONE atomicAdd (a CAS loop on a packed word was measured first: 296 producers
// retrying against each other cost 0.5 ms per launch).
struct PwsSched {
  unsigned int* counter;      // chunks claimed beyond the first one of every CTA; zero
                              // between launches (re-armed by the last CTA)
  unsigned int* done;         // CTAs that have finished claiming
  unsigned nchunks;
  unsigned ncol;              // columns (tiles)
  unsigned phase_chunk0[PWS_PHASES + 1];   // first chunk index of each phase
  int phase_x0[PWS_PHASES + 1];            // first site plane of each phase
  int tiles_z;                // col = ty * tiles_z + tz
};

// chunk i -> (column, [out_lo, out_hi))
__host__ __device__ inline void pws_chunk(const PwsSched& s, unsigned i, int* col, int* out_lo,
                                          int* out_hi) {
  int len = PWS_MAX_CHUNK, ph = 0;
#pragma unroll
  for (int j = 1; j < PWS_PHASES; ++j) {
    if (i >= s.phase_chunk0[j]) {
      ph = j;
      len = PWS_MAX_CHUNK >> j;
    }
  }
  const unsigned local = i - s.phase_chunk0[ph];
  const int xb = (int)(local / s.ncol);
  *col = (int)(local % s.ncol);
  *out_lo = s.phase_x0[ph] + xb * len;
  const int end = s.phase_x0[ph + 1];
  *out_hi = *out_lo + len < end ? *out_lo + len : end;
}

template <typename T, int H, int NF, int BY, int TZ, int P, bool ENERGY, int MINB, int VEC_ = 2>
__global__ void __launch_bounds__(BY* TZ + 32, MINB)
    fused_stage_pws_kernel(const __grid_constant__ FusedParams<T, NF> p,
                           const __grid_constant__ PwsSched sched,
                           const __grid_constant__ CUtensorMap map_f,
                           const __grid_constant__ CUtensorMap map_d,
                           const __grid_constant__ CUtensorMap map_kf,
                           const __grid_constant__ CUtensorMap map_kd) {
  using C = WsCfg<T, H, NF, BY, TZ, P, VEC_, false>;
  constexpr int VEC = C::VEC, GV = C::GV, W = C::W, HL = C::HL, RF = C::RF, RP = C::RP;
  using V = Vec<T, VEC>;
  constexpr int NT = BY * TZ;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  unsigned char* fring = base;
  unsigned char* pring = base + RF * NF * C::FPLANE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(pring + RP * C::PWSLOT);
  uint64_t* full_f = bars;
  uint64_t* empty_f = bars + RF;
  uint64_t* full_p = bars + 2 * RF;
  uint64_t* empty_p = bars + 2 * RF + RP;
  PwsHeader* hdr = reinterpret_cast<PwsHeader*>(bars + 2 * RF + 2 * RP);
  double* red = reinterpret_cast<double*>(hdr + RF);   // [NCW][2 NF + 1]

  const int tid = threadIdx.x;
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

  if (tid >= NT) {
    // ======================= producer warp (one elected lane) ===================
    if (tid == NT) {
      tma_prefetch_desc(&map_f);
      tma_prefetch_desc(&map_d);
      tma_prefetch_desc(&map_kf);
      tma_prefetch_desc(&map_kd);
      const uint32_t pw_bytes = (read_k ? 3u : 1u) * NF * C::PWTILE_RAW;
      unsigned g = 0, gp = 0;   // f planes / pointwise planes issued so far
      auto f_slot = [&]() -> int {
        const int slot = (int)(g % RF);
        if (g >= (unsigned)RF) mbar_wait(&empty_f[slot], ((g / RF) - 1) & 1);
        return slot;
      };
      bool first = true;
      for (;;) {
        // ---- claim the next chunk: the CTA's own index first, then one atomicAdd each
        unsigned ci;
        if (first) {
          ci = blockIdx.x;
          first = false;
        } else {
          ci = gridDim.x + atomicAdd(sched.counter, 1u);
        }
        if (ci >= sched.nchunks) break;
        const unsigned row = ci;
        {
          int col, out_lo, out_hi;
          pws_chunk(sched, ci, &col, &out_lo, &out_hi);
          const int z0 = (col % sched.tiles_z) * C::BZ, y0 = (col / sched.tiles_z) * BY;
          const int nseg = (out_hi - out_lo) + 2 * H;
          for (int m = 0; m < nseg; ++m) {
            const int slot = f_slot();
            PwsHeader h;
            h.col = col;
            h.xpad = out_lo + m;
            h.out_lo = out_lo;
            h.out_hi = out_hi;
            h.row = (int)row;
            h.flags = (m == nseg - 1) ? (PWS_SEG_END | PWS_FLUSH) : 0;
            h.pad0 = h.pad1 = 0;
            hdr[slot] = h;
            mbar_expect_tx(&full_f[slot], NF * C::FPLANE_RAW);
#pragma unroll
            for (int f = 0; f < NF; ++f)
              tma_load_4d(fring + (slot * NF + f) * C::FPLANE, &map_f, &full_f[slot],
                          z0 + C::ZF, y0, out_lo + m, f);
            ++g;
            const int c = m - 2 * H;
            if (c >= 0) {   // pointwise planes of site plane out_lo + c
              const int pslot = (int)(gp % RP);
              if (gp >= (unsigned)RP) mbar_wait(&empty_p[pslot], ((gp / RP) - 1) & 1);
              mbar_expect_tx(&full_p[pslot], pw_bytes);
              unsigned char* dst = pring + pslot * C::PWSLOT;
              const int px = out_lo + c + H;
#pragma unroll
              for (int f = 0; f < NF; ++f) {
                tma_load_4d(dst + (0 * NF + f) * C::PWTILE, &map_d, &full_p[pslot],
                            z0 + C::ZP, y0 + H, px, f);
                if (read_k) {
                  tma_load_4d(dst + (1 * NF + f) * C::PWTILE, &map_kf, &full_p[pslot],
                              z0 + C::ZP, y0 + H, px, f);
                  tma_load_4d(dst + (2 * NF + f) * C::PWTILE, &map_kd, &full_p[pslot],
                              z0 + C::ZP, y0 + H, px, f);
                }
              }
              ++gp;
            }
          }
        }
      }
      // ---- no more work: tell the consumers, then re-arm the scheduler when every CTA is
      // through (only atomics touch these words: no fence needed)
      {
        const int slot = f_slot();
        PwsHeader h;
        h.col = h.xpad = h.out_lo = h.out_hi = h.row = h.pad0 = h.pad1 = 0;
        h.flags = PWS_TERMINATE;
        hdr[slot] = h;
        mbar_arrive(&full_f[slot]);
      }
      if (atomicAdd(sched.done, 1u) == gridDim.x - 1) {
        atomicExch(sched.counter, 0u);
        atomicExch(sched.done, 0u);
      }
    }
    return;
  }

  // ============================ consumer warps ==================================
  const int tz = tid % TZ, ty = tid / TZ;
  const int lane = tid & 31;
  const int64_t Pz = p.Nz + 2 * H, Py = p.Ny + 2 * H;
  const int64_t plane_stride = Py * Pz;
  using TagPlain = WsBool<false>;

  T a = p.a, hubble = p.hubble;
  if (p.a_dev) a = (T)(*p.a_dev);
  if (p.hubble_dev) hubble = (T)(*p.hubble_dev);
  const T two_h = (T)2 * hubble;
  const T a_sq = a * a;

  const int own_off = (ty + H) * W + HL + tz * VEC;   // elements, within an f plane
  const int pw_off = ty * C::PWW + C::PWL + tz * VEC;  // elements, within a pw tile

  V q[NF][2 * H + 1];
#pragma unroll
  for (int f = 0; f < NF; ++f)
#pragma unroll
    for (int s = 0; s < 2 * H + 1; ++s)
#pragma unroll
      for (int v = 0; v < VEC; ++v) q[f][s].v[v] = (T)0;

  double e_kin[NF], e_grad[NF], e_pot = 0.0;
#pragma unroll
  for (int f = 0; f < NF; ++f) e_kin[f] = e_grad[f] = 0.0;

  int cur_col = -1, j = 0, k = 0;
  bool active = false;
  int64_t site0 = 0;
  unsigned g = 0, gp = 0;
  int64_t rows_published = 0;

  for (;;) {
    const int slot = (int)(g % RF);
    mbar_wait(&full_f[slot], (g / RF) & 1);
    const PwsHeader hd = hdr[slot];
    if (hd.flags & PWS_TERMINATE) break;
    if (hd.col != cur_col) {
      cur_col = hd.col;
      const int z0 = (cur_col % sched.tiles_z) * C::BZ, y0 = (cur_col / sched.tiles_z) * BY;
      j = y0 + ty;
      k = z0 + tz * VEC;
      active = (j < p.Ny) && (k < p.Nz);
      site0 = ((int64_t)H * Py + (j + H)) * Pz + (k + H);
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
#pragma unroll
      for (int s = 0; s < 2 * H; ++s) q[f][s] = q[f][s + 1];
      q[f][2 * H] = ld_pieces<T, VEC, GV>(
          reinterpret_cast<const T*>(fring + (slot * NF + f) * C::FPLANE) + own_off);
    }

    const int m = hd.xpad - hd.out_lo;   // plane index within the segment
    const int i = hd.xpad - 2 * H;       // site plane whose x stencil is complete now
    if (m >= 2 * H) {
      const int cslot = (int)((g - H) % RF);
      const int pslot = (int)(gp % RP);
      mbar_wait(&full_p[pslot], (gp / RP) & 1);
      if (active) {
        const unsigned char* pw = pring + pslot * C::PWSLOT;
        V lap[NF], dold[NF], kfo[NF], kdo[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          dold[f] = ld_pieces<T, VEC, GV>(
              reinterpret_cast<const T*>(pw + (0 * NF + f) * C::PWTILE) + pw_off);
          if (read_k) {
            kfo[f] = ld_pieces<T, VEC, GV>(
                reinterpret_cast<const T*>(pw + (1 * NF + f) * C::PWTILE) + pw_off);
            kdo[f] = ld_pieces<T, VEC, GV>(
                reinterpret_cast<const T*>(pw + (2 * NF + f) * C::PWTILE) + pw_off);
          } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) kfo[f].v[v] = kdo[f].v[v] = (T)0;
          }
          const T* pl = reinterpret_cast<const T*>(fring + (cslot * NF + f) * C::FPLANE);
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
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const T cc = q[f][H].v[v];
            T sx = (T)lap_coef<H>(0) * cc;
            T sy = (T)lap_coef<H>(0) * cc;
            T sz = (T)lap_coef<H>(0) * cc;
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
#pragma unroll
          for (int f = 0; f < NF; ++f) {
            T sk = (T)0, sg = (T)0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              const T dv = dold[f].v[v];
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
          const V dVv = poly_eval<T, NF, VEC>(p.nterms[f], p.coef[f], p.power[f], fc);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const T dV = dVv.v[v];
            const T d_old = dold[f].v[v];
            const T rhs_d = (lap[f].v[v] - two_h * d_old) - a_sq * dV;
            T kfv, kdv;
            if (read_k) {
              kfv = p.A * kfo[f].v[v] + p.dt * d_old;
              kdv = p.A * kdo[f].v[v] + p.dt * rhs_d;
            } else {
              kfv = p.dt * d_old;
              kdv = p.dt * rhs_d;
            }
            kfn.v[v] = kfv;
            kdn.v[v] = kdv;
            fo.v[v] = q[f][H].v[v] + p.B * kfv;
            dn.v[v] = d_old + p.B * kdv;
          }
          const int64_t o = off + f * p.field_stride;
          st_pieces<T, VEC, GV, false>(p.kf + o, kfn);
          st_pieces<T, VEC, GV, false>(p.kdfdt + o, kdn);
          st_pieces<T, VEC, GV, false>(p.dfdt + o, dn);
          T* fout = p.f_out + o;
          st_pieces<T, VEC, GV, false>(fout, fo);
          if (p.wrap_x) {
            if (i < H) st_pieces<T, VEC, GV, false>(fout + (int64_t)p.Nx * plane_stride, fo);
            if (i >= p.Nx - H)
              st_pieces<T, VEC, GV, false>(fout - (int64_t)p.Nx * plane_stride, fo);
          }
          // NVLink: my boundary planes -> the neighbours' x halos (peer memory)
          if (p.peer_lo != nullptr && i < H)
            st_pieces<T, VEC, GV, false>(
                p.peer_lo + f * p.peer_lo_stride +
                    (int64_t)(p.peer_lo_Nx + H + i) * plane_stride + (int64_t)(j + H) * Pz +
                    (k + H),
                fo);
          if (p.peer_hi != nullptr && i >= p.Nx - H)
            st_pieces<T, VEC, GV, false>(
                p.peer_hi + f * p.peer_hi_stride + (int64_t)(i - p.Nx + H) * plane_stride +
                    (int64_t)(j + H) * Pz + (k + H),
                fo);
          if (p.wrap_y) {
            if (j < H) st_pieces<T, VEC, GV, false>(fout + (int64_t)p.Ny * Pz, fo);
            if (j >= p.Ny - H) st_pieces<T, VEC, GV, false>(fout - (int64_t)p.Ny * Pz, fo);
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
      // this warp is done with the pointwise slot
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_p[pslot]);
      ++gp;
    }
    // ... with the f plane that just served as centre plane, and at the end of a segment
    // with the H planes behind it (they only served as +x neighbours)
    __syncwarp();
    if (lane == 0) {
      if (m >= H) mbar_arrive(&empty_f[(g - H) % RF]);
      if (hd.flags & PWS_SEG_END) {
#pragma unroll
        for (int s = H - 1; s >= 0; --s) mbar_arrive(&empty_f[(g - s) % RF]);
      }
    }
    ++g;

    if (hd.flags & PWS_FLUSH) {
      // ---- chunk finished: publish its row (consumer warps only: named barrier 1)
      if (p.tail.mode != 0 || ENERGY) {
        constexpr int NV = 2 * NF + 1;
        double x = 0.0;
        if constexpr (ENERGY) {
          double vals[NV];
#pragma unroll
          for (int f = 0; f < NF; ++f) {
            vals[f] = e_kin[f];
            vals[NF + 1 + f] = e_grad[f];
            e_kin[f] = e_grad[f] = 0.0;
          }
          vals[NF] = e_pot;
          e_pot = 0.0;
          asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
#pragma unroll
          for (int r = 0; r < NV; ++r) {
            double y = vals[r];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) y += __shfl_xor_sync(0xffffffffu, y, d);
            if (lane == 0) red[(tid >> 5) * NV + r] = y;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
          if (tid < NV)
            for (int w = 0; w < C::NCW; ++w) x += red[w * NV + tid];
        }
        if (p.tail.mode != 0) {
          tail_publish_row<1, NT>(p.tail, hd.row, tid, ENERGY ? NV : 0, x, false);
        } else if (ENERGY && tid < NV) {
          p.energy_partials[(p.energy_slot_offset + hd.row) * PSB_ENERGY_STRIDE + tid] = x;
        }
        ++rows_published;
      }
    }
  }

  if (p.tail.mode != 0 && blockIdx.x == 0)
    tail_watch<1, NT>(p.tail, (int64_t)p.pws_rows, ENERGY ? 2 * NF + 1 : 0, tid);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
#ifndef __CUDACC_RTC__

// tabulate the schedule for columns of Nx site planes
static inline void pws_make_schedule(PwsSched& s, int Nx, unsigned ncol) {
  int x = 0;
  unsigned chunk = 0;
  s.ncol = ncol;
  for (int j = 0; j < PWS_PHASES; ++j) {
    const int len = PWS_MAX_CHUNK >> j;
    s.phase_chunk0[j] = chunk;
    s.phase_x0[j] = x;
    const int rem = Nx - x;
    int n;   // chunks of this length along x, per column
    if (j == PWS_PHASES - 1)
      n = (rem + len - 1) / len;                 // whatever is left, in the shortest chunks
    else
      n = rem > len ? (rem - len) / len : 0;     // leave at least `len` planes for them
    x += (j == PWS_PHASES - 1) ? rem : n * len;
    chunk += (unsigned)n * ncol;
  }
  s.phase_chunk0[PWS_PHASES] = chunk;
  s.phase_x0[PWS_PHASES] = Nx;
  s.nchunks = chunk;
}

template <typename T, int H, int NF, int BY, int TZ, int P, int VEC>
struct PwsGeom {
  using C = WsCfg<T, H, NF, BY, TZ, P, VEC, false>;
  static constexpr int SMEM_BYTES =
      C::SMEM_BYTES + C::RF * (int)sizeof(PwsHeader) + C::NCW * (2 * NF + 1) * 8 + 64;
};

// CTAs of the persistent launch: resident slots (MINB per SM), never more than chunks
template <typename T, int H, int NF, int BY, int TZ, int P, int MINB, int VEC>
static unsigned pws_grid(const psb_scalar_stage_args* a, PwsSched* s) {
  using C = WsCfg<T, H, NF, BY, TZ, P, VEC, false>;
  const int tz = (int)((a->Nz + C::BZ - 1) / C::BZ), ty = (int)((a->Ny + BY - 1) / BY);
  s->tiles_z = tz;
  pws_make_schedule(*s, (int)a->Nx, (unsigned)(tz * ty));
  unsigned grid = (unsigned)(num_sms() * MINB);
  if (grid > s->nchunks) grid = s->nchunks;
  return grid;
}

// the persistent kernel serves whole-slab launches only (single x range from 0 to Nx)
static inline bool pws_eligible(const psb_scalar_stage_args* a) {
  return a->x_begin == 0 && a->x_end == a->Nx && a->x2_end <= a->x2_begin &&
         a->pws_sched != nullptr &&
         (a->tail == nullptr || !(a->tail->mode & PSB_TAIL_PEER));
}

template <typename T, int H, int NF, int BY, int TZ, int P, int MINB, int VEC = 2>
static int64_t pws_rows(const psb_scalar_stage_args* a) {
  PwsSched s;
  pws_grid<T, H, NF, BY, TZ, P, MINB, VEC>(a, &s);
  return s.nchunks;
}

template <typename T, int H, int NF, int BY, int TZ, int P, int MINB, int VEC = 2>
static int launch_pws(const psb_scalar_stage_args* a, cudaStream_t stream) {
  using C = WsCfg<T, H, NF, BY, TZ, P, VEC, false>;
  using G = PwsGeom<T, H, NF, BY, TZ, P, VEC>;
  static_assert(G::SMEM_BYTES <= SMEM_LIMIT + 4096, "rings do not fit in shared memory");
  const int64_t Px = a->Nx + 2 * H, Py = a->Ny + 2 * H, Pz = a->Nz + 2 * H;
  CUtensorMap mf, md, mkf, mkd;
  if (encode_field_tensor_map(&mf, a->dtype, a->f_in, NF, a->field_stride, Px, Py, Pz, 1,
                              C::ROWS, C::W) != 0 ||
      encode_field_tensor_map(&md, a->dtype, a->dfdt, NF, a->field_stride, Px, Py, Pz, 1,
                              BY, C::PWW) != 0 ||
      encode_field_tensor_map(&mkf, a->dtype, a->kf, NF, a->field_stride, Px, Py, Pz, 1, BY,
                              C::PWW) != 0 ||
      encode_field_tensor_map(&mkd, a->dtype, a->kdfdt, NF, a->field_stride, Px, Py, Pz, 1,
                              BY, C::PWW) != 0)
    return -1;
  FusedParams<T, NF> p;
  fill_params<T, NF>(p, a);
  PwsSched s;
  const unsigned grid = pws_grid<T, H, NF, BY, TZ, P, MINB, VEC>(a, &s);
  s.counter = reinterpret_cast<unsigned int*>(a->pws_sched);
  s.done = reinterpret_cast<unsigned int*>(a->pws_sched) + 2;
  p.pws_rows = (int)s.nchunks;
  dim3 block(C::THREADS, 1, 1);
  if (a->energy_partials != nullptr) {
    auto kern = fused_stage_pws_kernel<T, H, NF, BY, TZ, P, true, MINB, VEC>;
    PSB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        G::SMEM_BYTES));
    kern<<<grid, block, G::SMEM_BYTES, stream>>>(p, s, mf, md, mkf, mkd);
  } else {
    auto kern = fused_stage_pws_kernel<T, H, NF, BY, TZ, P, false, MINB, VEC>;
    PSB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        G::SMEM_BYTES));
    kern<<<grid, block, G::SMEM_BYTES, stream>>>(p, s, mf, md, mkf, mkd);
  }
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#endif  // !__CUDACC_RTC__

}  // namespace psb
