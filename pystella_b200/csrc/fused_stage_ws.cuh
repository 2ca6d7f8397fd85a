// Warp-specialised fused stage (the default path when the layout is TMA-able).
//
// One producer warp streams *all four* input arrays of the stage through shared
// memory with TMA; the consumer warps never touch global memory for loads and
// never meet at a CTA-wide barrier:
//
//   f ring   : R_f = H+1+P slots, each a padded (BY+2H) x W plane per field
//              (own values -> x register queue; centre plane -> y/z neighbours)
//   pw ring  : R_p = P+1 slots, each {dfdt, kf, kdfdt} x (BY x BZ) per field
//   full_*   : mbarriers completed by TMA transaction bytes
//   empty_*  : mbarriers on which every consumer warp arrives when it is done
//              with a slot; the producer waits on them before refilling
//
// so P+1 planes (f + pointwise) are always in flight per CTA regardless of how
// many warps are resident: memory-level parallelism comes from the TMA queue, not
// from occupancy.  Results are stored straight from registers (16-byte vectors),
// f_out additionally onto the periodic face halos.
#pragma once
#include "fused_stage_impl.cuh"

namespace psb {

constexpr int ws_gcd(int a, int b) { return b == 0 ? a : ws_gcd(b, a % b); }

// VEC z-sites per thread (one 16-byte shared-memory vector: 2 doubles or 4 floats).
// GV = elements per global / pointwise access piece: the interior starts H elements
// into a row, so only gcd(VEC, H) consecutive elements are guaranteed aligned
// (fp32, h = 2: 16-byte LDS of the f tile, 8-byte pieces for everything else).
//
// AL ("aligned mapping", EXPERIMENTAL, tune_variant 16): the thread's VEC sites are shifted
// left by SH = H mod VEC so that they start on a VEC-element boundary of the padded row:
// every access is one full 16-byte vector for any radius; sites that fall outside
// 0 <= z < Nz at the two ends of a row are masked.
template <typename T, int H, int NF, int BY, int TZ, int P, int VEC_ = 2, bool AL = false>
struct WsCfg {
  static constexpr int VEC = VEC_;
  static constexpr int SH = AL ? H % VEC_ : 0;
  static constexpr int GV = AL ? VEC_ : ws_gcd(VEC_, H);
  static constexpr int BZ = TZ * VEC;
  static constexpr int GVO = VEC_;   // un-padded arrays: a thread's sites start VEC-aligned
  // left pad of the f tile: the TMA box must start 16-byte aligned in global memory,
  // i.e. at padded column z0 = z0 + H - HL with HL = H rounded up to GV elements
  static constexpr int HL = ((H + GV - 1) / GV) * GV;
  static constexpr int WQ = 16 / (int)sizeof(T);
  static constexpr int W = ((BZ + 2 * HL + WQ - 1) / WQ) * WQ;
  static constexpr int ROWS = BY + 2 * H;
  static constexpr int RF = H + 1 + P;
  static constexpr int RP = P + 1;
  static constexpr int FPLANE_RAW = ROWS * W * (int)sizeof(T);
  static constexpr int FPLANE = ((FPLANE_RAW + 127) / 128) * 128;   // bytes, per field
  // TMA needs a 16-byte aligned start along z: pointwise tiles begin PWL elements
  // left of the first interior site of the tile (fp32, h = 2: two floats)
  static constexpr int PWL = AL ? 0 : H % WQ;
  static constexpr int PWW = ((BZ + PWL + WQ - 1) / WQ) * WQ;
  // padded z coordinate (relative to z0) at which the f box / the pointwise boxes start
  static constexpr int ZF = H - HL - SH;
  static constexpr int ZP = H - PWL - SH;
  static constexpr int PWTILE_RAW = BY * PWW * (int)sizeof(T);
  static constexpr int PWTILE = ((PWTILE_RAW + 127) / 128) * 128;   // bytes, per array
  static constexpr int PWSLOT = 3 * NF * PWTILE;
  static constexpr int NCW = BY * TZ / 32;                           // consumer warps
  static constexpr int THREADS = BY * TZ + 32;                       // + producer warp
  static constexpr int NBAR = 2 * RF + 2 * RP;
  static constexpr int SMEM_BYTES = RF * NF * FPLANE + RP * PWSLOT + NBAR * 8 + 128;
};

constexpr int SMEM_LIMIT = 200 * 1024;
// x-planes marched per CTA: short chunks balance the 148 SMs better than they cost
// in re-read prologue planes (measured at 512^3: 16 > 32 > 64 > 128)
constexpr int WS_DEFAULT_XCHUNK = 16;

constexpr int ws_smem_bytes(int es, int H, int NF, int BY, int TZ, int P, int VEC = 2) {
  const int BZ = TZ * VEC;
  const int GV = ws_gcd(VEC, H);
  const int HL = ((H + GV - 1) / GV) * GV;
  const int WQ = 16 / es;
  const int W = ((BZ + 2 * HL + WQ - 1) / WQ) * WQ;
  const int fplane = (((BY + 2 * H) * W * es + 127) / 128) * 128;
  const int pww = ((BZ + H % WQ + WQ - 1) / WQ) * WQ;
  const int pwtile = ((BY * pww * es + 127) / 128) * 128;
  return (H + 1 + P) * NF * fplane + (P + 1) * 3 * NF * pwtile + (2 * (H + 1 + P) + 2 * (P + 1)) * 8 +
         128;
}
// widest z tile (64, 32 or 16 threads of 2 sites) whose rings fit; 0 = none
constexpr int ws_pick_tz(int es, int H, int NF, int BY, int P, int VEC = 2,
                         int limit = SMEM_LIMIT) {
  for (int tz = 128 / VEC; tz >= 16; tz /= 2)
    if (ws_smem_bytes(es, H, NF, BY, tz, P, VEC) <= limit) return tz;
  return 0;
}
// three CTAs per SM (227 KB of shared memory, 1 KB reserved per CTA)
constexpr int SMEM_LIMIT_3 = 74 * 1024;

// VEC elements moved as VEC/GV aligned pieces
template <typename T, int VEC, int GV>
__device__ __forceinline__ Vec<T, VEC> ld_pieces(const T* p) {
  Vec<T, VEC> r;
#pragma unroll
  for (int g = 0; g < VEC / GV; ++g) {
    Vec<T, GV> t = ld_vec<T, GV>(p + g * GV);
#pragma unroll
    for (int v = 0; v < GV; ++v) r.v[g * GV + v] = t.v[v];
  }
  return r;
}
template <typename T, int VEC, int GV, bool STREAM>
__device__ __forceinline__ void st_pieces(T* p, const Vec<T, VEC>& x) {
#pragma unroll
  for (int g = 0; g < VEC / GV; ++g) {
    Vec<T, GV> t;
#pragma unroll
    for (int v = 0; v < GV; ++v) t.v[v] = x.v[g * GV + v];
    if constexpr (STREAM)
      st_stream<T, GV>(p + g * GV, t);
    else
      st_plain<T, GV>(p + g * GV, t);
  }
}

// Measured (B200, 512^3 fp32): the split mapping halves the shared-memory wavefronts (ncu: Mem
// Busy 76 % -> 37 %) but the kernels are latency-bound, not LSU-bound, and its per-piece
// predicates cost more than it saves: 0.92 -> 0.70 of the roofline for one field.  Off.
#ifndef PSB_SPLIT_VECTORS
#define PSB_SPLIT_VECTORS 0
#endif

// "Split" vectors: when the interior starts only GV < VEC elements aligned (fp32, h = 2:
// 8 bytes), a thread's VEC sites are NOT consecutive but VEC/GV pieces of GV consecutive
// sites, GS = TZ*GV elements apart, so that the lanes of a warp touch CONSECUTIVE GV-element
// pieces: every shared-memory load and every global store of a piece is one contiguous
// 32*GV-element span (ncu, consecutive mapping: 8-byte pieces at a 16-byte lane stride took
// twice the ideal shared-memory wavefronts and L2 store sectors -- 43 % of all wavefronts).
template <typename T, int VEC, int GV, int GS>
__device__ __forceinline__ Vec<T, VEC> ld_pieces_s(const T* p) {
  Vec<T, VEC> r;
#pragma unroll
  for (int g = 0; g < VEC / GV; ++g) {
    Vec<T, GV> t = ld_vec<T, GV>(p + g * GS);
#pragma unroll
    for (int v = 0; v < GV; ++v) r.v[g * GV + v] = t.v[v];
  }
  return r;
}
// pact[g]: piece g lies inside the lattice
template <typename T, int VEC, int GV, int GS>
__device__ __forceinline__ void st_pieces_s(T* p, const Vec<T, VEC>& x, const bool* pact) {
#pragma unroll
  for (int g = 0; g < VEC / GV; ++g) {
    Vec<T, GV> t;
#pragma unroll
    for (int v = 0; v < GV; ++v) t.v[v] = x.v[g * GV + v];
    if (pact[g]) st_plain<T, GV>(p + g * GS, t);
  }
}

#ifdef PSB_GENERATED_RHS
// Generated-RHS instantiation (run-time compiled with NVRTC by pystella_b200/codegen.py):
// the translation unit defines, before including this header,
//   template <typename T> __device__ void psb_generated_rhs(int n, const T* f, const T* d,
//       const T* lap, T a, T hubble, const double* s, T& rhs_f, T& rhs_d);
//   template <typename T> __device__ T psb_generated_potential(const T* f, T a, T hubble,
//       const double* s);
// (values of all fields at one site; user scalars in s) and the kernel gains one
// parameter carrying those scalars.  Everything else -- rings, stencil, stores, halos,
// energy folding -- is the code below, unchanged.
struct PsbJitScalars {
  double v[32];
};
#endif

template <bool B> struct WsBool {
  static constexpr bool value = B;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <typename T, int H, int NF, int BY, int TZ, int P, bool ENERGY, int MINB,
          int VEC_ = 2, bool AL = false>
__global__ void __launch_bounds__(BY* TZ + 32, MINB)
    fused_stage_ws_kernel(const __grid_constant__ FusedParams<T, NF> p,
                          const __grid_constant__ CUtensorMap map_f,
                          const __grid_constant__ CUtensorMap map_d,
                          const __grid_constant__ CUtensorMap map_kf,
                          const __grid_constant__ CUtensorMap map_kd
#ifdef PSB_GENERATED_RHS
                          ,
                          const __grid_constant__ PsbJitScalars jit_scalars
#endif
) {
  using C = WsCfg<T, H, NF, BY, TZ, P, VEC_, AL>;
  constexpr int VEC = C::VEC, GV = C::GV, W = C::W, HL = C::HL, RF = C::RF, RP = C::RP;
  using V = Vec<T, VEC>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  unsigned char* fring = base;
  unsigned char* pring = base + RF * NF * C::FPLANE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(pring + RP * C::PWSLOT);
  uint64_t* full_f = bars;
  uint64_t* empty_f = bars + RF;
  uint64_t* full_p = bars + 2 * RF;
  uint64_t* empty_p = bars + 2 * RF + RP;

  const int tid = threadIdx.x;
  const int z0 = blockIdx.x * C::BZ, y0 = blockIdx.y * BY;
  const ChunkRange cr = chunk_range(blockIdx.z, p.x_begin, p.x_end, p.x2_begin, p.x2_end,
                                    p.xchunk, p.nchunks1, p.nmain, p.split_len,
                                    p.peer_hi != nullptr);
  const int xs = cr.xs, xe = cr.xe;
  const int nsteps = (xe - xs) + 2 * H;   // f planes streamed
  const int ncomp = xe - xs;              // site planes computed
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
  // PSB_LAUNCH_PDL: everything above overlaps the previous kernel's tail; nothing below may
  // run before that kernel has completed (inputs, ping-pong buffers, marker rows).
  pdl_launch_dependents();
  pdl_wait();

  if (tid >= BY * TZ) {
    // ======================= producer warp (one elected lane) ===================
    if (tid == BY * TZ) {
      tma_prefetch_desc(&map_f);
      tma_prefetch_desc(&map_d);
      tma_prefetch_desc(&map_kf);
      tma_prefetch_desc(&map_kd);
      const uint32_t pw_bytes = (read_k ? 3u : 1u) * NF * C::PWTILE_RAW;
      const int pf = p.pf_dist;
      auto prefetch_f = [&](int m) {
#pragma unroll
        for (int f = 0; f < NF; ++f) tma_prefetch_4d(&map_f, z0 + C::ZF, y0, xs + m, f);
      };
      auto prefetch_pw = [&](int c) {
        const int px = xs + c + H;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          tma_prefetch_4d(&map_d, z0 + C::ZP, y0 + H, px, f);
          if (read_k) {
            tma_prefetch_4d(&map_kf, z0 + C::ZP, y0 + H, px, f);
            tma_prefetch_4d(&map_kd, z0 + C::ZP, y0 + H, px, f);
          }
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
        {  // f plane m: interior x index xs - H + m -> padded index xs + m
          const int slot = m % RF;
          if (m >= RF) mbar_wait(&empty_f[slot], ((m / RF) - 1) & 1);
          mbar_expect_tx(&full_f[slot], NF * C::FPLANE_RAW);
#pragma unroll
          for (int f = 0; f < NF; ++f)
            tma_load_4d(fring + (slot * NF + f) * C::FPLANE, &map_f, &full_f[slot],
                        z0 + C::ZF, y0, xs + m, f);
        }
        // consumer step m needs f plane m and the pointwise planes of compute index
        // m - 2H: issue both in the same iteration so that both rings run P+1 ahead
        const int c = m - 2 * H;
        if (c >= 0 && c < ncomp) {
          const int slot = c % RP;
          if (c >= RP) mbar_wait(&empty_p[slot], ((c / RP) - 1) & 1);
          mbar_expect_tx(&full_p[slot], pw_bytes);
          unsigned char* dst = pring + slot * C::PWSLOT;
          const int px = xs + c + H;  // padded x index of site plane xs + c
#pragma unroll
          for (int f = 0; f < NF; ++f) {
            tma_load_4d(dst + (0 * NF + f) * C::PWTILE, &map_d, &full_p[slot],
                        z0 + C::ZP, y0 + H, px, f);
            if (read_k) {
              tma_load_4d(dst + (1 * NF + f) * C::PWTILE, &map_kf, &full_p[slot],
                          z0 + C::ZP, y0 + H, px, f);
              tma_load_4d(dst + (2 * NF + f) * C::PWTILE, &map_kd, &full_p[slot],
                          z0 + C::ZP, y0 + H, px, f);
            }
          }
        }
      }
    }
    return;
  }

  // ============================ consumer warps ==================================
  // split vectors (fp32, h = 2, consecutive mapping impossible at 16-byte alignment): the
  // thread's VEC sites are VEC/GV pieces of GV sites, GS apart; see ld_pieces_s
  constexpr bool SPLIT = PSB_SPLIT_VECTORS && !AL && (GV < VEC);
  constexpr int NG = VEC / GV;
  constexpr int TS = SPLIT ? GV : VEC;        // z stride between threads (sites)
  constexpr int GS = SPLIT ? TZ * GV : GV;    // z stride between the pieces of a vector
  constexpr int ZG = SPLIT ? NG : 1;          // row segments a thread differentiates along z
  constexpr int ZW = SPLIT ? GV : VEC;        // sites per segment
  auto voff = [](int v) { return (v / GV) * GS + (v % GV); };   // element offset of site v
  const int tz = tid % TZ, ty = tid / TZ;
  const int lane = tid & 31;
  const int j = y0 + ty, k = z0 + tz * TS - C::SH;
  // AL: sites k .. k+VEC-1 may stick out of the row at either end; SPLIT: whole pieces may
  // lie beyond the end of the row
  bool ok[VEC];
  bool all_ok = true, any_ok = false;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    ok[v] = (j < p.Ny) && (k + voff(v) >= 0) && (k + voff(v) < p.Nz);
    all_ok = all_ok && ok[v];
    any_ok = any_ok || ok[v];
  }
  bool pact[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) pact[g] = ok[g * GV];
  const bool active = (AL || SPLIT) ? any_ok : ((j < p.Ny) && (k < p.Nz));
  // store VEC results: one (or VEC/GV) vector stores, element-wise where masked
  auto put = [&](T* dst, const V& x, auto stream_tag) {
    constexpr bool STREAM = decltype(stream_tag)::value;
    if constexpr (SPLIT) {
      st_pieces_s<T, VEC, GV, GS>(dst, x, pact);
    } else if constexpr (!AL) {
      st_pieces<T, VEC, GV, STREAM>(dst, x);
    } else {
      if (all_ok) {
        st_pieces<T, VEC, GV, STREAM>(dst, x);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          if (ok[v]) dst[v] = x.v[v];
      }
    }
  };
  using TagPlain = WsBool<false>;
#ifdef PSB_STREAM_POINTWISE
  using TagPointwise = WsBool<true>;
#else
  using TagPointwise = WsBool<false>;
#endif
  const int64_t Pz = p.Nz + 2 * H, Py = p.Ny + 2 * H;
  const int64_t plane_stride = Py * Pz;
  const int64_t site0 = ((int64_t)H * Py + (j + H)) * Pz + (k + H);

  T a = p.a, hubble = p.hubble;
  if (p.a_dev) a = (T)(*p.a_dev);
  if (p.hubble_dev) hubble = (T)(*p.hubble_dev);
  const T two_h = (T)2 * hubble;
  const T a_sq = a * a;

  const int own_off = (ty + H) * W + HL + tz * TS;   // elements, within an f plane
  const int pw_off = ty * C::PWW + C::PWL + tz * TS;  // elements, within a pw tile

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

  // running output pointers of site plane i (advance by one plane per compute step)
  T* o_f = p.f_out + site0 + (int64_t)xs * plane_stride;
  T* o_d = p.dfdt + site0 + (int64_t)xs * plane_stride;
  T* o_kf = p.kf + site0 + (int64_t)xs * plane_stride;
  T* o_kd = p.kdfdt + site0 + (int64_t)xs * plane_stride;

  for (int n = 0; n < nsteps; ++n) {
    const int slot = n % RF;
    mbar_wait(&full_f[slot], (n / RF) & 1);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
#pragma unroll
      for (int s = 0; s < 2 * H; ++s) q[f][s] = q[f][s + 1];
      q[f][2 * H] = ld_pieces_s<T, VEC, GV, GS>(
          reinterpret_cast<const T*>(fring + (slot * NF + f) * C::FPLANE) + own_off);
    }

    if (n >= 2 * H) {
      const int c = n - 2 * H;       // compute index; site plane i = xs + c
      const int i = xs + c;
      const int cslot = (n - H) % RF;
      const int pslot = c % RP;
      mbar_wait(&full_p[pslot], (c / RP) & 1);
      if (active) {
        const unsigned char* pw = pring + pslot * C::PWSLOT;
        V lap[NF], dold[NF], kfo[NF], kdo[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          dold[f] = ld_pieces_s<T, VEC, GV, GS>(
              reinterpret_cast<const T*>(pw + (0 * NF + f) * C::PWTILE) + pw_off);
          if (read_k) {
            kfo[f] = ld_pieces_s<T, VEC, GV, GS>(
                reinterpret_cast<const T*>(pw + (1 * NF + f) * C::PWTILE) + pw_off);
            kdo[f] = ld_pieces_s<T, VEC, GV, GS>(
                reinterpret_cast<const T*>(pw + (2 * NF + f) * C::PWTILE) + pw_off);
          } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) kfo[f].v[v] = kdo[f].v[v] = (T)0;
          }
          const T* pl = reinterpret_cast<const T*>(fring + (cslot * NF + f) * C::FPLANE);
          constexpr int NVZ = (ZW + 2 * HL) / GV;  // aligned pieces covering a row segment
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
              sz = sz + cs * zrow[SPLIT ? v / GV : 0][HL + (SPLIT ? v % GV : v) + s];
              sz = sz + cs * zrow[SPLIT ? v / GV : 0][HL + (SPLIT ? v % GV : v) - s];
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
              if constexpr (AL || SPLIT) {
                if (!ok[v]) continue;
              }
              const T dv = dold[f].v[v];
              sk = sk + dv * dv;
              sg = sg + (-(fc[f].v[v] * lap[f].v[v]));
            }
            e_kin[f] += (double)sk;
            e_grad[f] += (double)sg;
          }
          T sp = (T)0;
#ifdef PSB_GENERATED_RHS
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if constexpr (AL || SPLIT) {
              if (!ok[v]) continue;
            }
            T fsite[NF];
#pragma unroll
            for (int m = 0; m < NF; ++m) fsite[m] = fc[m].v[v];
            sp = sp + psb_generated_potential<T>(fsite, a, hubble, jit_scalars.v);
          }
#else
          V Vv;
          if (NF == 1 && p.quadratic) {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
              Vv.v[v] = p.nvterms ? (T)0 + (p.vcoef[0] * fc[0].v[v]) * fc[0].v[v] : (T)0;
          } else {
            Vv = poly_eval<T, NF, VEC>(p.nvterms, p.vcoef, p.vpower, fc);
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            if constexpr (AL || SPLIT) {
              if (!ok[v]) continue;
            }
            sp = sp + Vv.v[v];
          }
#endif
          e_pot += (double)sp;
        }

#pragma unroll
        for (int f = 0; f < NF; ++f) {
          V fo, dn, kfn, kdn;
#ifndef PSB_GENERATED_RHS
          V dVv;
          if (NF == 1 && p.quadratic) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) dVv.v[v] = (T)0 + p.coef[0][0] * fc[0].v[v];
          } else {
            dVv = poly_eval<T, NF, VEC>(p.nterms[f], p.coef[f], p.power[f], fc);
          }
#endif
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
#ifdef PSB_GENERATED_RHS
            T fsite[NF], dsite[NF], lsite[NF];
#pragma unroll
            for (int m = 0; m < NF; ++m) {
              fsite[m] = q[m][H].v[v];
              dsite[m] = dold[m].v[v];
              lsite[m] = lap[m].v[v];
            }
            const T d_old = dold[f].v[v];
            T rhs_f, rhs_d;
            psb_generated_rhs<T>(f, fsite, dsite, lsite, a, hubble, jit_scalars.v, rhs_f,
                                 rhs_d);
            T kfv, kdv;
            if (read_k) {
              kfv = p.A * kfo[f].v[v] + p.dt * rhs_f;
              kdv = p.A * kdo[f].v[v] + p.dt * rhs_d;
            } else {
              kfv = p.dt * rhs_f;
              kdv = p.dt * rhs_d;
            }
#else
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
#endif
            kfn.v[v] = kfv;
            kdn.v[v] = kdv;
            fo.v[v] = q[f][H].v[v] + p.B * kfv;
            dn.v[v] = d_old + p.B * kdv;
          }
          const int64_t fo_off = f * p.field_stride;
          // plain stores for all four outputs: measured on B200 they beat evict-first
          // (st.cs) stores at both traffic mixes (stage 0: 2R:4W, stages 1-4: 4R:4W)
          put(o_kf + fo_off, kfn, TagPointwise{});
          put(o_kd + fo_off, kdn, TagPointwise{});
          put(o_d + fo_off, dn, TagPointwise{});
          T* fout = o_f + fo_off;
          put(fout, fo, TagPlain{});
          if (p.wrap_x) {
            if (i < H) put(fout + (int64_t)p.Nx * plane_stride, fo, TagPlain{});
            if (i >= p.Nx - H) put(fout - (int64_t)p.Nx * plane_stride, fo, TagPlain{});
          }
          // NVLink: my boundary planes -> the neighbours' x halos (peer memory)
          if (p.peer_lo != nullptr && i < H)
            put(p.peer_lo + f * p.peer_lo_stride +
                    (int64_t)(p.peer_lo_Nx + H + i) * plane_stride + (int64_t)(j + H) * Pz +
                    (k + H),
                fo, TagPlain{});
          if (p.peer_hi != nullptr && i >= p.Nx - H)
            put(p.peer_hi + f * p.peer_hi_stride + (int64_t)(i - p.Nx + H) * plane_stride +
                    (int64_t)(j + H) * Pz + (k + H),
                fo, TagPlain{});
          if (p.wrap_y) {
            if (j < H) put(fout + (int64_t)p.Ny * Pz, fo, TagPlain{});
            if (j >= p.Ny - H) put(fout - (int64_t)p.Ny * Pz, fo, TagPlain{});
          }
          if (p.wrap_z) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              if constexpr (AL || SPLIT) {
                if (!ok[v]) continue;
              }
              if (k + voff(v) < H) fout[voff(v) + p.Nz] = fo.v[v];
              if (k + voff(v) >= p.Nz - H) fout[voff(v) - p.Nz] = fo.v[v];
            }
          }
        }
      }
      o_f += plane_stride;
      o_d += plane_stride;
      o_kf += plane_stride;
      o_kd += plane_stride;
      // this warp is done with the pointwise slot
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_p[pslot]);
    }
    // ... and with the f plane that just served as centre plane
    if (n >= H) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_f[(n - H) % RF]);
    }
  }

  double xrow = 0.0;
  if constexpr (ENERGY) {
    // consumer-only reduction (the producer warp has exited): named barrier 1
    constexpr int NV = 2 * NF + 1;
    constexpr int NT = BY * TZ;
    double vals[NV];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      vals[f] = e_kin[f];
      vals[NF + 1 + f] = e_grad[f];
    }
    vals[NF] = e_pot;
    double* red = reinterpret_cast<double*>(pring);  // all pw slots are drained by now
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
#pragma unroll
    for (int r = 0; r < NV; ++r) {
      double x = vals[r];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
      if (lane == 0) red[(tid >> 5) * NV + r] = x;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
    if (tid < NV) {
      double x = 0.0;
      for (int w = 0; w < C::NCW; ++w) x += red[w * NV + tid];
      if (p.tail.mode == 0) {
        const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x +
                            blockIdx.x;
        p.energy_partials[(p.energy_slot_offset + blk) * PSB_ENERGY_STRIDE + tid] = x;
      }
      xrow = x;
    }
  }
  if (p.tail.mode != 0) {
    // consumer warps only (the producer warp has exited): named barrier 1
    constexpr int NVROW = ENERGY ? 2 * NF + 1 : 0;
    const int64_t blk = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const bool peer_store = (p.peer_lo != nullptr && xs < H) ||
                            (p.peer_hi != nullptr && xe > p.Nx - H);
    tail_publish_row<1, BY * TZ>(p.tail, blk, tid, NVROW, xrow, peer_store);
    if (blk == 0)
      tail_watch<1, BY * TZ>(p.tail, (int64_t)gridDim.x * gridDim.y * gridDim.z, NVROW, tid);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
#ifndef __CUDACC_RTC__

template <typename T, int H, int NF, int BY, int TZ, int P, int VEC = 2, bool AL = false>
static int64_t ws_grid_blocks(const psb_scalar_stage_args* a) {
  using C = WsCfg<T, H, NF, BY, TZ, P, VEC, AL>;
  const XLayout L = x_layout(a, WS_DEFAULT_XCHUNK, ((a->Nz + C::SH + C::BZ - 1) / C::BZ) *
                                                        ((a->Ny + BY - 1) / BY));
  return ((a->Nz + C::SH + C::BZ - 1) / C::BZ) * ((a->Ny + BY - 1) / BY) *
         (int64_t)(L.n1 + L.n2);
}

// parameters of the launch of this instance, for an externally compiled kernel
template <typename T, int H, int NF, int BY, int TZ, int P, int MINB, int VEC = 2,
          bool AL = false>
static int plan_ws(const psb_scalar_stage_args* a, psb_fused_plan* plan) {
  using C = WsCfg<T, H, NF, BY, TZ, P, VEC, AL>;
  if (AL) {
    set_error("psb_scalar_stage_fused_plan: the aligned-mapping variant has no plan");
    return 3;
  }
  static_assert(sizeof(FusedParams<T, NF>) <= PSB_FUSED_PLAN_PARAM_BYTES, "plan too small");
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  const int64_t Px = a->Nx + 2 * H, Py = a->Ny + 2 * H, Pz = a->Nz + 2 * H;
  const bool dry_run = plan->reserved != 0;   // geometry only (no driver needed)
  memset(plan, 0, sizeof(*plan));
  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  if (dry_run) {
    // descriptors stay zero
  } else if (encode_field_tensor_map(&maps[0], a->dtype, a->f_in, NF, a->field_stride, Px, Py, Pz,
                              1, C::ROWS, C::W) != 0 ||
      encode_field_tensor_map(&maps[1], a->dtype, a->dfdt, NF, a->field_stride, Px, Py, Pz,
                              1, BY, C::PWW) != 0 ||
      encode_field_tensor_map(&maps[2], a->dtype, a->kf, NF, a->field_stride, Px, Py, Pz, 1,
                              BY, C::PWW) != 0 ||
      encode_field_tensor_map(&maps[3], a->dtype, a->kdfdt, NF, a->field_stride, Px, Py, Pz,
                              1, BY, C::PWW) != 0) {
    set_error("psb_scalar_stage_fused_plan: TMA descriptors could not be encoded");
    return 3;
  }
  for (int m = 0; m < 4; ++m) memcpy(plan->tensor_maps[m], &maps[m], sizeof(CUtensorMap));
  FusedParams<T, NF> p;
  memset(&p, 0, sizeof(p));
  fill_params<T, NF>(p, a);
  const XLayout L = x_layout(a, WS_DEFAULT_XCHUNK, ((a->Nz + C::SH + C::BZ - 1) / C::BZ) *
                                                        ((a->Ny + BY - 1) / BY));
  p.xchunk = L.xchunk;
  p.nchunks1 = L.n1;
  p.nmain = L.nmain;
  p.split_len = L.split_len;
  memcpy(plan->params, &p, sizeof(p));
  plan->params_bytes = (uint32_t)sizeof(p);
  plan->dtype = a->dtype;
  plan->h = H;
  plan->nscalars = NF;
  plan->by = BY;
  plan->tz = TZ;
  plan->p = P;
  plan->minb = MINB;
  plan->vec = VEC;
  plan->energy = a->energy_partials != nullptr ? 1 : 0;
  plan->grid[0] = (uint32_t)((a->Nz + C::SH + C::BZ - 1) / C::BZ);
  plan->grid[1] = (uint32_t)((a->Ny + BY - 1) / BY);
  plan->grid[2] = (uint32_t)(L.n1 + L.n2);
  plan->threads = C::THREADS;
  plan->smem_bytes = C::SMEM_BYTES;
  return 0;
}

// returns 0 ok, >0 error, -1 = layout not TMA-able (caller falls back)
template <typename T, int H, int NF, int BY, int TZ, int P, int MINB, int VEC = 2,
          bool AL = false>
static int launch_ws(const psb_scalar_stage_args* a, cudaStream_t stream) {
  using C = WsCfg<T, H, NF, BY, TZ, P, VEC, AL>;
  static_assert(C::SMEM_BYTES <= SMEM_LIMIT, "rings do not fit in shared memory");
  const int64_t Px = a->Nx + 2 * H, Py = a->Ny + 2 * H, Pz = a->Nz + 2 * H;
  CUtensorMap mf, md, mkf, mkd;
  if (encode_field_tensor_map(&mf, a->dtype, a->f_in, NF, a->field_stride, Px, Py, Pz, 1,
                              C::ROWS, C::W) != 0)
    return -1;
  if (encode_field_tensor_map(&md, a->dtype, a->dfdt, NF, a->field_stride, Px, Py, Pz, 1,
                              BY, C::PWW) != 0 ||
      encode_field_tensor_map(&mkf, a->dtype, a->kf, NF, a->field_stride, Px, Py, Pz, 1, BY,
                              C::PWW) != 0 ||
      encode_field_tensor_map(&mkd, a->dtype, a->kdfdt, NF, a->field_stride, Px, Py, Pz, 1,
                              BY, C::PWW) != 0)
    return -1;

  FusedParams<T, NF> p;
  fill_params<T, NF>(p, a);
  const XLayout L = x_layout(a, WS_DEFAULT_XCHUNK, ((a->Nz + C::SH + C::BZ - 1) / C::BZ) *
                                                        ((a->Ny + BY - 1) / BY));
  p.xchunk = L.xchunk;
  p.nchunks1 = L.n1;
  p.nmain = L.nmain;
  p.split_len = L.split_len;
  dim3 grid((unsigned)((a->Nz + C::SH + C::BZ - 1) / C::BZ),
            (unsigned)((a->Ny + BY - 1) / BY), (unsigned)(L.n1 + L.n2));
  dim3 block(C::THREADS, 1, 1);
  if (a->energy_partials != nullptr) {
    auto kern = fused_stage_ws_kernel<T, H, NF, BY, TZ, P, true, MINB, VEC, AL>;
    PSB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C::SMEM_BYTES));
    PSB_CHECK_CUDA(launch_kernel(kern, grid, block, C::SMEM_BYTES, stream, a->launch_flags, p,
                                 mf, md, mkf, mkd));
  } else {
    auto kern = fused_stage_ws_kernel<T, H, NF, BY, TZ, P, false, MINB, VEC, AL>;
    PSB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C::SMEM_BYTES));
    PSB_CHECK_CUDA(launch_kernel(kern, grid, block, C::SMEM_BYTES, stream, a->launch_flags, p,
                                 mf, md, mkf, mkd));
  }
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#endif  // !__CUDACC_RTC__

}  // namespace psb
