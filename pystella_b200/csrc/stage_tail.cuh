// In-kernel tail of an RK stage: what used to be two or three tiny launches between the
// stage kernels (psb_energy_sums, psb_expansion_stage / psb_expansion_stage_peer) runs inside
// the stage kernel itself, in CTA 0 (the "watcher"), with NO fence, atomic or barrier added
// to the other CTAs:
//
//   * every CTA publishes one row of doubles when it is done -- its energy partial sums, or
//     a single marker for kernels without energy sums -- with plain 8-byte stores.  Rows are
//     pre-set to a sentinel bit pattern (PSB_TAIL_SENTINEL, a NaN payload no computation
//     produces), so a value is valid exactly when it differs from the sentinel: 8-byte
//     stores are single-copy atomic, no ordering between a row's values is needed (the
//     Lamport-style self-validating packets of the low-latency all-reduce protocols);
//   * CTA 0 first does its own tile, then consumes the rows in index order as they appear
//     (thread t adds rows t, t+NT, ... in increasing order, then a fixed tree: the sum is
//     deterministic, independent of timing), re-arms each row to the sentinel for the next
//     launch, and finally
//   FOLD    writes the sums to sums[0..2ns];
//   PEER    all-gathers them over peer memory and runs the flag barrier between the GPUs of
//           the node (CTAs that stored halo planes into a neighbour's memory publish their
//           row only after a system-scope fence, so "all rows seen" implies "all peer stores
//           visible"; this orders the halo stores before anybody's next stage and protects
//           the ping-pong buffers);
//   EXPAND  forms the energies with the current scale factor and advances the FLRW
//           background by one low-storage stage (pystella/expansion.py:119-157).
//
// Why not the classic last-CTA ticket (threadFenceReduction)?  Measured on B200 (512^3 fp64,
// 16384 CTAs of ~25 us): the __threadfence() + atomic at the end of EVERY CTA keeps its SM
// slot occupied until all of its stores have drained -- 7 % of the kernel time.  The watcher
// costs one of 296 CTA slots.
//
// Per stage this replaces Reduction.__call__ -> get_rho_and_p -> Expansion.step
// (pystella/reduction.py:237-255, sectors.py:211-229, expansion.py:140-157) and, multi-rank,
// the allreduce of decomp.py:470-491, with zero extra launches.
#pragma once
#ifndef __CUDACC_RTC__
#include <string.h>
#endif
#include "common.cuh"

namespace psb {

constexpr int TAIL_FOLD = 1, TAIL_EXPAND = 2, TAIL_PEER = 4;
constexpr unsigned long long TAIL_SENTINEL = PSB_TAIL_SENTINEL;

struct TailParams {
  int mode;             // 0 = no tail
  int ns;               // scalar fields in the energy sums (nvals = 2 ns + 1); 0 = none
  double* rows;         // one row of PSB_ENERGY_STRIDE doubles per CTA, sentinel-initialised
  double* sums;         // [16]: folded sums (written with FOLD, read otherwise)
  double grid_size, mpl, dt, A, B;
  double* state;        // {a, adot, hubble, ka, kadot}
  double* energies;     // {kinetic.., potential, gradient.., total, pressure}
  psb_peer_args peers;
};

#ifndef __CUDACC_RTC__
// rows_energy: row 0 of the launch's energy partial rows (kernels with energy sums publish
// through them), else the tail's own marker rows are used
inline void fill_tail(TailParams& t, const psb_stage_tail* a, double* rows_energy) {
  memset(&t, 0, sizeof(t));
  if (a == nullptr || a->mode == 0) return;
  t.mode = a->mode;
  t.ns = a->nscalars;
  t.rows = rows_energy != nullptr ? rows_energy : a->rows;
  t.sums = a->sums;
  t.grid_size = a->grid_size;
  t.mpl = a->mpl;
  t.dt = a->dt;
  t.A = a->A;
  t.B = a->B;
  t.state = a->state;
  t.energies = a->energies;
  t.peers = a->peers;
}
int validate_tail(const psb_stage_tail* a, int64_t nblocks, bool has_partials, const char* who);
#endif

template <int BAR, int NT> __device__ __forceinline__ void tail_bar() {
  asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory");
}

__device__ __forceinline__ void tail_publish(double* p, double x) {
  asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(x) : "memory");
}
__device__ __forceinline__ unsigned long long tail_peek(const double* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

struct TailEnergies {
  double total, pressure;
};

__device__ inline TailEnergies tail_form_energies(const double* sums, int ns, double grid_size,
                                                  double a, double* energies) {
  // kinetic_n = <dfdt_n^2> / 2 / a^2 ; potential = <V> ; gradient_n = <-f lap f> / 2 / a^2
  const double a2 = a * a;
  double kin = 0.0, grad = 0.0;
  for (int n = 0; n < ns; ++n) {
    const double k = sums[n] / 2.0 / a2 / grid_size;
    const double g = sums[ns + 1 + n] / 2.0 / a2 / grid_size;
    energies[n] = k;
    energies[ns + 1 + n] = g;
    kin += k;
    grad += g;
  }
  const double pot = sums[ns] / grid_size;
  energies[ns] = pot;
  TailEnergies e;
  e.total = kin + pot + grad;
  e.pressure = kin - grad / 3.0 - pot;
  energies[2 * ns + 1] = e.total;
  energies[2 * ns + 2] = e.pressure;
  return e;
}

__device__ inline void tail_expansion_update(const double* sums, int ns, double grid_size,
                                             double mpl, double dt, double A, double B,
                                             double* state, double* energies) {
  double a = state[0], adot = state[1], ka = state[3], kadot = state[4];
  const TailEnergies e = tail_form_energies(sums, ns, grid_size, a, energies);
  const double pi = 3.141592653589793;
  // a' = adot ; adot' = 4 pi a^2 / 3 / mpl^2 (rho - 3 P) a   (expansion.py:119-138)
  const double rhs_a = adot;
  const double rhs_adot =
      4.0 * pi * (a * a) / 3.0 / (mpl * mpl) * (e.total - 3.0 * e.pressure) * a;
  ka = A * ka + dt * rhs_a;
  a = a + B * ka;
  kadot = A * kadot + dt * rhs_adot;
  adot = adot + B * kadot;
  state[0] = a;
  state[1] = adot;
  state[2] = adot / a;
  state[3] = ka;
  state[4] = kadot;
}

// All-gather of `sums` over peer memory + inter-GPU flag barrier; on return total[0..nvals)
// holds the rank-ordered sum (bit-identical on every rank).  Called by one CTA, NT threads.
template <int BAR, int NT>
__device__ inline void tail_peer_exchange(const psb_peer_args& peers, const double* sums,
                                          int nvals, double* total, int tid) {
  const int P = peers.nranks, me = peers.rank;
  const unsigned long long epoch = *peers.epoch + 1ull;
  const int par = (int)(epoch & 1ull);
  // 1. my sums -> slot [par][me] of every rank (incl. myself)
  for (int e = tid; e < P * nvals; e += NT) {
    const int q = e / nvals, r = e % nvals;
    double* dst = static_cast<double*>(peers.peer_slots[q]) +
                  ((int64_t)par * P + me) * PSB_ENERGY_STRIDE;
    dst[r] = sums[r];
  }
  __threadfence_system();
  tail_bar<BAR, NT>();
  // 2. raise my flag on every rank: everything this rank stored to its peers before (the
  //    halo planes of this kernel: their CTAs fenced at system scope before publishing the
  //    rows this CTA has consumed) is ordered before it
  for (int q = tid; q < P; q += NT) {
    unsigned long long* fl =
        static_cast<unsigned long long*>(peers.peer_flags[q]) + (int64_t)par * P + me;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fl), "l"(epoch) : "memory");
  }
  // 3. wait for every rank's flag
  for (int q = tid; q < P; q += NT) {
    const unsigned long long* mine =
        static_cast<unsigned long long*>(peers.peer_flags[me]) + (int64_t)par * P + q;
    unsigned long long seen;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
    } while (seen < epoch);
  }
  __threadfence_system();
  tail_bar<BAR, NT>();
  if (tid == 0) *peers.epoch = epoch;
  // 4. add the slots in rank order
  if (tid < nvals) {
    const volatile double* sl =
        static_cast<double*>(peers.peer_slots[me]) + (int64_t)par * P * PSB_ENERGY_STRIDE;
    double x = 0.0;
    for (int q = 0; q < P; ++q) x += sl[q * PSB_ENERGY_STRIDE + tid];
    total[tid] = x;
  }
  tail_bar<BAR, NT>();
}

// Every CTA: publish this CTA's row.  vals[0..nv) are the values of thread `tid` == r (the
// caller passes x for tid < nv); kernels without energy sums pass nv = 0 and a marker is
// written.  CTAs that stored into a neighbour's memory fence at system scope first.
template <int BAR, int NT>
__device__ __forceinline__ void tail_publish_row(const TailParams& t, int64_t blk, int tid,
                                                 int nv, double x, bool did_peer_store) {
  if ((t.mode & TAIL_PEER) && did_peer_store) {
    __threadfence_system();
    tail_bar<BAR, NT>();
  }
  double* row = t.rows + blk * PSB_ENERGY_STRIDE;
  if (nv > 0) {
    if (tid < nv) tail_publish(row + tid, x);
  } else if (tid == 0) {
    tail_publish(row, 1.0);
  }
}

// CTA 0 only, after its own tile: consume all rows, then fold / exchange / expand.
template <int BAR, int NT>
__device__ __noinline__ void tail_watch(const TailParams& t, int64_t nblk, int nv, int tid) {
  __shared__ double s_red[NT / 32][PSB_ENERGY_STRIDE];
  __shared__ double s_total[PSB_ENERGY_STRIDE];
  const int nvals = t.ns > 0 ? 2 * t.ns + 1 : 0;   // values of the energy sums
  const int nread = nv > 0 ? nv : 1;               // values per row of THIS launch
  const int lane = tid & 31, warp = tid >> 5;
  double acc[2 * PSB_MAX_SCALARS + 1];
#pragma unroll
  for (int r = 0; r < 2 * PSB_MAX_SCALARS + 1; ++r) acc[r] = 0.0;
  const double sentinel = __longlong_as_double((long long)TAIL_SENTINEL);
  for (int64_t row = tid; row < nblk; row += NT) {
    double* p = t.rows + row * PSB_ENERGY_STRIDE;
#pragma unroll
    for (int r = 0; r < 2 * PSB_MAX_SCALARS + 1; ++r) {
      if (r < nread) {
        unsigned long long v = tail_peek(p + r);
        unsigned ns_sleep = 32;
        while (v == TAIL_SENTINEL) {
          __nanosleep(ns_sleep);
          if (ns_sleep < 512) ns_sleep *= 2;
          v = tail_peek(p + r);
        }
        acc[r] += __longlong_as_double((long long)v);
        tail_publish(p + r, sentinel);   // re-arm for the next launch
      }
    }
  }
  if (t.mode & TAIL_FOLD) {
    // fixed tree: lanes (xor shuffle), then warps in index order
#pragma unroll
    for (int r = 0; r < 2 * PSB_MAX_SCALARS + 1; ++r) {
      if (r < nv) {
        double x = acc[r];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if (lane == 0) s_red[warp][r] = x;
      }
    }
    tail_bar<BAR, NT>();
    if (tid < nv) {
      double x = 0.0;
      for (int w = 0; w < NT / 32; ++w) x += s_red[w][tid];
      t.sums[tid] = x;
      s_total[tid] = x;
    }
  } else {
    tail_bar<BAR, NT>();   // every thread has seen its rows
    if (tid < nvals) s_total[tid] = t.sums[tid];   // folded by an earlier kernel of the stage
  }
  tail_bar<BAR, NT>();
  if (!(t.mode & (TAIL_EXPAND | TAIL_PEER))) return;
  if (t.mode & TAIL_PEER) {
    if (tid < nvals) s_red[0][tid] = s_total[tid];
    tail_bar<BAR, NT>();
    tail_peer_exchange<BAR, NT>(t.peers, s_red[0], nvals, s_total, tid);
  }
  if ((t.mode & TAIL_EXPAND) && tid == 0)
    tail_expansion_update(s_total, t.ns, t.grid_size, t.mpl, t.dt, t.A, t.B, t.state,
                          t.energies);
}

}  // namespace psb
