// In-kernel tail of an RK stage: what used to be two or three tiny launches between the
// stage kernels (psb_energy_sums, psb_expansion_stage / psb_expansion_stage_peer) runs in
// the LAST CTA of the stage kernel itself (atomic tickets, threadFenceReduction pattern):
//
//   FOLD    the per-CTA energy partial rows are folded in a fixed order -- groups of 64
//           consecutive CTAs by the last CTA of each group, the group rows by the last
//           group -- into sums[0..2ns] (deterministic: the order of the additions depends
//           on indices only, never on which CTA happens to be last);
//   PEER    all-gather of the sums over peer memory, flag barrier between the GPUs of the
//           node (orders the halo planes this kernel stored into the neighbours' memory
//           before anybody's next stage, and protects the ping-pong buffers);
//   EXPAND  energies with the current scale factor, then one low-storage RK stage of the
//           FLRW background (pystella/expansion.py:119-157).
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
constexpr int TAIL_GROUP = 64;   // CTAs per first-level fold

struct TailParams {
  int mode;             // 0 = no tail
  int ns;               // scalar fields in the energy sums (nvals = 2 ns + 1); 0 = none
  double* partials;     // row 0 of this launch's partial rows (FOLD)
  double* ws;           // sums[16] | group rows [ng][16] | tickets (uint32: global, groups)
  int64_t ws_groups;    // capacity (group rows) of ws; tickets start behind them
  double grid_size, mpl, dt, A, B;
  double* state;        // {a, adot, hubble, ka, kadot}
  double* energies;     // {kinetic.., potential, gradient.., total, pressure}
  psb_peer_args peers;
};

#ifndef __CUDACC_RTC__
inline void fill_tail(TailParams& t, const psb_stage_tail* a, double* partials_row0) {
  memset(&t, 0, sizeof(t));
  if (a == nullptr || a->mode == 0) return;
  t.mode = a->mode;
  t.ns = a->nscalars;
  t.partials = partials_row0;
  t.ws = a->workspace;
  t.ws_groups = a->workspace_groups;
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
  //    halo planes of this kernel, published by every CTA with a system fence ahead of its
  //    ticket) is ordered before it
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

// Called by the NT threads of barrier BAR of EVERY CTA of the launch, after their last
// global store.  blk = linear CTA index, nblk = CTAs in the launch; did_peer_store: this CTA
// stored into a neighbour's memory.
template <int BAR, int NT>
__device__ inline void stage_tail(const TailParams& t, int64_t blk, int64_t nblk, int tid,
                                  bool did_peer_store) {
  __shared__ int s_last;
  __shared__ double s_total[PSB_ENERGY_STRIDE];
  const int nvals = t.ns > 0 ? 2 * t.ns + 1 : 0;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  double* sums = t.ws;
  double* grows = t.ws + PSB_ENERGY_STRIDE;
  unsigned int* tickets =
      reinterpret_cast<unsigned int*>(t.ws + PSB_ENERGY_STRIDE * (1 + t.ws_groups));

  if ((t.mode & TAIL_PEER) && did_peer_store)
    __threadfence_system();
  else
    __threadfence();
  tail_bar<BAR, NT>();

  if (t.mode & TAIL_FOLD) {
    const int64_t ng = (nblk + TAIL_GROUP - 1) / TAIL_GROUP;
    const int64_t g = blk / TAIL_GROUP;
    const int64_t left = nblk - g * TAIL_GROUP;
    const int gsize = left < TAIL_GROUP ? (int)left : TAIL_GROUP;
    if (tid == 0) s_last = (atomicAdd(&tickets[1 + g], 1u) == (unsigned)(gsize - 1));
    tail_bar<BAR, NT>();
    if (!s_last) return;
    __threadfence();
    for (int r = warp; r < nvals; r += NW) {
      double x = 0.0;
      for (int row = lane; row < gsize; row += 32)
        x += __ldcg(&t.partials[(g * TAIL_GROUP + row) * PSB_ENERGY_STRIDE + r]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
      if (lane == 0) grows[g * PSB_ENERGY_STRIDE + r] = x;
    }
    __threadfence();
    tail_bar<BAR, NT>();
    if (tid == 0) {
      tickets[1 + g] = 0u;   // re-arm for the next launch
      s_last = (atomicAdd(&tickets[0], 1u) == (unsigned)(ng - 1));
    }
    tail_bar<BAR, NT>();
    if (!s_last) return;
    __threadfence();
    for (int r = warp; r < nvals; r += NW) {
      double x = 0.0;
      for (int64_t row = lane; row < ng; row += 32)
        x += __ldcg(&grows[row * PSB_ENERGY_STRIDE + r]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
      if (lane == 0) sums[r] = x;
    }
    if (tid == 0) tickets[0] = 0u;
    __threadfence();
    tail_bar<BAR, NT>();
  } else {
    if (tid == 0) s_last = (atomicAdd(&tickets[0], 1u) == (unsigned)(nblk - 1));
    tail_bar<BAR, NT>();
    if (!s_last) return;
    if (tid == 0) tickets[0] = 0u;
    __threadfence();
  }
  if (!(t.mode & (TAIL_EXPAND | TAIL_PEER))) return;

  // ---- this is the last CTA of the launch: every other CTA has published its stores
  if (t.mode & TAIL_PEER) {
    tail_peer_exchange<BAR, NT>(t.peers, sums, nvals, s_total, tid);
  } else {
    if (tid < nvals) s_total[tid] = __ldcg(&sums[tid]);
    tail_bar<BAR, NT>();
  }
  if ((t.mode & TAIL_EXPAND) && tid == 0)
    tail_expansion_update(s_total, t.ns, t.grid_size, t.mpl, t.dt, t.A, t.B, t.state,
                          t.energies);
}

}  // namespace psb
