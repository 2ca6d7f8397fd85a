// Device-resident FLRW background: folds the fused stage's energy partials and
// advances (a, adot) by one low-storage RK stage without leaving the GPU.
// Replaces, per stage, Reduction.__call__ -> get_rho_and_p -> Expansion.step
// (pystella/reduction.py:237-255, pystella/sectors.py:211-229,
// pystella/expansion.py:101-157) and the blocking device-to-host reads between.
#include <math.h>

#include "common.cuh"

namespace psb {

// Two-level fold in ONE launch: FOLD_BLOCKS blocks each fold a contiguous share of
// the rows (thread t owns column t % 16 of rows t/16, t/16 + 16, ... of its share, so
// consecutive threads read consecutive doubles); the last block to finish (atomic
// ticket) folds the FOLD_BLOCKS block rows in index order -> deterministic result.
constexpr int FOLD_BLOCKS = 64;
constexpr int FOLD_THREADS = 256;

__global__ void __launch_bounds__(FOLD_THREADS)
    energy_sums_kernel(const double* __restrict__ partials, int64_t nslots, int nvals,
                       double* __restrict__ sums, double* __restrict__ scratch,
                       unsigned int* __restrict__ ticket) {
  constexpr int GROUPS = FOLD_THREADS / PSB_ENERGY_STRIDE;  // 16 rows in parallel
  __shared__ double sh[GROUPS][PSB_ENERGY_STRIDE];
  __shared__ bool last;
  const int col = threadIdx.x % PSB_ENERGY_STRIDE, grp = threadIdx.x / PSB_ENERGY_STRIDE;
  const int64_t per = (nslots + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = per * blockIdx.x;
  const int64_t r1 = r0 + per < nslots ? r0 + per : nslots;
  double x0 = 0.0, x1 = 0.0;
  int64_t row = r0 + grp;
  for (; row + GROUPS < r1; row += 2 * GROUPS) {
    x0 += partials[row * PSB_ENERGY_STRIDE + col];
    x1 += partials[(row + GROUPS) * PSB_ENERGY_STRIDE + col];
  }
  if (row < r1) x0 += partials[row * PSB_ENERGY_STRIDE + col];
  sh[grp][col] = x0 + x1;
  __syncthreads();
  for (int half = GROUPS / 2; half > 0; half >>= 1) {
    if (grp < half) sh[grp][col] += sh[grp + half][col];
    __syncthreads();
  }
  if (threadIdx.x < PSB_ENERGY_STRIDE)
    scratch[blockIdx.x * PSB_ENERGY_STRIDE + threadIdx.x] = sh[0][threadIdx.x];
  __threadfence();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    if (threadIdx.x < nvals) {
      double y = 0.0;
      for (unsigned b = 0; b < gridDim.x; ++b)
        y += scratch[b * PSB_ENERGY_STRIDE + threadIdx.x];
      sums[threadIdx.x] = y;
    }
    if (threadIdx.x == 0) *ticket = 0u;  // re-arm for the next launch
  }
}

struct Energies {
  double total, pressure;
};

__device__ inline Energies form_energies(const double* sums, int ns, double grid_size,
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
  Energies e;
  e.total = kin + pot + grad;
  e.pressure = kin - grad / 3.0 - pot;
  energies[2 * ns + 1] = e.total;
  energies[2 * ns + 2] = e.pressure;
  return e;
}

__global__ void expansion_stage_kernel(const double* sums, int ns, double grid_size,
                                       double mpl, double dt, double A, double B,
                                       double* state, double* energies) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = state[0], adot = state[1], ka = state[3], kadot = state[4];
  const Energies e = form_energies(sums, ns, grid_size, a, energies);
  const double pi = 3.141592653589793;
  // a' = adot ; adot' = 4 pi a^2 / 3 / mpl^2 (rho - 3 P) a   (expansion.py:119-138)
  const double rhs_a = adot;
  const double rhs_adot = 4.0 * pi * (a * a) / 3.0 / (mpl * mpl) *
                          (e.total - 3.0 * e.pressure) * a;
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

// All-gather of the energy sums over peer memory + inter-GPU barrier + expansion
// stage, one CTA.  slots: [2][nranks][STRIDE] doubles, flags: [2][nranks] uint64 on
// every rank (symmetric); parity = epoch & 1 double-buffers both.
__global__ void __launch_bounds__(64)
    expansion_stage_peer_kernel(psb_peer_args peers, const double* sums, int ns,
                                double grid_size, double mpl, double dt, double A, double B,
                                double* state, double* energies) {
  __shared__ double total[PSB_ENERGY_STRIDE];
  // PSB_LAUNCH_PDL: scheduled while the stage kernel drains; nothing is read or signalled
  // before that kernel has completed (its peer stores are what the flag below publishes)
  pdl_launch_dependents();
  pdl_wait();
  const int P = peers.nranks, me = peers.rank;
  const unsigned long long epoch = *peers.epoch + 1ull;
  const int par = (int)(epoch & 1ull);
  const int nvals = ns > 0 ? 2 * ns + 1 : 0;
  const int t = threadIdx.x;
  // 1. my sums -> slot [par][me] of every rank (incl. myself)
  for (int e = t; e < P * nvals; e += blockDim.x) {
    const int q = e / nvals, r = e % nvals;
    double* dst = static_cast<double*>(peers.peer_slots[q]) +
                  ((int64_t)par * P + me) * PSB_ENERGY_STRIDE;
    dst[r] = sums[r];
  }
  __threadfence_system();
  __syncthreads();
  // 2. raise my flag on every rank; everything this rank stored to its peers before
  //    this kernel (the halo planes of the stage kernel) is ordered before it
  if (t < P) {
    unsigned long long* fl = static_cast<unsigned long long*>(peers.peer_flags[t]) +
                             (int64_t)par * P + me;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fl), "l"(epoch) : "memory");
  }
  // 3. wait for every rank's flag
  if (t < P) {
    const unsigned long long* mine =
        static_cast<unsigned long long*>(peers.peer_flags[me]) + (int64_t)par * P + t;
    unsigned long long seen;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
    } while (seen < epoch);
  }
  __threadfence_system();
  __syncthreads();
  if (t == 0) *peers.epoch = epoch;
  if (ns <= 0) return;
  // 4. add the slots in rank order (identical on all ranks)
  if (t < nvals) {
    const volatile double* sl =
        static_cast<double*>(peers.peer_slots[me]) + (int64_t)par * P * PSB_ENERGY_STRIDE;
    double x = 0.0;
    for (int q = 0; q < P; ++q) x += sl[q * PSB_ENERGY_STRIDE + t];
    total[t] = x;
  }
  __syncthreads();
  if (t != 0) return;
  double a = state[0], adot = state[1], ka = state[3], kadot = state[4];
  const Energies e = form_energies(total, ns, grid_size, a, energies);
  const double pi = 3.141592653589793;
  const double rhs_a = adot;
  const double rhs_adot = 4.0 * pi * (a * a) / 3.0 / (mpl * mpl) *
                          (e.total - 3.0 * e.pressure) * a;
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

__global__ void expansion_init_kernel(const double* sums, int ns, double grid_size,
                                      double mpl, double a0, double* state,
                                      double* energies) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const Energies e = form_energies(sums, ns, grid_size, a0, energies);
  const double pi = 3.141592653589793;
  // adot = sqrt(8 pi a^2 / 3 / mpl^2 * rho) * a   (expansion.py:101-117)
  const double adot = sqrt(8.0 * pi * (a0 * a0) / 3.0 / (mpl * mpl) * e.total) * a0;
  state[0] = a0;
  state[1] = adot;
  state[2] = adot / a0;
  state[3] = 0.0;
  state[4] = 0.0;
}

}  // namespace psb

extern "C" int psb_energy_sums(const double* partials, int64_t nslots, int nvals,
                               double* sums, void* stream) {
  using namespace psb;
  PSB_REQUIRE(partials && sums, "psb_energy_sums: null pointer");
  PSB_REQUIRE(nslots > 0 && nvals > 0 && nvals <= PSB_ENERGY_STRIDE,
              "psb_energy_sums: bad sizes (nslots=%lld, nvals=%d)", (long long)nslots, nvals);
  // sums[16 .. 16 + 64*16 + 1): scratch rows and the ticket (zeroed by the caller once)
  double* scratch = sums + PSB_ENERGY_STRIDE;
  unsigned int* ticket =
      reinterpret_cast<unsigned int*>(scratch + FOLD_BLOCKS * PSB_ENERGY_STRIDE);
  int blocks = FOLD_BLOCKS;
  if (nslots < 4 * FOLD_BLOCKS) blocks = 1;
  energy_sums_kernel<<<blocks, FOLD_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      partials, nslots, nvals, sums, scratch, ticket);
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int psb_expansion_stage(const double* sums, int nscalars, double grid_size,
                                   double mpl, double dt, double A, double B, double* state,
                                   double* energies, void* stream) {
  using namespace psb;
  PSB_REQUIRE(sums && state && energies, "psb_expansion_stage: null pointer");
  PSB_REQUIRE(nscalars >= 1 && nscalars <= PSB_MAX_SCALARS,
              "psb_expansion_stage: nscalars=%d out of range", nscalars);
  PSB_REQUIRE(grid_size > 0 && mpl > 0, "psb_expansion_stage: bad grid_size/mpl");
  expansion_stage_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, nscalars, grid_size, mpl, dt, A, B, state, energies);
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int psb_expansion_init(const double* sums, int nscalars, double grid_size,
                                  double mpl, double a0, double* state, double* energies,
                                  void* stream) {
  using namespace psb;
  PSB_REQUIRE(sums && state && energies, "psb_expansion_init: null pointer");
  PSB_REQUIRE(nscalars >= 1 && nscalars <= PSB_MAX_SCALARS,
              "psb_expansion_init: nscalars=%d out of range", nscalars);
  PSB_REQUIRE(grid_size > 0 && mpl > 0 && a0 > 0, "psb_expansion_init: bad scalars");
  expansion_init_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, nscalars, grid_size, mpl, a0, state, energies);
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int psb_expansion_stage_peer_v(const psb_peer_args* peers, const double* sums,
                                          int nscalars, double grid_size, double mpl,
                                          double dt, double A, double B, double* state,
                                          double* energies, int launch_flags,
                                          void* stream) {
  using namespace psb;
  PSB_REQUIRE(peers != nullptr, "psb_expansion_stage_peer_v: null peers");
  PSB_REQUIRE(peers->nranks >= 1 && peers->nranks <= 64 && peers->rank >= 0 &&
                  peers->rank < peers->nranks,
              "psb_expansion_stage_peer_v: bad rank/nranks");
  PSB_REQUIRE(peers->peer_slots && peers->peer_flags && peers->epoch,
              "psb_expansion_stage_peer_v: null peer tables");
  PSB_REQUIRE(nscalars >= 0 && nscalars <= PSB_MAX_SCALARS,
              "psb_expansion_stage_peer_v: nscalars=%d out of range", nscalars);
  if (nscalars > 0) {
    PSB_REQUIRE(sums && state && energies, "psb_expansion_stage_peer_v: null pointer");
    PSB_REQUIRE(grid_size > 0 && mpl > 0, "psb_expansion_stage_peer_v: bad grid_size/mpl");
  }
  PSB_CHECK_CUDA(launch_kernel(expansion_stage_peer_kernel, dim3(1), dim3(64), 0,
                               static_cast<cudaStream_t>(stream), launch_flags, *peers, sums,
                               nscalars, grid_size, mpl, dt, A, B, state, energies));
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int psb_expansion_stage_peer(const psb_peer_args* peers, const double* sums,
                                        int nscalars, double grid_size, double mpl,
                                        double dt, double A, double B, double* state,
                                        double* energies, void* stream) {
  return psb_expansion_stage_peer_v(peers, sums, nscalars, grid_size, mpl, dt, A, B, state,
                                    energies, 0, stream);
}
