// Device code of the diagnostics that sit next to the time loop (SURVEY.md section 8, row
// f-4).  Kept free of host headers so that the CPU test-suite can execute it under
// tests/cuda_emu/cuda_emu.h.
#pragma once
#include "hist_accum.cuh"
#include "pystella_b200.h"

namespace psb {

// Binned power spectrum of a momentum-space array (PowerSpectra.bin_power,
// pystella/fourier/spectra.py:103-131):
//     spectrum[round(|k| / bin_width)] += count * |k|^k_power * |fk|^2
// with |k| = sqrt(sum_mu (dk_mu * momenta_mu[index_mu])^2) formed on the fly from the
// three 1-D momenta arrays (as the reference's kernel does) and count = 2 for 0 < k_z <
// Nyquist of an r2c transform.  HBM traffic: fk once = 2 S bytes per mode.
template <typename R> struct Cplx;
template <> struct Cplx<double> { typedef double2 type; };
template <> struct Cplx<float> { typedef float2 type; };

template <typename R> struct BinPowerParams {
  const typename Cplx<R>::type* fk;
  const R* kx;
  const R* ky;
  const R* kz;
  long long Kx, Ky, Kz;
  double dkx, dky, dkz, inv_bin_width, k_power, nyquist;
  int num_bins, is_real, copies, use_shared;
  int ipow;      // k_power when it is a small non-negative integer (no pow()), else -1
  double* spectrum;
};

constexpr int kDiagWarps = 8;      // block = (32, kDiagWarps)

template <typename R>
__device__ __forceinline__ void bin_power_mode(const BinPowerParams<R>& p, double kxy2,
                                               long long k, typename Cplx<R>::type v,
                                               int& bin, double& w) {
  const double mz = (double)p.kz[k];
  const double qz = p.dkz * mz;
  const double kmag = sqrt(kxy2 + qz * qz);
  const double b = floor(kmag * p.inv_bin_width + 0.5);      // C rounding of a value >= 0
  bin = (b >= 0.0 && b < (double)p.num_bins) ? (int)b : -1;
  const double count = (p.is_real && mz > 0.0 && mz < p.nyquist) ? 2.0 : 1.0;
  const double re = (double)v.x, im = (double)v.y;
  double kp = 1.0;
  if (p.ipow >= 0) {
    for (int e = 0; e < p.ipow; ++e) kp *= kmag;   // uniform trip count
  } else {
    kp = pow(kmag, p.k_power);
  }
  w = count * kp * (re * re + im * im);
}

// SHARED: block-private copies of the spectrum in shared memory (the pointer's address
// space is then known at compile time: shared-memory atomics, and RED for the flush);
// otherwise (more bins than fit) global atomics straight away.
template <typename R, bool SHARED>
__global__ void __launch_bounds__(32 * kDiagWarps) bin_power_kernel(const BinPowerParams<R> p) {
#ifdef PSB_CUDA_EMU
  double* temp = emu::dynamic_smem();
#else
  extern __shared__ double psb_diag_smem[];
  double* temp = psb_diag_smem;
#endif
  const int lane = threadIdx.x, warp = threadIdx.y;
  const int tid = warp * 32 + lane, nthreads = 32 * kDiagWarps;
  double* mine;
  if (SHARED) {
    for (int b = tid; b < p.copies * p.num_bins; b += nthreads) temp[b] = 0.0;
    __syncthreads();
    mine = temp + (warp % p.copies) * p.num_bins;
  } else {
    mine = p.spectrum;
  }
  typedef typename Cplx<R>::type C2;
  const long long ngy = (p.Ky + kDiagWarps - 1) / kDiagWarps;
  const long long ngroups = p.Kx * ngy;
  for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const long long i = g / ngy;
    const long long j = (g - i * ngy) * kDiagWarps + warp;
    if (j >= p.Ky) continue;                       // uniform in the warp
    const double qx = p.dkx * (double)p.kx[i], qy = p.dky * (double)p.ky[j];
    const double kxy2 = qx * qx + qy * qy;
    const C2* row = p.fk + (i * p.Ky + j) * p.Kz;
    for (long long k0 = 0; k0 < p.Kz; k0 += 64) {  // two modes in flight per lane
      const long long ka = k0 + lane, kb = k0 + 32 + lane;
      C2 va, vb;
      va.x = va.y = vb.x = vb.y = (R)0;
      if (ka < p.Kz) va = row[ka];
      if (kb < p.Kz) vb = row[kb];
      int bin_a = -1, bin_b = -1;
      double wa = 0.0, wb = 0.0;
      if (ka < p.Kz) bin_power_mode<R>(p, kxy2, ka, va, bin_a, wa);
      if (kb < p.Kz) bin_power_mode<R>(p, kxy2, kb, vb, bin_b, wb);
      hist_accumulate(mine, p.num_bins, bin_a, wa, lane);
      if (k0 + 32 < p.Kz) hist_accumulate(mine, p.num_bins, bin_b, wb, lane);
    }
  }
  if (SHARED) {
    __syncthreads();
    hist_flush(temp, p.copies, p.num_bins, p.spectrum, tid, nthreads);
  }
}

// Host side: kernel parameters and launch geometry for `a` (everything but the launch;
// shared with the CPU execution model of the tests).  Block-private copies of the spectrum
// in shared memory: one per warp while they fit in 48 KB, a single one up to 6144 bins,
// global atomics beyond that.  4 CTAs per SM unless tune_blocks says otherwise.
template <typename R>
inline void bin_power_setup(const psb_bin_power_args* a, int num_sms, BinPowerParams<R>& p,
                            size_t& smem_bytes, long long& blocks) {
  p.fk = static_cast<const typename Cplx<R>::type*>(a->fk);
  p.kx = static_cast<const R*>(a->momenta_x);
  p.ky = static_cast<const R*>(a->momenta_y);
  p.kz = static_cast<const R*>(a->momenta_z);
  p.Kx = a->Kx;
  p.Ky = a->Ky;
  p.Kz = a->Kz;
  p.dkx = a->dk[0];
  p.dky = a->dk[1];
  p.dkz = a->dk[2];
  p.inv_bin_width = 1.0 / a->bin_width;
  p.k_power = a->k_power;
  p.nyquist = a->nyquist;
  p.ipow = -1;
  if (a->k_power >= 0 && a->k_power <= 16 && a->k_power == double(int(a->k_power)))
    p.ipow = int(a->k_power);
  p.num_bins = a->num_bins;
  p.is_real = a->is_real;
  p.spectrum = a->spectrum;
  const size_t bytes_one = size_t(a->num_bins) * sizeof(double);
  const size_t budget = 48 * 1024;
  p.use_shared = bytes_one <= budget ? 1 : 0;
  p.copies = 1;
  if (p.use_shared) {
    p.copies = int(budget / bytes_one);
    if (p.copies > kDiagWarps) p.copies = kDiagWarps;
  }
  smem_bytes = p.use_shared ? bytes_one * p.copies : 0;
  const long long ngroups = a->Kx * ((a->Ky + kDiagWarps - 1) / kDiagWarps);
  blocks = a->tune_blocks > 0 ? a->tune_blocks : 4LL * num_sms;
  if (blocks > ngroups) blocks = ngroups;
  if (blocks < 1) blocks = 1;
}

}  // namespace psb
