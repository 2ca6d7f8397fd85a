// Device code of the diagnostics that sit next to the time loop (SURVEY.md section 8, row
// f-4).  Kept free of host headers so that the CPU test-suite can execute it under
// tests/cuda_emu/cuda_emu.h.
#pragma once
#include "hist_accum.cuh"

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

}  // namespace psb
