// Test driver: runs the device code of pystella_b200/csrc/diag_kernels.cuh on the CPU
// execution model of cuda_emu.h.  Built and called by tests/test_cuda_emu.py.
#include "cuda_emu.h"
#include "diag_kernels.cuh"

template <typename R>
static void run_bin_power(const void* fk, const void* kx, const void* ky, const void* kz,
                          long long Kx, long long Ky, long long Kz, const double* dk,
                          double bin_width, double k_power, double nyquist, int num_bins,
                          int is_real, int copies, int use_shared, int blocks,
                          double* spectrum) {
  psb::BinPowerParams<R> p;
  p.fk = static_cast<const typename psb::Cplx<R>::type*>(fk);
  p.kx = static_cast<const R*>(kx);
  p.ky = static_cast<const R*>(ky);
  p.kz = static_cast<const R*>(kz);
  p.Kx = Kx; p.Ky = Ky; p.Kz = Kz;
  p.dkx = dk[0]; p.dky = dk[1]; p.dkz = dk[2];
  p.inv_bin_width = 1.0 / bin_width;
  p.k_power = k_power;
  p.nyquist = nyquist;
  p.ipow = (k_power >= 0 && k_power <= 16 && k_power == double(int(k_power))) ? int(k_power) : -1;
  p.num_bins = num_bins;
  p.is_real = is_real;
  p.copies = copies;
  p.use_shared = use_shared;
  p.spectrum = spectrum;
  const size_t smem = use_shared ? size_t(copies) * num_bins * sizeof(double) : 0;
  if (use_shared)
    emu::launch_smem(dim3(blocks), dim3(32, psb::kDiagWarps), smem,
                     [&] { psb::bin_power_kernel<R, true>(p); });
  else
    emu::launch_smem(dim3(blocks), dim3(32, psb::kDiagWarps), 0,
                     [&] { psb::bin_power_kernel<R, false>(p); });
}

extern "C" void emu_bin_power(int is_f64, const void* fk, const void* kx, const void* ky,
                              const void* kz, long long Kx, long long Ky, long long Kz,
                              const double* dk, double bin_width, double k_power,
                              double nyquist, int num_bins, int is_real, int copies,
                              int use_shared, int blocks, double* spectrum) {
  if (is_f64)
    run_bin_power<double>(fk, kx, ky, kz, Kx, Ky, Kz, dk, bin_width, k_power, nyquist,
                          num_bins, is_real, copies, use_shared, blocks, spectrum);
  else
    run_bin_power<float>(fk, kx, ky, kz, Kx, Ky, Kz, dk, bin_width, k_power, nyquist,
                         num_bins, is_real, copies, use_shared, blocks, spectrum);
}
