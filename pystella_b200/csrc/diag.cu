// Diagnostics next to the time loop: the binning pass of
// PowerSpectra (pystella/fourier/spectra.py:103-163).  The Fourier transform itself stays
// with cuFFT (torch.fft), as the reference leaves it to its FFT library.
#include "common.cuh"
#include "diag_kernels.cuh"

namespace psb {

template <typename R> static int launch_bin_power(const psb_bin_power_args* a, cudaStream_t stream) {
  BinPowerParams<R> p;
  size_t smem = 0;
  long long blocks = 1;
  bin_power_setup<R>(a, num_sms(), p, smem, blocks);
  if (p.use_shared)
    bin_power_kernel<R, true><<<dim3((unsigned)blocks), dim3(32, kDiagWarps), smem, stream>>>(p);
  else
    bin_power_kernel<R, false><<<dim3((unsigned)blocks), dim3(32, kDiagWarps), 0, stream>>>(p);
  PSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace psb

extern "C" int psb_bin_power(const psb_bin_power_args* a, void* stream) {
  using namespace psb;
  PSB_REQUIRE(a != nullptr, "psb_bin_power: null args");
  PSB_REQUIRE(a->dtype == PSB_F64 || a->dtype == PSB_F32, "psb_bin_power: bad dtype %d",
              a->dtype);
  PSB_REQUIRE(a->Kx >= 0 && a->Ky >= 0 && a->Kz >= 0, "psb_bin_power: negative extent");
  PSB_REQUIRE(a->num_bins >= 1, "psb_bin_power: num_bins = %d", a->num_bins);
  PSB_REQUIRE(a->bin_width > 0 && a->bin_width == a->bin_width,
              "psb_bin_power: bin_width must be positive");
  PSB_REQUIRE(a->spectrum != nullptr, "psb_bin_power: null spectrum");
  if (a->Kx == 0 || a->Ky == 0 || a->Kz == 0) return 0;
  PSB_REQUIRE(a->fk && a->momenta_x && a->momenta_y && a->momenta_z,
              "psb_bin_power: null input array");
  const size_t es = a->dtype == PSB_F64 ? 16 : 8;
  PSB_REQUIRE(reinterpret_cast<uintptr_t>(a->fk) % es == 0,
              "psb_bin_power: fk is not aligned to its element size");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return a->dtype == PSB_F64 ? launch_bin_power<double>(a, s) : launch_bin_power<float>(a, s);
}
