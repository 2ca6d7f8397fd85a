// Test driver: runs the device code AND the launch set-up of pystella_b200/csrc/diag_kernels.cuh
// on the CPU execution model of cuda_emu.h; `emu_psb_bin_power` has the signature of the C-ABI
// entry point psb_bin_power.  Built and called by tests/test_cuda_emu.py.
#include "cuda_emu.h"
#include "diag_kernels.cuh"

template <typename R> static void run_bin_power(const psb_bin_power_args* a, int num_sms) {
  psb::BinPowerParams<R> p;
  size_t smem = 0;
  long long blocks = 1;
  psb::bin_power_setup<R>(a, num_sms, p, smem, blocks);
  if (p.use_shared)
    emu::launch_smem(dim3((unsigned)blocks), dim3(32, psb::kDiagWarps), smem,
                     [&] { psb::bin_power_kernel<R, true>(p); });
  else
    emu::launch_smem(dim3((unsigned)blocks), dim3(32, psb::kDiagWarps), 0,
                     [&] { psb::bin_power_kernel<R, false>(p); });
}

extern "C" int emu_psb_bin_power(const psb_bin_power_args* a, int num_sms) {
  if (a->Kx == 0 || a->Ky == 0 || a->Kz == 0) return 0;
  if (a->dtype == PSB_F64)
    run_bin_power<double>(a, num_sms);
  else
    run_bin_power<float>(a, num_sms);
  return 0;
}
