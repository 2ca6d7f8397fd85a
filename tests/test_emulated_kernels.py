"""
The GPU parity tests of the stand-alone kernels, EXECUTED in the CPU suite: the bodies of
tests/test_gpu_kernels.py and tests/test_gpu_api.py run unchanged against a HOST build of the
same CUDA sources (tests/emu_library.py: halo, element-wise LSRK / classical RK / tensor stages,
finite differences (generic tile kernel), reductions, background; launch syntax transformed,
everything else compiled as it is with -ffp-contract=off, i.e. the strict build's arithmetic).
On the GPU box the same functions are the hard gates; here they check the kernels' logic,
the C ABI and the Python layer above it against the oracle without a device.  The TMA pipelines
(fused stages) exist on the hardware only and are not covered here.
"""
import numpy as np
import pytest
import torch

import test_gpu_api as API
import test_gpu_kernels as KER
from emu_device import emulated_device


@pytest.fixture
def device(monkeypatch):
    def host(a):
        return torch.from_numpy(np.ascontiguousarray(a))

    def strip(fn):
        def wrapped(*a, **k):
            k.pop("device", None)
            return fn(*a, **k)
        return wrapped

    for mod in (API, KER):
        monkeypatch.setattr(mod, "to_dev", host)
    for name in ("zeros", "empty", "rand", "randn", "ones", "tensor", "full"):
        monkeypatch.setattr(torch, name, strip(getattr(torch, name)))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    with emulated_device(monkeypatch) as dev:
        yield dev


# ---- kernels through the thin wrappers of pystella_b200._kernels ------------------------

@pytest.mark.parametrize("strict", [True, False])
def test_unfused_scalar_stage(device, strict):
    KER.test_unfused_stage(strict)


@pytest.mark.parametrize("h, grid_shape", [(1, (32, 32, 32)), (1, (19, 7, 32)),
                                           ((2, 0, 3), (16, 16, 16)), ((0, 2, 1), (8, 8, 8)),
                                           (2, (16, 12, 20)), (4, (9, 8, 11))])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_halo_wrap_bit_exact(device, h, grid_shape, dtype):
    KER.test_halo_wrap_bit_exact(h, grid_shape, dtype)


def test_pack_unpack_and_padding(device):
    KER.test_pack_unpack_and_padding()


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 5e-14), (np.float32, 1e-5)])
def test_reduce_poly(device, dtype, rtol):
    KER.test_reduce_poly(dtype, rtol)


@pytest.mark.parametrize("h", [1, 2, 3, 4])
@pytest.mark.parametrize("halo_extra", [0, 1])
def test_finite_difference_vs_oracle(device, h, halo_extra):
    KER.test_finite_difference_vs_oracle(h, halo_extra, True)


@pytest.mark.parametrize("h", [1, 2])
def test_plane_wave_eigenvalues(device, h):
    KER.test_plane_wave_eigenvalues_cuda(h)


# ---- the reference-facing Python API ----------------------------------------------------

def test_wave_equation_example_unfused(device):
    API.test_wave_equation_example((24, 20, 36), 2, False)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_scalar_sector_two_fields_unfused(device, dtype, tol):
    API.test_scalar_sector_two_fields(False, dtype, tol)


@pytest.mark.parametrize("h", [1, 2])
def test_scalar_energy_reduction(device, h):
    API.test_scalar_energy_reduction(h, np.float64, 1e-13)


@pytest.mark.parametrize("op", ["avg", "max"])
def test_reduction_with_tmp_instructions(device, op):
    API.test_reduction_with_tmp_instructions(op, True)


@pytest.mark.parametrize("max_min", [False, True])
def test_field_statistics(device, max_min):
    API.test_field_statistics(max_min)


def test_finite_differencer_api(device):
    API.test_finite_differencer_api()


def test_share_halos_and_padding_api(device):
    API.test_share_halos_and_padding_api()


def test_two_scalars_plus_gw_sector_unfused(device):
    API.test_two_scalars_plus_gw_sector(False, np.float64, 1e-12)


@pytest.mark.parametrize("name", ["RungeKutta3SSP"])
def test_classical_rk_wave_equation(device, name):
    API.test_classical_rk_wave_equation_device(name)
