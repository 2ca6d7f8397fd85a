"""
Generated right-hand sides inside the fused TMA pipeline (row f-2), everything that needs
no GPU: recognition, the functor's arithmetic (compiled as host C++ and compared with the
symbolic right-hand side), NVRTC compilation of the pipeline with the functor for the
configurations the dispatcher picks (dry-run plans), and the stepper wiring.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

import pystella_b200 as ps
from pystella_b200 import _lib, fused_jit
from pystella_b200.field import evaluate, var

cos, tanh, exp = var("cos"), var("tanh"), var("exp")
lam = var("lam")


def attractor(f):          # alpha-attractor T-model + an axion-like coupling
    return lam * tanh(f[0])**2 + (1 - cos(f[1])) * f[0]**2 / 2


def test_recognition_and_refusals():
    sector = ps.ScalarSector(2, potential=attractor)
    system = fused_jit.recognise_general_system(sector.rhs_dict, sector.potential(sector.f))
    assert (system.nscalars, system.user_scalars) == (2, ["lam"])
    assert system.f_name == "f" and system.dot_name == "dfdt" and system.lap_name == "lap_f"
    src = system.generated.source
    assert "#define PSB_GENERATED_RHS 1" in src and '#include "fused_stage_ws.cuh"' in src
    assert "tanh(f[0])" in src and "cos(f[1])" in src and "((T)s[0])" in src

    f = ps.DynamicField("f", offset="h")
    g = ps.Field("g", offset="h")
    with pytest.raises(NotImplementedError, match="array 'g'"):
        fused_jit.recognise_general_system({f: f.dot, f.dot: f.lap + g})
    from pystella_b200.field import shift_fields
    with pytest.raises(NotImplementedError, match="shifted"):
        fused_jit.recognise_general_system({f: f.dot, f.dot: shift_fields(f, (1, 0, 0))})
    with pytest.raises(NotImplementedError):
        fused_jit.recognise_general_system({g: g})

    # stepper wiring: the default fallback of fused steppers (jit=False switches it off)
    kw = dict(halo_shape=2, dt=0.01, dx=(0.1,) * 3, MapKernel=ps.FusedStencilStage)
    with pytest.raises(NotImplementedError):
        ps.LowStorageRK54(sector, jit=False, **kw).systems
    st = ps.LowStorageRK54(sector, **kw)
    assert isinstance(st.systems[0], fused_jit.GeneralScalarSystem)
    assert st.systems[0].potential is not None
    with pytest.raises(NotImplementedError):
        ps.LowStorageRK54(sector, halo_shape=2, dt=0.01, jit=True).systems   # not fused
    # polynomial sectors keep the hand-written kernel even with jit=True
    poly = ps.ScalarSector(1, potential=lambda g_: g_[0]**2 / 2)
    assert not hasattr(ps.LowStorageRK54(poly, jit=True, **kw).systems[0], "generated")


def test_functor_arithmetic_on_host():
    """psb_generated_rhs / psb_generated_potential == the symbolic expressions."""
    sector = ps.ScalarSector(2, potential=attractor)
    system = fused_jit.recognise_general_system(sector.rhs_dict, sector.potential(sector.f))
    src = system.generated.source
    src = src.replace('#include "fused_stage_ws.cuh"', "")
    src = src.replace("__device__ __forceinline__", "static inline")
    src = "#include <cmath>\nusing namespace std;\n" + src + '''
extern "C" void run(const double* f, const double* d, const double* lap, double a, double H,
                    const double* s, double* out) {
  for (int n = 0; n < 2; ++n) psb_generated_rhs<double>(n, f, d, lap, a, H, s, out[2 * n],
                                                        out[2 * n + 1]);
  out[4] = psb_generated_potential<double>(f, a, H, s);
}
'''
    rng = np.random.default_rng(0)
    with tempfile.TemporaryDirectory() as tmp:
        cpp, so = os.path.join(tmp, "f.cpp"), os.path.join(tmp, "f.so")
        open(cpp, "w").write(src)
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", cpp,
                               "-o", so])
        lib = ctypes.CDLL(so)
        lib.run.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_double] * 2 + [ctypes.c_void_p] * 2
        f = ps.DynamicField("f", offset="h", shape=(2,))
        for _ in range(5):
            fv, dv, lv = rng.random(2) + 0.2, rng.random(2), rng.random(2)
            a, hub, lam_v = 1.3, 0.21, 0.7
            out = np.zeros(5)
            s = np.array([lam_v])
            lib.run(fv.ctypes.data, dv.ctypes.data, lv.ctypes.data, a, hub, s.ctypes.data,
                    out.ctypes.data)
            ctx = dict(f=fv, dfdt=dv, lap_f=lv, a=a, hubble=hub, lam=lam_v)
            for n in range(2):
                assert np.isclose(out[2 * n], evaluate(sector.rhs_dict[f[n]], ctx), rtol=1e-14)
                assert np.isclose(out[2 * n + 1], evaluate(sector.rhs_dict[f.dot[n]], ctx),
                                  rtol=1e-13)
            assert np.isclose(out[4], evaluate(sector.potential(f), ctx), rtol=1e-14)


def _plan(dtype, nscalars, N, h, energy):
    """Dry-run plan of the configuration the dispatcher picks for this layout."""
    lib = _lib.load()
    a = _lib.ScalarStageArgs()
    a.dtype, a.nscalars, a.h = dtype, nscalars, h
    a.Nx = a.Ny = a.Nz = N
    a.x_begin, a.x_end = 0, N
    a.field_stride = (N + 2 * h) ** 3
    for i, name in enumerate(("f_in", "f_out", "dfdt", "kf", "kdfdt")):
        setattr(a, name, 0x10000 * (i + 1))
    for d in range(3):
        a.dx[d] = 0.1
    a.dt, a.A, a.B = 0.01, -0.4, 0.3
    if energy:
        a.energy_partials = 0x900000
    plan = _lib.FusedPlan()
    plan.reserved = 1
    rc = lib.psb_scalar_stage_fused_plan(ctypes.byref(a), ctypes.byref(plan))
    assert rc == 0, lib.psb_last_error()
    return plan


@pytest.mark.parametrize("dtype, nscalars, h, energy", [
    (_lib.PSB_F64, 1, 2, False), (_lib.PSB_F64, 2, 2, True), (_lib.PSB_F32, 2, 2, True),
    (_lib.PSB_F64, 3, 4, True), (_lib.PSB_F32, 1, 4, False)])
def test_pipeline_compiles_with_generated_rhs(dtype, nscalars, h, energy, tmp_path):
    """The hand-written pipeline source + the generated functor -> sm_100a cubin with the
    TMA / mbarrier instructions of the production kernel (no GPU needed)."""
    if nscalars == 2:
        sector = ps.ScalarSector(2, potential=attractor)
    elif nscalars == 3:
        sector = ps.ScalarSector(3, potential=lambda f: exp(-f[0]**2) * f[1]**2 + cos(f[2]))
    else:
        sector = ps.ScalarSector(1, potential=lambda f: 1 - cos(f[0]))
    system = fused_jit.recognise_general_system(sector.rhs_dict, sector.potential(sector.f))
    plan = _plan(dtype, nscalars, 64, h, energy)
    assert (plan.h, plan.nscalars, plan.energy) == (h, nscalars, int(energy))
    assert plan.threads == plan.by * plan.tz + 32
    assert plan.params_bytes <= 4096 and plan.smem_bytes < 227 * 1024
    handle = system.generated.compile(plan)
    assert handle > 0 and system.generated.compile(plan) == handle          # cached
    from pystella_b200.codegen import JitLibrary
    jit = JitLibrary.get()
    symbol = jit.psb_jit_kernel_symbol(handle).decode()
    assert "fused_stage_ws_kernel" in symbol and "PsbJitScalars" in symbol
    cubin = str(tmp_path / "k.cubin")
    assert jit.psb_jit_write_cubin(handle, cubin.encode()) == 0
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass and "SYNCS" in sass       # TMA loads, mbarrier waits
    res = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True,
                         text=True).stdout
    regs = int(res.split("REG:")[1].split()[0])
    assert regs * plan.threads * plan.minb <= 65536, (regs, plan.threads, plan.minb)


def test_odd_radius_is_refused():
    """Odd stencil radii run on the cooperative loader, which has no generated variant."""
    lib = _lib.load()
    a = _lib.ScalarStageArgs()
    a.dtype, a.nscalars, a.h = _lib.PSB_F64, 1, 1
    a.Nx = a.Ny = a.Nz = 64
    a.x_end = 64
    a.field_stride = 66 ** 3
    for i, name in enumerate(("f_in", "f_out", "dfdt", "kf", "kdfdt")):
        setattr(a, name, 0x10000 * (i + 1))
    for d in range(3):
        a.dx[d] = 0.1
    plan = _lib.FusedPlan()
    plan.reserved = 1
    assert lib.psb_scalar_stage_fused_plan(ctypes.byref(a), ctypes.byref(plan)) == 3
    assert b"cooperative loader" in lib.psb_last_error()


def test_compile_errors_are_reported():
    sector = ps.ScalarSector(1, potential=lambda f: 1 - cos(f[0]))
    system = fused_jit.recognise_general_system(sector.rhs_dict)
    system.generated.source = system.generated.source.replace("rhs_f = d[0];", "rhs_f = oops;")
    with pytest.raises(ps.PsbError, match="oops"):
        system.generated.compile(_plan(_lib.PSB_F64, 1, 64, 2, False))
