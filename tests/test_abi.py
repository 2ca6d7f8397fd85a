"""The C-ABI library loads and exports every symbol include/pystella_b200.h declares."""
import ctypes
import os
import re

import pytest

from pystella_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "pystella_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psb_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("strict", [False, True])
def test_library_exports_header_symbols(strict):
    names = header_functions()
    assert len(names) >= 12
    lib = ctypes.CDLL(_lib.lib_path(strict))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SYMBOLS) == names


def test_loader_and_error_channel():
    lib = _lib.load()
    assert lib.psb_abi_version() == 2
    assert lib.psb_strict_fp() == 0
    assert _lib.load(strict=True).psb_strict_fp() == 1
    # argument validation happens before any CUDA call: safe without a GPU
    rc = lib.psb_scalar_stage_fused(None, None)
    assert rc != 0 and b"null args" in lib.psb_last_error()
    args = _lib.ScalarStageArgs()
    rc = lib.psb_scalar_stage_fused(ctypes.byref(args), None)
    assert rc != 0 and b"null array" in lib.psb_last_error()
    rc = lib.psb_halo_wrap(None, 0, 4, 4, 4, 1, 1, 1, 7, None)
    assert rc != 0
    assert lib.psb_reduce_workspace_doubles(4) > 0


def test_struct_layout_matches_header():
    """sizeof of the ctypes mirror must equal the C struct (checked via a tiny C probe)."""
    import subprocess
    import tempfile
    src = ('#include <stdio.h>\n#include "pystella_b200.h"\nint main(){printf("%zu %zu\\n",'
           'sizeof(psb_scalar_stage_args), sizeof(psb_reduce_args));'
           'printf("%zu %zu %zu %zu\\n", sizeof(psb_tensor_stage_args), sizeof(psb_peer_args),'
           'sizeof(psb_rk_stage_args), sizeof(psb_tensor_fused_args));'
           'printf("%zu %zu %zu\\n", sizeof(psb_fused_plan), '
           '__builtin_offsetof(psb_fused_plan, tensor_maps), '
           '__builtin_offsetof(psb_fused_plan, params));'
           'printf("%zu %zu %zu\\n", sizeof(psb_stage_tail), '
           '__builtin_offsetof(psb_scalar_stage_args, tail), '
           '__builtin_offsetof(psb_tensor_fused_args, tail));'
           'printf("%zu %zu %zu %d\\n", __builtin_offsetof(psb_scalar_stage_args, launch_flags), '
           '__builtin_offsetof(psb_tensor_fused_args, launch_flags), '
           '__builtin_offsetof(psb_scalar_stage_args, pws_sched), PSB_LAUNCH_PDL);'
           'printf("%zu %zu\\n", sizeof(psb_bin_power_args), '
           '__builtin_offsetof(psb_bin_power_args, spectrum));'
           'return 0;}')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        (a, b, c, d, e, f, g, off_maps, off_params, tail_size, off_tail,
         off_ttail, off_lf, off_tlf, off_sched, pdl, bp_size,
         bp_spec) = subprocess.check_output([exe]).split()
    assert int(bp_size) == ctypes.sizeof(_lib.BinPowerArgs)
    assert int(bp_spec) == _lib.BinPowerArgs.spectrum.offset
    assert int(a) == ctypes.sizeof(_lib.ScalarStageArgs)
    assert int(b) == ctypes.sizeof(_lib.ReduceArgs)
    assert int(c) == ctypes.sizeof(_lib.TensorStageArgs)
    assert int(d) == ctypes.sizeof(_lib.PeerArgs)
    assert int(e) == ctypes.sizeof(_lib.RKStageArgs)
    assert int(f) == ctypes.sizeof(_lib.TensorFusedArgs)
    assert int(g) == ctypes.sizeof(_lib.FusedPlan)
    assert int(off_maps) == _lib.FusedPlan.tensor_maps.offset == 64
    assert int(off_params) == _lib.FusedPlan.params.offset == 576
    assert int(tail_size) == ctypes.sizeof(_lib.StageTail)
    assert int(off_tail) == _lib.ScalarStageArgs.tail.offset
    assert int(off_ttail) == _lib.TensorFusedArgs.tail.offset
    assert int(off_lf) == _lib.ScalarStageArgs.launch_flags.offset
    assert int(off_tlf) == _lib.TensorFusedArgs.launch_flags.offset
    assert int(off_sched) == _lib.ScalarStageArgs.pws_sched.offset
    assert int(pdl) == _lib.LAUNCH_PDL


def test_no_cpu_fallback():
    import torch
    import pystella_b200 as ps
    from pystella_b200 import _kernels
    t = torch.zeros(6, 6, 6, dtype=torch.float64)
    with pytest.raises(ps.PsbError):
        _kernels.halo_wrap(t, (1, 1, 1))
    with pytest.raises(ps.PsbError):
        _kernels.finite_difference(t, lap=torch.zeros(4, 4, 4, dtype=torch.float64), h=1,
                                   dx=(1, 1, 1))
    d = ps.DomainDecomposition((1, 1, 1), 1)
    with pytest.raises(ps.PsbError):
        d.share_halos(None, t)


def _valid_stage_args():
    """Arguments that pass every check up to the launch (pointers are dummies: the tests
    below break one field each and must be refused BEFORE any CUDA call)."""
    a = _lib.ScalarStageArgs()
    a.dtype, a.nscalars, a.h = 0, 1, 2
    a.Nx = a.Ny = a.Nz = 16
    a.x_begin, a.x_end = 0, 16
    a.field_stride = 20 ** 3
    for name, val in (("f_in", 0x1000), ("f_out", 0x2000), ("dfdt", 0x3000), ("kf", 0x4000),
                      ("kdfdt", 0x5000)):
        setattr(a, name, val)
    for d in range(3):
        a.dx[d] = 0.1
    a.nterms[0] = 1
    a.coef[0][0] = 1.0
    a.power[0][0][0] = 1
    return a


@pytest.mark.parametrize("mutate, message", [
    (lambda a: setattr(a, "h", 5), b"h=5 out of range"),
    (lambda a: setattr(a, "h", 0), b"out of range"),
    (lambda a: setattr(a, "nscalars", 9), b"nscalars=9"),
    (lambda a: setattr(a, "x_end", 17), b"bad x range"),
    (lambda a: (setattr(a, "x2_begin", 4), setattr(a, "x2_end", 8)), b"overlap"),
    (lambda a: setattr(a, "f_out", 0x1000), b"alias"),
    (lambda a: a.dx.__setitem__(1, 0.0), b"dx[1]"),
    (lambda a: a.dx.__setitem__(2, float("nan")), b"dx[2]"),
    (lambda a: a.power[0][0].__setitem__(0, -1), b"negative power"),
    (lambda a: (setattr(a, "nvterms", 1), a.vpower[0].__setitem__(0, -2)), b"negative power"),
    (lambda a: a.nterms.__setitem__(0, 17), b"too many potential terms"),
    (lambda a: (setattr(a, "nscalars", 2), setattr(a, "field_stride", 10)), b"field_stride"),
    (lambda a: (setattr(a, "wrap_x", 1), setattr(a, "peer_lo_f_out", 0x9000)),
     b"mutually exclusive"),
    (lambda a: setattr(a, "dtype", 7), b"unknown dtype"),
    (lambda a: setattr(a, "Nx", 1), b"smaller than the halo"),
])
def test_fused_stage_argument_validation(mutate, message):
    lib = _lib.load()
    a = _valid_stage_args()
    mutate(a)
    rc = lib.psb_scalar_stage_fused(ctypes.byref(a), None)
    assert rc != 0
    assert message in lib.psb_last_error(), lib.psb_last_error()


def test_other_entry_points_validate_before_launching():
    lib = _lib.load()
    a = _valid_stage_args()
    a.f_in = 0
    a.lap_f = 0x6000
    a.power[0][0][0] = -3
    assert lib.psb_scalar_stage_unfused(ctypes.byref(a), None) != 0
    assert b"negative power" in lib.psb_last_error()
    a.power[0][0][0] = 1
    a.nscalars = 0
    assert lib.psb_scalar_stage_unfused(ctypes.byref(a), None) != 0
    assert b"nscalars=0" in lib.psb_last_error()

    r = _lib.ReduceArgs()
    r.dtype, r.narrays, r.nterms, r.noutputs = 0, 1, 1, 1
    r.Nx = r.Ny = r.Nz = 8
    r.arrays[0] = 0x1000
    r.partials, r.out = 0x2000, 0x3000
    r.power[0][0] = -1
    assert lib.psb_reduce(ctypes.byref(r), None) != 0
    assert b"negative power" in lib.psb_last_error()
    r.power[0][0] = 1
    r.halo[0] = -2
    assert lib.psb_reduce(ctypes.byref(r), None) != 0
    assert b"negative halo" in lib.psb_last_error()
    r.halo[0] = 0
    r.op[0] = 5
    assert lib.psb_reduce(ctypes.byref(r), None) != 0
    assert b"unknown op" in lib.psb_last_error()

    k = _lib.RKStageArgs()
    k.dtype, k.nscalars, k.h, k.ncopies = 0, 1, 2, 3
    k.Nx = k.Ny = k.Nz = 8
    k.f, k.dfdt, k.lap_f = 0x1000, 0x2000, 0x3000
    k.nstatements = 1
    k.dst[0], k.ncomb[0] = 1, 1
    k.nterms[0] = 1
    k.power[0][0][0] = -1
    assert lib.psb_scalar_rk_stage(ctypes.byref(k), None) != 0
    assert b"negative power" in lib.psb_last_error()
    k.power[0][0][0] = 1
    k.dst[0] = 3
    assert lib.psb_scalar_rk_stage(ctypes.byref(k), None) != 0
    assert b"bad dst" in lib.psb_last_error()


def _stage_args(nscalars=1, N=(8, 8, 8), h=2):
    """ScalarStageArgs that pass validation up to the stage tail (fake, aligned pointers: the
    calls below must be refused before anything is launched)."""
    a = _lib.ScalarStageArgs()
    a.dtype, a.nscalars, a.h = _lib.PSB_F64, nscalars, h
    a.Nx, a.Ny, a.Nz = N
    a.x_begin, a.x_end = 0, N[0]
    a.field_stride = (N[0] + 2 * h) * (N[1] + 2 * h) * (N[2] + 2 * h)
    for i, name in enumerate(("f_in", "f_out", "dfdt", "kf", "kdfdt")):
        setattr(a, name, 0x10000 * (i + 1))
    for d in range(3):
        a.dx[d] = 0.1
    a.dt, a.A, a.B, a.a = 1e-3, -0.4, 0.3, 1.0
    a.wrap_x = a.wrap_y = a.wrap_z = 1
    return a


def test_stage_tail_validation():
    """psb_stage_tail is checked before the launch (include/pystella_b200.h): unknown mode
    bits, FOLD without energy partials, missing marker rows / sums / state, bad peer tables."""
    lib = _lib.load()
    cases = [
        (dict(mode=64), b"unknown stage-tail mode"),
        (dict(mode=_lib.TAIL_FOLD, nscalars=1, sums=0x9000), b"needs energy_partials"),
        (dict(mode=_lib.TAIL_PEER, nscalars=0), b"marker rows"),
        (dict(mode=_lib.TAIL_PEER, nscalars=0, rows=0x9000, nrows=1), b"marker rows"),
        (dict(mode=_lib.TAIL_EXPAND, nscalars=1, rows=0x9000, nrows=1 << 20), b"sums"),
        (dict(mode=_lib.TAIL_EXPAND, nscalars=1, rows=0x9000, nrows=1 << 20, sums=0x9100),
         b"state and energies"),
        (dict(mode=_lib.TAIL_PEER, nscalars=0, rows=0x9000, nrows=1 << 20), b"rank/nranks"),
    ]
    for fields, msg in cases:
        tail = _lib.StageTail()
        for k, v in fields.items():
            setattr(tail, k, v)
        args = _stage_args()
        args.tail = ctypes.pointer(tail)
        rc = lib.psb_scalar_stage_fused(ctypes.byref(args), None)
        assert rc != 0 and msg in lib.psb_last_error(), (fields, lib.psb_last_error())
    # the same rules guard the tensor stage (no FOLD there: it has no energy partials)
    t = _lib.TensorFusedArgs()
    t.dtype, t.nscalars, t.h = _lib.PSB_F32, 2, 2
    t.Nx, t.Ny, t.Nz = 8, 8, 8
    for i, name in enumerate(("hij_in", "hij_out", "dhijdt", "khij", "kdhijdt", "src")):
        setattr(t, name, 0x10000 * (i + 1))
    for d in range(3):
        t.dx[d] = 0.1
    assert lib.psb_tensor_stage_fused_rows(ctypes.byref(t)) == 6 * 1 * 2 * 1
    tail = _lib.StageTail()
    tail.mode = _lib.TAIL_FOLD
    tail.nscalars = 2
    tail.sums = 0x9000
    tail.rows, tail.nrows = 0x9100, 1 << 20
    t.tail = ctypes.pointer(tail)
    rc = lib.psb_tensor_stage_fused(ctypes.byref(t), None)
    assert rc != 0 and b"needs energy_partials" in lib.psb_last_error()


def test_tail_mode_selection():
    """Which part of the stage tail runs where (pystella_b200/step.py::_make_tail): single GPU
    -- fold + background in the stage's last kernel; x-slab ranks over peer memory -- fold
    only (the exchange kernel does the rest) unless tune['peer_tail']."""
    import torch
    import pystella_b200 as ps
    sector = ps.ScalarSector(1, potential=lambda f: f[0]**2 / 2)

    class FakeExpansion:
        nscalars, grid_size, mpl = 1, 512.0, 1.0
        _A, _B = ps.LowStorageRK54._A, ps.LowStorageRK54._B
        state = torch.zeros(5, dtype=torch.float64)
        energies = torch.zeros(5, dtype=torch.float64)

    def make(**tune):
        st = ps.LowStorageRK54(sector, halo_shape=2, dt=1e-3, dx=(0.1,) * 3,
                               MapKernel=ps.FusedStencilStage, energy=True, tune=tune)
        st.energy_nvals, st._stage_index = 3, 1
        return st

    F, E, P = _lib.TAIL_FOLD, _lib.TAIL_EXPAND, _lib.TAIL_PEER
    st = make()
    assert st._make_tail(fold=True, dt=1e-3, device="cpu").mode == F
    st.attach_expansion(FakeExpansion())
    tail = st._make_tail(fold=True, dt=1e-3, device="cpu")
    assert tail.mode == F | E and tail.nscalars == 1 and tail.A == ps.LowStorageRK54._A[1]
    st._last_of_stage = False                       # a tensor sector follows
    assert st._make_tail(fold=True, dt=1e-3, device="cpu").mode == F
    st._last_of_stage = True
    assert st._make_tail(fold=False, dt=1e-3, nrows=7, device="cpu").mode == E  # tensor kernel

    class FakePeer:
        args = _lib.PeerArgs()
    for tune, want in ((dict(), F), (dict(peer_tail=True), F | E | P)):
        st = make(**tune)
        st._halo_mode, st._peer = "peer", FakePeer()
        st.attach_expansion(FakeExpansion())
        assert st._make_tail(fold=True, dt=1e-3, device="cpu").mode == want
    st = make(peer_tail=True)
    st._halo_mode, st._peer, st.fused_energy = "peer", FakePeer(), False
    tail = st._make_tail(fold=False, dt=1e-3, nrows=5, device="cpu")       # pure barrier
    assert tail.mode == P and tail.nscalars == 0 and tail.nrows >= 5
    rows = st._tail_rows.view(torch.int64)
    assert int(rows[0, 0]) == _lib.TAIL_SENTINEL and bool((rows == _lib.TAIL_SENTINEL).all())
