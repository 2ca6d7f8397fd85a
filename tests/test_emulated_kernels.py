"""
The GPU parity tests, EXECUTED in the CPU suite: the bodies of tests/test_gpu_kernels.py and
tests/test_gpu_api.py run unchanged against a HOST build of the same CUDA sources
(tests/emu_library.py) on the execution model of tests/cuda_emu/cuda_emu.h, which includes
mbarriers, TMA tile loads (cp.async.bulk.tensor with a software tensor map), named barriers
and warp shuffles.  Covered: halos, element-wise LSRK / classical RK / tensor stages, finite
differences (TMA pipeline and generic tile kernel), reductions, background -- and the hot
kernels themselves: the warp-specialised TMA pipeline of the fused scalar stage
(fused_stage_ws.cuh), its persistent variant (fused_stage_pws.cuh), the fused tensor stage
(tensor_stage_fused.cu) and the generic instance, bit-exact against the oracle
(-ffp-contract=off = the strict build's arithmetic), with the in-kernel stage tail.
On the GPU box the same functions are the hard gates; here they check the kernels' logic
(ring protocol, parities, x chunks, halo write-through, energy rows), the C ABI and the Python
layer above it without a device.  Not modelled: the memory model, asynchrony, performance.  The NVRTC-generated instances of the pipeline run
here as well: their functor source is compiled for the model instead of with NVRTC.
"""
import numpy as np
import pytest
import torch

import test_gpu_api as API
import test_gpu_generated as GEN
import test_gpu_kernels as KER
import test_zz_gpu_diagnostics as ZZ
from emu_device import emulated_device

# a kernel that hangs on the execution model (a protocol bug) must fail the run, not stall it
pytestmark = pytest.mark.timeout(600, method="thread")


@pytest.fixture
def device(monkeypatch):
    def host(a):          # like .cuda(): a copy, never a view of the NumPy original
        return torch.from_numpy(np.array(a, copy=True, order="C"))

    def strip(fn):
        def wrapped(*a, **k):
            k.pop("device", None)
            return fn(*a, **k)
        return wrapped

    for mod in (API, GEN, KER, ZZ):
        monkeypatch.setattr(mod, "to_dev", host)
    for name in ("zeros", "empty", "rand", "randn", "ones", "tensor", "full"):
        monkeypatch.setattr(torch, name, strip(getattr(torch, name)))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)

    # polynomial potentials through the coefficient tables (the NVRTC-generated instances of
    # the pipeline are a GPU-only path)
    monkeypatch.setenv("PSB_JIT", "0")
    with emulated_device(monkeypatch) as dev:
        # the watcher CTA of the fused stage (csrc/stage_tail.cuh) waits for the rows of all
        # other CTAs: on a model that runs blocks one after the other it has to come last
        dev.main.emu_set_reverse_blocks(1)
        yield dev
        dev.main.emu_set_reverse_blocks(0)


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


@pytest.mark.parametrize("h, halo_extra", [(1, 0), (2, 0), (2, 1), (3, 0), (4, 1)])
def test_finite_difference_vs_oracle(device, h, halo_extra):
    KER.test_finite_difference_vs_oracle(h, halo_extra, True)


@pytest.mark.parametrize("h", [1, 2])
def test_plane_wave_eigenvalues(device, h):
    KER.test_plane_wave_eigenvalues_cuda(h)


# ---- randomised sweeps of the stand-alone kernels ------------------------------------------
# (80 cases of each generator were run once while this was written: no failure)

def test_finite_difference_random_sweep(device):
    """Radius 1..4, padding >= radius, every output combination incl. accumulation, random
    extents, fp64 / fp32: TMA pipeline or tile kernel as the layout decides, vs the oracle."""
    import oracle as O
    from pystella_b200 import _kernels as K
    rng = np.random.default_rng(8100)
    for _ in range(8):
        h = int(rng.integers(1, 5))
        halo = h + int(rng.choice([0, 0, 1, 2]))
        N = (int(rng.integers(1, 24)), int(rng.integers(1, 24)), int(rng.integers(1, 100)))
        if rng.random() < 0.6:
            N = (N[0], N[1], N[2] + N[2] % 2)
        dtype = np.float64 if rng.random() < 0.6 else np.float32
        dx = (0.1, 0.23, 0.17)
        fx = rng.random(tuple(m + 2 * halo for m in N)).astype(dtype)
        inner = fx[tuple(slice(halo - h, halo + m + h) for m in N)].astype(np.float64)
        lap_ref, grd_ref = O.laplacian(inner, h, dx), O.gradient(inner, h, dx)
        t = torch.from_numpy(fx.copy())

        def z():
            return torch.from_numpy(rng.random(N).astype(dtype))
        mode = rng.choice(["lap", "grad", "lapgrad", "pdz", "pdy_incr"])
        kw = dict(h=h, halo=halo, dx=dx, strict=True)
        if mode == "lap":
            lap = z()
            K.finite_difference(t, lap=lap, **kw)
            outs = [(lap, lap_ref)]
        elif mode == "grad":
            a, b, c = z(), z(), z()
            K.finite_difference(t, pdx=a, pdy=b, pdz=c, **kw)
            outs = [(a, grd_ref[0]), (b, grd_ref[1]), (c, grd_ref[2])]
        elif mode == "lapgrad":
            lap, a, b, c = z(), z(), z(), z()
            K.finite_difference(t, lap=lap, pdx=a, pdy=b, pdz=c, **kw)
            outs = [(lap, lap_ref), (a, grd_ref[0]), (b, grd_ref[1]), (c, grd_ref[2])]
        elif mode == "pdz":
            a = z()
            K.finite_difference(t, pdz=a, **kw)
            outs = [(a, grd_ref[2])]
        else:
            a = z()
            a0 = a.numpy().astype(np.float64).copy()
            K.finite_difference(t, pdy=a, accumulate=True, **kw)
            outs = [(a, a0 + grd_ref[1])]
        tol = 1e-13 if dtype == np.float64 else 2e-5
        for got, ref in outs:
            err = np.max(np.abs(got.numpy().astype(np.float64) - ref)) / max(np.max(np.abs(ref)),
                                                                            1e-300)
            assert err < tol, (N, h, halo, dtype, mode, err)
        assert np.array_equal(t.numpy(), fx)


def test_reduction_random_sweep(device):
    """1-3 arrays with their own paddings, up to 6 outputs of random polynomials with sum /
    max / min and absolute values, random extents (vectorised and generic kernel)."""
    from pystella_b200 import _kernels as K
    rng = np.random.default_rng(8200)
    for _ in range(8):
        N = (int(rng.integers(1, 20)), int(rng.integers(1, 20)), int(rng.integers(1, 90)))
        if rng.random() < 0.5:
            N = (N[0], N[1], N[2] + N[2] % 2)
        dtype = np.float64 if rng.random() < 0.6 else np.float32
        na = int(rng.integers(1, 4))
        halos = [int(rng.choice([0, 0, 1, 2, 3])) for _ in range(na)]
        arrs = [(rng.random(tuple(m + 2 * h for m in N)) - 0.4).astype(dtype) for h in halos]
        terms, ops, flags = [], [], []
        for r in range(int(rng.integers(1, 7))):
            ops.append(int(rng.choice([0, 0, 1, 2])))
            flags.append(int(rng.random() < 0.3))
            for _t in range(int(rng.integers(1, 4))):
                terms.append((r, float(rng.normal()), [int(rng.integers(0, 4))
                                                       for _a in range(na)]))
        out = K.reduce_poly([torch.from_numpy(a.copy()) for a in arrs], halos, terms, ops,
                            flags, N, strict=True).numpy()
        inner = [a[tuple(slice(h, h + m) for m in N)].astype(np.float64) for a, h in
                 zip(arrs, halos)]
        for r, (op, fl) in enumerate(zip(ops, flags)):
            val = np.zeros(N)
            for (o, c, p) in terms:
                if o == r:
                    t = np.full(N, c)
                    for v, pw in zip(inner, p):
                        t = t * v**pw
                    val = val + t
            if fl & 1:
                val = np.abs(val)
            ref = {0: val.sum(), 1: val.max(), 2: val.min()}[op]
            scale = max(abs(ref), np.abs(val).sum() if op == 0 else np.abs(val).max(), 1e-300)
            assert abs(out[r] - ref) <= (1e-12 if dtype == np.float64 else 3e-5) * scale, \
                (N, dtype, halos, ops, flags, r)


def test_halo_kernels_random_sweep(device):
    """Per-axis paddings 0..3 (zeros included), random extents: periodic wrap equals a globally
    periodic array bit for bit; face packing and padding removal."""
    from pystella_b200 import _kernels as K
    rng = np.random.default_rng(8300)
    for _ in range(10):
        hh = tuple(int(rng.integers(0, 4)) for _a in range(3))
        N = tuple(int(rng.integers(max(h, 1), 40)) for h in hh)
        dtype = np.float64 if rng.random() < 0.5 else np.float32
        data = KER.periodic_global(rng, N, hh, dtype)
        sub = np.zeros_like(data)
        slc = tuple(slice(hi, -hi) if hi > 0 else slice(None) for hi in hh)
        sub[slc] = data[slc]
        t = torch.from_numpy(sub.copy())
        K.halo_wrap(t, hh)
        assert np.array_equal(t.numpy(), data), (N, hh, dtype)
        arr = rng.random(tuple(m + 2 * x for m, x in zip(N, hh))).astype(dtype)
        tt = torch.from_numpy(arr.copy())
        for axis in (0, 1):
            if hh[axis] == 0:
                continue
            shape = ((2, hh[0], N[1] + 2 * hh[1], N[2] + 2 * hh[2]) if axis == 0
                     else (2, N[0] + 2 * hh[0], hh[1], N[2] + 2 * hh[2]))
            buf = torch.zeros(shape, dtype=tt.dtype)
            K.halo_pack(tt, buf, hh, axis)
            b = buf.numpy()
            if axis == 0:
                assert np.array_equal(b[1], arr[N[0]:N[0] + hh[0]])
                assert np.array_equal(b[0], arr[hh[0]:2 * hh[0]])
            else:
                assert np.array_equal(b[1], arr[:, N[1]:N[1] + hh[1]])
                assert np.array_equal(b[0], arr[:, hh[1]:2 * hh[1]])
        un = torch.zeros(N, dtype=tt.dtype)
        K.remove_halos(tt, un, hh)
        assert np.array_equal(un.numpy(), arr[tuple(slice(h, h + m) for h, m in zip(hh, N))])


# ---- the reference-facing Python API ----------------------------------------------------

@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12)])
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


@pytest.mark.parametrize("name", ["RungeKutta3SSP"])
def test_classical_rk_wave_equation(device, name):
    API.test_classical_rk_wave_equation_device(name)


# ---- the fused stages ---------------------------------------------------------------------

@pytest.mark.parametrize("N,h,n,pot,variant,stage",
                         [c + (2,) for c in KER.CASES] + [c + (0,) for c in KER.CASES[::3]])
def test_fused_stage_bit_exact(device, N, h, n, pot, variant, stage):
    """Every case of the GPU test: the TMA pipeline for even radii and aligned rows, the
    generic instance (cooperative loader / scalar) for the others."""
    KER.test_fused_stage_fp64(N, h, n, pot, variant, True, stage)


@pytest.mark.parametrize("N,n,pot", [((32, 32, 32), 1, KER.POT1), ((24, 20, 64), 2, KER.POT2)])
def test_persistent_fused_stage_bit_exact(device, N, n, pot):
    """tune_variant 30: persistent CTAs, continuous pipeline, atomically claimed chunks."""
    KER.test_fused_stage_fp64(N, 2, n, pot, 30, True, 2)


@pytest.mark.parametrize("seed", [1, 2, 3, 5, 8, 13])
def test_fused_stage_random_sweep(device, seed):
    """Randomised cases of the fused stage beyond the fixed list: radius 1..4, one to four
    fields, random extents (odd ones included), automatic / cooperative / scalar / persistent
    variants, uniform and split x-chunk schedules, first and later stages, with and without
    the periodic wrap in x -- bit-exact against the oracle, write-through halos included.
    (160 cases of the same generator were run once while this was written: no failure.)"""
    import oracle as O
    from pystella_b200 import _kernels as K
    rng = np.random.default_rng(4000 + seed)
    h = int(rng.choice([1, 2, 2, 2, 3, 4]))
    n = int(rng.choice([1, 1, 2, 3, 4]))
    N = (int(rng.integers(h, 50)), int(rng.integers(h, 30)), int(rng.integers(max(h, 2), 150)))
    if rng.random() < 0.7:
        N = (N[0], N[1], N[2] + N[2] % 2)
    variant = int(rng.choice([0, 0, 0, 2, 3, 30]))
    if variant == 30 and (h != 2 or n > 2):
        variant = 0
    xchunk = int(rng.choice([0, 0, 8, 16, 32, -4, -8]))
    stage = int(rng.choice([0, 1, 4]))
    wrapx = bool(rng.random() < 0.8)
    pots = {1: [None, KER.POT1, [(0.5, (2,))]], 2: [None, KER.POT2], 3: [None], 4: [None]}
    pot = pots[n][int(rng.integers(len(pots[n])))]
    f, d, kf, kd = KER._state(rng, n, N, h, np.float64)
    dx, dt, a, hub = (0.1, 0.13, 0.07), 0.01, 1.3, 0.2
    A, B = KER.RK54[0][stage], KER.RK54[1][stage]
    fo, do, kfo, kdo = f.copy(), d.copy(), kf.copy(), kd.copy()
    O.reference_schedule_stage(stage, fo, do, kfo, kdo, np.zeros((n,) + N), h, dx, dt, a=a,
                               hubble=hub, potential_terms=pot)
    tf, td, tkf, tkd = (torch.from_numpy(x.copy()) for x in (f, d, kf, kd))
    tfo = torch.full_like(tf, float("nan"))
    K.scalar_stage(fused=True, f_in=tf, f_out=tfo, dfdt=td, kf=tkf, kdfdt=tkd, h=h, dx=dx,
                   dt=dt, A=A, B=B, a=a, hubble=hub, dV_terms=KER._dV(pot, n), strict=True,
                   variant=variant, xchunk=xchunk, wrap=(wrapx, True, True))
    what = (N, h, n, variant, xchunk, stage, wrapx)
    for got, ref in ((tfo, fo), (td, do), (tkf, kfo), (tkd, kdo)):
        for i in range(n):
            assert np.array_equal(O.interior(got[i].numpy(), h), O.interior(ref[i], h)), what
    full = fo.copy()
    for i in range(n):
        O.share_halos(full[i], h)
    s_ = slice(h, -h)
    faces = [(s_, slice(0, h), s_), (s_, slice(-h, None), s_), (s_, s_, slice(0, h)),
             (s_, s_, slice(-h, None))]
    if wrapx:
        faces += [(slice(0, h), s_, s_), (slice(-h, None), s_, s_)]
    g = tfo.numpy()
    for i in range(n):
        for sl in faces:
            assert np.array_equal(g[i][sl], full[i][sl]), (what, sl)


def test_tma_path_is_the_one_that_ran(device):
    """With tensor-map encoding switched off the automatic variant has nothing to launch on a
    TMA-eligible layout (-1), while the cooperative-loader variant still runs: the cases above
    did go through the pipeline."""
    from pystella_b200._lib import PsbError
    device.main.emu_set_tma(0)
    try:
        with pytest.raises(PsbError, match="code -1"):
            KER.test_fused_stage_fp64((20, 24, 36), 2, 1, KER.POT1, 0, True, 2)
        KER.test_fused_stage_fp64((20, 24, 36), 2, 1, KER.POT1, 2, True, 2)
    finally:
        device.main.emu_set_tma(1)


def test_fused_x_range_split_equals_whole(device):
    KER.test_fused_x_range_split_equals_whole()


@pytest.mark.parametrize("N,h,n,pot", [((32, 32, 32), 2, 1, KER.POT1), ((17, 9, 33), 1, 1, None)])
def test_fused_stage_fp32(device, N, h, n, pot):
    KER.test_fused_stage_fp32(N, h, n, pot)


@pytest.mark.parametrize("seed", [0, 1, 3, 6])
def test_fused_tensor_stage_random_sweep(device, seed):
    """Config 4's stage (two coupled scalars + tensor sector under FusedStencilStage) on
    random lattices: radius 2 or 4, fp64 / fp32, even and odd row lengths, rows that are and
    are not multiples of 16 bytes (the latter take the element-wise tensor stage fed by
    FiniteDifferencer), two or three RK stages -- against the oracle's scalar_rhs + gw_rhs +
    lsrk_stage (tests/helpers.py::gw_stage_case, as tests/test_gpu_api.py does for one
    lattice).  This sweep found that fp32 lattices with Nz + 2h not a multiple of 4 raised
    instead of taking the element-wise stage (gw._fused_ok; fixed)."""
    from helpers import gw_stage_case
    rng = np.random.default_rng(7000 + seed)
    h = int(rng.choice([2, 2, 4]))
    N = (int(rng.integers(h, 14)), int(rng.integers(h, 14)), int(rng.integers(max(h, 4), 40)))
    if rng.random() < 0.75:
        N = (N[0], N[1], N[2] + N[2] % 2)
    dtype = np.float64 if seed % 2 == 0 else np.float32
    gw_stage_case(N, h, dtype, int(rng.integers(2, 4)), rng,
                  lambda a: torch.from_numpy(np.array(a, copy=True, order="C")))


@pytest.mark.parametrize("N, h, dtype", [((10, 12, 18), 2, np.float32)])
def test_config4_stage_gpu_gate_on_rows_without_tma_description(device, N, h, dtype):
    ZZ.test_config4_stage_on_rows_without_tma_description(N, h, dtype)


@pytest.mark.parametrize("seed", [0, 2, 4, 5])
def test_fused_stepper_api_random_sweep(device, monkeypatch, seed):
    """The fused stepper through the public API on random systems: radius 1..4, one to three
    fields, quadratic / quartic / coupled potentials, generated instances on or off, random
    extents, fp64 / fp32, one or two steps; between steps the field is left alone, replaced by
    a clone (a new tensor: halos must be refilled), rescaled in place (halos stale) or is a
    view into a larger tensor.  Against the oracle's schedule.  (110 cases of the same generator
    were run once while this was written: no failure.)"""
    import oracle as O
    import pystella_b200 as ps
    rng = np.random.default_rng(9000 + seed)
    pots = {1: [([(0.5, (2,))], lambda f: f[0]**2 / 2),
                ([(0.5, (2,)), (0.25, (4,))], lambda f: f[0]**2 / 2 + f[0]**4 / 4)],
            2: [([(0.5, (2, 0)), (0.5, (0, 2)), (0.5, (2, 2))],
                 lambda f: f[0]**2 / 2 + f[1]**2 / 2 + f[0]**2 * f[1]**2 / 2),
                ([(0.5, (2, 0)), (0.3, (0, 2))], lambda f: f[0]**2 / 2 + 0.3 * f[1]**2)],
            3: [([(0.5, (2, 0, 0)), (0.5, (0, 2, 0)), (0.1, (1, 1, 1))],
                 lambda f: f[0]**2 / 2 + f[1]**2 / 2 + 0.1 * f[0] * f[1] * f[2])]}
    h = int(rng.choice([1, 2, 2, 2, 3, 4]))
    n = int(rng.choice([1, 1, 2, 3]))
    N = (int(rng.integers(h, 20)), int(rng.integers(h, 16)), int(rng.integers(max(h, 2), 50)))
    if rng.random() < 0.7:
        N = (N[0], N[1], N[2] + N[2] % 2)
    dtype = np.float64 if rng.random() < 0.7 else np.float32
    monkeypatch.setenv("PSB_JIT", "1" if rng.random() < 0.5 else "0")
    terms, pot = pots[n][int(rng.integers(len(pots[n])))]
    nsteps = int(rng.integers(1, 3))
    edit = str(rng.choice(["none", "clone", "scale", "view"]))
    pad = tuple(m + 2 * h for m in N)
    dx = tuple(5 / m for m in N)
    dt = 0.1 * min(dx)
    f0 = (0.2 + 1e-2 * rng.standard_normal((n,) + pad)).astype(dtype)
    d0 = (-0.1 + 1e-2 * rng.standard_normal((n,) + pad)).astype(dtype)
    fr, dr = f0.astype(np.float64), d0.astype(np.float64)
    kf, kd, lap = np.zeros_like(fr), np.zeros_like(fr), np.zeros((n,) + N)
    stepper = ps.LowStorageRK54(ps.ScalarSector(n, potential=pot), halo_shape=h, dt=dt, dx=dx,
                                decomp=ps.DomainDecomposition((1, 1, 1), h, N),
                                MapKernel=ps.FusedStencilStage)
    if edit == "view":
        big = torch.zeros((n + 1,) + pad, dtype=torch.from_numpy(f0).dtype)
        big[1:] = torch.from_numpy(f0)
        f = big[1:]
    else:
        f = torch.from_numpy(f0.copy())
    d = torch.from_numpy(d0.copy())
    for step in range(nsteps):
        for s_ in range(5):
            O.reference_schedule_stage(s_, fr, dr, kf, kd, lap, h, dx, dt, a=1.05, hubble=0.2,
                                       potential_terms=terms)
            stepper(s_, f=f, dfdt=d, a=1.05, hubble=0.2)
        if step + 1 < nsteps:
            if edit == "clone":
                f = f.clone()
            elif edit == "scale":
                f *= 1.5
                fr *= 1.5
    tol = 1e-12 if dtype == np.float64 else 3e-5
    for i in range(n):
        for got, ref in ((f, fr), (d, dr)):
            g = O.interior(got[i].numpy().astype(np.float64), h)
            r = O.interior(ref[i], h)
            assert np.max(np.abs(g - r)) / np.max(np.abs(r)) < tol, (N, h, n, dtype, edit, i)


def test_fused_two_fields_and_generic_dict(device):
    API.test_scalar_sector_two_fields(True, np.float64, 1e-12)
    API.test_generic_dict_with_friction_matches_sector()


@pytest.mark.parametrize("nscalars", [1, 2])
def test_device_resident_preheating_loop(device, nscalars):
    """Fused stage + in-kernel energy sums + watcher-CTA tail (fold, FLRW background) +
    DeviceExpansion, against the oracle's loop."""
    API.test_device_resident_preheating_loop(nscalars)


@pytest.mark.parametrize("h, N", [(1, (12, 10, 16)), (2, (12, 10, 17)), (3, (8, 12, 14))])
def test_fused_two_fields_on_non_tma_layouts(device, h, N, monkeypatch):
    monkeypatch.setenv("PSB_JIT", "1")      # the generated instance is preferred ...
    ZZ.test_fused_two_fields_on_non_tma_layouts(h, N)   # ... and the tables serve the layout


# ---- NVRTC-generated instances of the pipeline (non-quadratic potentials by default) --------

@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_generated_instance_two_coupled_fields(device, monkeypatch, dtype, tol):
    """Config 4's scalar stage: the warp-specialised pipeline compiled with the potential
    m^2 phi^2 / 2 + g^2 phi^2 chi^2 / 2 inlined (pystella_b200/fused_jit.py).  The functor
    source NVRTC would get is compiled for the model and launched with the library's launch
    plan (FusedParams, four tensor maps, user scalars) -- five stages against the oracle."""
    monkeypatch.setenv("PSB_JIT", "1")
    API.test_scalar_sector_two_fields(True, dtype, tol)
    assert device.jit.launches[-1][0].startswith("psb::fused_stage_ws_kernel<")


def test_generated_instance_cosine_potential_and_device_loop(device, monkeypatch):
    """A non-polynomial right-hand side (tests/test_gpu_generated.py) in the fused pipeline,
    and the device-resident loop (energy sums of the generated potential + in-kernel tail)."""
    monkeypatch.setenv("PSB_JIT", "1")
    GEN._case_generated_rhs("float64", 1e-12)
    API.test_device_resident_preheating_loop(2)
    assert device.jit.launches[-1][0].startswith("psb::fused_stage_ws_kernel<double, 2, 2,")


# ---- several ranks -------------------------------------------------------------------------
# The bodies of the multi-GPU tests (tests/dist_worker.py) on gloo ranks without a GPU.  Every
# rank is a process that runs the stage kernels on the execution model; the ranks really run
# concurrently, so flag barriers wait on one another.

def _run_ranks(case, world, *args):
    import emu_library
    from test_distributed import _run
    emu_library.build()                 # once, before the ranks start
    _run(case, world, *args, timeout=900)


def test_fused_stepper_peer_memory_halos_on_two_ranks():
    """The default multi-GPU schedule: the stage kernels store their boundary planes straight
    into the neighbour's halos (POSIX shared memory stands in for NVLink peer mappings, see
    dist_worker._install_shared_memory_peers), psb_expansion_stage_peer all-gathers the
    energy sums over peer memory, is the inter-rank barrier and advances the background;
    uneven slabs; fields against the global oracle, energies bit-identical on both ranks."""
    _run_ranks("emu_fused", 2, "1", "1", "peer")


def test_config4_structure_over_peer_memory_with_in_kernel_tail():
    """Two coupled scalars + fused tensor stage on two ranks, fp32: both kernels store into
    the neighbour's halos, the tensor kernel's watcher CTA runs the inter-rank exchange and
    the background stage (tune peer_tail); == the C oracle on the global lattice."""
    _run_ranks("emu_fused_gw", 2, "f32", "1")


# ---- the GPU gates of the diagnostics kernels (tests/test_zz_gpu_diagnostics.py) -------------

@pytest.mark.parametrize("grid, is_real, dtype, k_power, bw_scale", [
    ((16, 20, 70), True, "f32", 2.5, 1.0), ((24, 24, 24), True, "f64", 1, 0.002)])
def test_bin_power_gpu_gate(device, grid, is_real, dtype, k_power, bw_scale):
    ZZ.test_bin_power_kernel(grid, is_real, dtype, k_power, bw_scale)


@pytest.mark.parametrize("dtype", ["float32"])
def test_histogram_gpu_gates(device, monkeypatch, dtype):
    monkeypatch.setenv("PSB_JIT", "1")
    ZZ.test_generated_histograms_match_numpy(dtype)
    ZZ.test_histograms_on_device()


@pytest.mark.parametrize("cls", ["Stencil", "StreamingStencil"])
def test_tiled_stencil_gpu_gate(device, monkeypatch, cls):
    monkeypatch.setenv("PSB_JIT", "1")
    ZZ.test_tiled_stencil_map(cls, "float64", 2, (37, 19, 150))


# ---- run-time generated kernels next to the hand-written ones ------------------------------

def test_generated_kernels_of_the_python_layer(device, monkeypatch):
    monkeypatch.setenv("PSB_JIT", "1")
    """Custom-stencil FiniteDifferencer, a classical RK stage for an arbitrary rhs_dict and
    non-polynomial reducers: generated sources on the model, hand-written kernels around
    them from the host build."""
    GEN.test_custom_stencil_differencer()
    GEN.test_classical_rk_generated_stage("RungeKutta2Heun")
    GEN.test_reduction_with_generated_reducers()
