"""
The reference-facing Python API (Stepper / FiniteDifferencer / DomainDecomposition
/ Reduction) on the GPU against oracle time loops written like the reference's
examples.
"""
import numpy as np
import pytest
import torch

import oracle as O
import pystella_b200 as ps
from helpers import scale_err, to_dev

pytestmark = pytest.mark.gpu


def _oracle_wave_loop(f, dfdt, h, dx, dt, nsteps):
    """examples/wave_equation.py:58-65 with the oracle's kernels."""
    f, dfdt = f.copy(), dfdt.copy()
    kf, kd = np.zeros_like(f), np.zeros_like(f)
    lap = np.zeros((1,) + tuple(n - 2 * h for n in f.shape[-3:]))
    for _ in range(nsteps):
        for s in range(5):
            O.reference_schedule_stage(s, f[None], dfdt[None], kf[None], kd[None], lap, h,
                                       dx, dt)
    return f, dfdt, kf, kd


@pytest.mark.parametrize("N,h", [((32, 32, 32), 1), ((64, 64, 64), 1), ((24, 20, 36), 2)])
@pytest.mark.parametrize("fused", [False, True])
def test_wave_equation_example(N, h, fused):
    """examples/wave_equation.py (config 1): LSRK54, random ICs, several steps."""
    rng = np.random.default_rng(12321)
    dx = tuple(10 / Ni for Ni in N)
    dt = min(dx) / 10
    pad = tuple(n + 2 * h for n in N)
    f0, d0 = rng.random(pad), rng.random(pad)
    nsteps = 4
    fr, dr, kfr, kdr = _oracle_wave_loop(f0, d0, h, dx, dt, nsteps)

    decomp = ps.DomainDecomposition((1, 1, 1), h, N)
    f, dfdt = to_dev(f0), to_dev(d0)
    lap_f = torch.zeros(N, dtype=torch.float64, device="cuda")
    f_ = ps.DynamicField("f", offset="h")
    rhs_dict = {f_: f_.dot, f_.dot: f_.lap}
    if fused:
        stepper = ps.LowStorageRK54(rhs_dict, dt=dt, halo_shape=h,
                                    MapKernel=ps.FusedStencilStage, dx=dx, decomp=decomp)
    else:
        stepper = ps.LowStorageRK54(rhs_dict, dt=dt, halo_shape=h)
    derivs = ps.FiniteDifferencer(decomp, h, dx)
    assert stepper.num_stages == 5 and stepper.expected_order == 4
    assert stepper.num_unknowns == 2
    for _ in range(nsteps):
        for s in range(stepper.num_stages):
            if not fused:
                derivs(None, fx=f, lap=lap_f)
            stepper(s, queue=None, f=f, dfdt=dfdt, lap_f=lap_f)
    for got, ref in ((f, fr), (dfdt, dr), (stepper.tmp_arrays["_f_tmp"], kfr),
                     (stepper.tmp_arrays["_dfdt_tmp"], kdr)):
        assert scale_err(O.interior(got.cpu().numpy(), h), O.interior(ref, h)) < 1e-12
    e_max, _ = O.get_errs(O.interior(f.cpu().numpy(), h), O.interior(fr, h))
    assert e_max < 1e-11
    assert set(stepper.tmp_arrays) == {"_f_tmp", "_dfdt_tmp"}
    assert stepper.tmp_arrays["_f_tmp"].shape == f.shape


def _example_potential(mphi=1.20e-6, mchi=0., gsq=2.5e-7, sigma=0., lambda4=0.):
    def potential(f):
        phi, chi = f[0], f[1]
        unscaled = (mphi**2 / 2 * phi**2 + mchi**2 / 2 * chi**2
                    + gsq / 2 * phi**2 * chi**2 + sigma / 2 * phi * chi**2
                    + lambda4 / 4 * chi**4)
        return unscaled / mphi**2
    terms = [(mphi**2 / 2 / mphi**2, (2, 0)), (mchi**2 / 2 / mphi**2, (0, 2)),
             (gsq / 2 / mphi**2, (2, 2)), (sigma / 2 / mphi**2, (1, 2)),
             (lambda4 / 4 / mphi**2, (0, 4))]
    return potential, terms


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_scalar_sector_two_fields(fused, dtype, tol):
    """ScalarSector(2) with the preheating potential, scalars a / hubble by keyword."""
    rng = np.random.default_rng(49279)
    N, h = (24, 16, 32), 2
    pad = tuple(n + 2 * h for n in N)
    dx = tuple(5 / n for n in N)
    dt = 0.1 * min(dx)
    potential, terms = _example_potential()
    f0 = (np.array([.193, 0.])[:, None, None, None] + 1e-3 * rng.standard_normal((2,) + pad))
    d0 = (np.array([-.142231, 0.])[:, None, None, None]
          + 1e-3 * rng.standard_normal((2,) + pad))
    f0, d0 = f0.astype(dtype), d0.astype(dtype)
    a, hubble = 1.07, 0.45

    fr, dr = f0.astype(np.float64), d0.astype(np.float64)
    kfr, kdr = np.zeros_like(fr), np.zeros_like(fr)
    lap = np.zeros((2,) + N)
    for s in range(5):
        O.reference_schedule_stage(s, fr, dr, kfr, kdr, lap, h, dx, dt, a=a, hubble=hubble,
                                   potential_terms=terms)

    sector = ps.ScalarSector(2, potential=potential)
    decomp = ps.DomainDecomposition((1, 1, 1), h, N)
    kwargs = dict(halo_shape=h, rank_shape=N, dt=dt)
    if fused:
        kwargs.update(MapKernel=ps.FusedStencilStage, dx=dx, decomp=decomp)
    stepper = ps.LowStorageRK54([sector], **kwargs)
    derivs = ps.FiniteDifferencer(decomp, h, dx, rank_shape=N)
    f, dfdt = to_dev(f0), to_dev(d0)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    lap_f = torch.zeros((2,) + N, dtype=tdt, device="cuda")
    for s in range(5):
        if not fused:
            derivs(None, fx=f, lap=lap_f)
        stepper(s, queue=None, a=np.array(a), hubble=np.array([hubble]), f=f, dfdt=dfdt,
                dfdx=None, lap_f=lap_f, hij=None, filter_args=True)
    for got, ref in ((f, fr), (dfdt, dr)):
        for i in range(2):
            assert scale_err(O.interior(got[i].cpu().numpy(), h),
                             O.interior(ref[i], h)) < tol


def test_generic_dict_with_friction_matches_sector():
    """A hand-written rhs_dict equal to ScalarSector(1) is recognised as such."""
    N, h = (16, 16, 16), 2
    dx, dt = (0.2,) * 3, 0.01
    f_ = ps.DynamicField("f", offset="h", shape=(1,))
    H = ps.Field("hubble", indices=[])
    a = ps.Field("a", indices=[])
    rhs = {f_[0]: f_.dot[0],
           f_.dot[0]: f_.lap[0] - 2 * H * f_.dot[0] - a**2 * (f_[0] + 0.3 * f_[0]**3)}
    st1 = ps.LowStorageRK54(rhs, halo_shape=h, dt=dt, MapKernel=ps.FusedStencilStage, dx=dx)
    sec = ps.ScalarSector(1, potential=lambda f: f[0]**2 / 2 + 0.3 / 4 * f[0]**4)
    st2 = ps.LowStorageRK54(sec, halo_shape=h, dt=dt, MapKernel=ps.FusedStencilStage, dx=dx)
    torch.manual_seed(1)
    pad = tuple(n + 2 * h for n in N)
    f0 = torch.rand((1,) + pad, dtype=torch.float64, device="cuda")
    d0 = torch.rand((1,) + pad, dtype=torch.float64, device="cuda")
    outs = []
    for st in (st1, st2):
        f, d = f0.clone(), d0.clone()
        for s in range(5):
            st(s, f=f, dfdt=d, a=1.2, hubble=0.1)
        outs.append((f, d))
    s = (slice(None),) + (slice(h, -h),) * 3
    assert float((outs[0][0] - outs[1][0])[s].abs().max()) < 1e-13
    assert float((outs[0][1] - outs[1][1])[s].abs().max()) < 1e-13


def test_unsupported_systems_fail_loudly():
    f_ = ps.DynamicField("f", offset="h")
    bad = {f_: f_.dot, f_.dot: f_.lap / (1 + f_)}
    st = ps.LowStorageRK54(bad, halo_shape=1, dt=0.1, jit=False)
    t = torch.zeros(6, 6, 6, dtype=torch.float64, device="cuda")
    with pytest.raises(NotImplementedError):
        st(0, f=t, dfdt=t.clone(), lap_f=torch.zeros(4, 4, 4, dtype=torch.float64,
                                                      device="cuda"))
    with pytest.raises(ValueError):
        ps.LowStorageRK54({ps.var("y"): 1}, dt=0.1)
    with pytest.raises(ps.PsbError):
        st2 = ps.LowStorageRK54({f_: f_.dot, f_.dot: f_.lap}, halo_shape=1, dt=0.1)
        c = torch.zeros(6, 6, 6, dtype=torch.float64)
        st2(0, f=c, dfdt=c.clone(), lap_f=torch.zeros(4, 4, 4, dtype=torch.float64))
    with pytest.raises(ValueError):
        ps.LowStorageRK54({f_: f_.dot, f_.dot: f_.lap}, halo_shape=1, dt=0.1,
                          MapKernel=ps.FusedStencilStage)   # no dx


@pytest.mark.parametrize("h", [1, 2])
@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-13), (np.float32, 1e-5)])
def test_scalar_energy_reduction(h, dtype, rtol):
    """test/test_energy.py:37-92."""
    rng = np.random.default_rng(8)
    N = (32, 32, 32)
    grid_size = np.prod(N)
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=N)

    def potential(f):
        phi, chi = f[0], f[1]
        return 1/2 * phi**2 + 1/2 * chi**2 + 1/2 * phi**2 * chi**2

    sector = ps.ScalarSector(2, potential=potential)
    energy_fn = ps.Reduction(decomp, sector, rank_shape=N, grid_size=grid_size,
                             halo_shape=h)
    pad = tuple(n + 2 * h for n in N)
    f = rng.random((2,) + pad).astype(dtype)
    dfdt = rng.random((2,) + pad).astype(dtype)
    lap = rng.random((2,) + N).astype(dtype)
    energy = energy_fn(None, f=to_dev(f), dfdt=to_dev(dfdt), lap_f=to_dev(lap),
                       a=np.array(1.))
    f64, d64, l64 = (x.astype(np.float64) for x in (f, dfdt, lap))
    ref = O.scalar_energy([O.interior(f64[i], h) for i in range(2)],
                          [O.interior(d64[i], h) for i in range(2)], [l64[0], l64[1]], 1.,
                          [(0.5, (2, 0)), (0.5, (0, 2)), (0.5, (2, 2))])
    assert set(energy) == {"kinetic", "potential", "gradient"}
    for key in ref:
        assert energy[key].shape == ref[key].shape
        max_err, _ = O.get_errs(energy[key], ref[key])
        assert max_err < rtol, key
    with_cb = ps.Reduction(decomp, sector, grid_size=grid_size, halo_shape=h,
                           callback=ps.get_rho_and_p)
    e2 = with_cb(None, f=to_dev(f), dfdt=to_dev(dfdt), lap_f=to_dev(lap), a=np.array(1.3))
    r = O.get_rho_and_p(O.scalar_energy(
        [O.interior(f64[i], h) for i in range(2)], [O.interior(d64[i], h) for i in range(2)],
        [l64[0], l64[1]], 1.3, [(0.5, (2, 0)), (0.5, (0, 2)), (0.5, (2, 2))]))
    assert np.isclose(e2["total"], r["total"], rtol=rtol * 10)
    assert np.isclose(e2["pressure"], r["pressure"], rtol=rtol * 100, atol=1e-12)


@pytest.mark.parametrize("op", ["avg", "sum", "max"])
@pytest.mark.parametrize("pass_grid_dims", [True, False])
def test_reduction_with_tmp_instructions(op, pass_grid_dims):
    """test/test_reduction.py:38-84."""
    rng = np.random.default_rng(3)
    N = (32, 16, 24)
    decomp = ps.DomainDecomposition((1, 1, 1), 1, grid_shape=N)
    tmp_insns = [(ps.var("x"), ps.Field("f") / 2 + .31)]
    reducers = {"avg": [(ps.var("x"), op)]}
    if pass_grid_dims:
        reducer = ps.Reduction(decomp, reducers, rank_shape=N, tmp_instructions=tmp_insns,
                               grid_size=np.prod(N))
    else:
        reducer = ps.Reduction(decomp, reducers, tmp_instructions=tmp_insns)
    f = rng.random(N)
    result = reducer(None, f=to_dev(f))
    x = f / 2 + .31
    ref = {"avg": x.sum() / np.prod(N), "sum": x.sum(), "max": x.max()}[op]
    assert np.allclose(result["avg"], ref, rtol=5e-14, atol=0)
    # re-use with a different shape (test_reduction.py:90-131)
    if not pass_grid_dims:
        g = rng.random((8, 8, 16))
        r2 = reducer(None, f=to_dev(g))
        x2 = g / 2 + .31
        ref2 = {"avg": x2.mean(), "sum": x2.sum(), "max": x2.max()}[op]
        assert np.allclose(r2["avg"], ref2, rtol=5e-14, atol=0)
    assert reducer.reduce_array(to_dev(f), "max") == f.max()


@pytest.mark.parametrize("max_min", [False, True])
def test_field_statistics(max_min):
    """test/test_reduction.py:137-177."""
    rng = np.random.default_rng(9)
    N, h = (16, 12, 20), 2
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=N)
    stats = ps.FieldStatistics(decomp, h, rank_shape=N, grid_size=np.prod(N),
                               max_min=max_min)
    f = rng.standard_normal((2, 3) + tuple(n + 2 * h for n in N))
    out = stats(to_dev(f))
    ref = O.field_statistics(f, h, max_min=max_min)
    assert set(out) == set(ref)
    for k in ref:
        assert out[k].shape == (2, 3)
        assert np.allclose(out[k], ref[k], rtol=1e-12, atol=1e-14), k


def test_finite_differencer_api():
    """Outer axes, grd as array / tuple, divergence (derivs.py:339-470)."""
    rng = np.random.default_rng(6)
    N, h = (12, 10, 14), 2
    dx = (0.1, 0.2, 0.3)
    decomp = ps.DomainDecomposition((1, 1, 1), h, rank_shape=N)
    derivs = ps.FiniteDifferencer(decomp, h, dx, rank_shape=N)
    pad = tuple(n + 2 * h for n in N)
    fx = rng.random((2, 3) + pad)
    t = to_dev(fx)
    lap = torch.zeros((2, 3) + N, dtype=torch.float64, device="cuda")
    grd = torch.zeros((2, 3, 3) + N, dtype=torch.float64, device="cuda")
    derivs(None, fx=t, lap=lap, grd=grd)
    got_f = t.cpu().numpy()
    for s in np.ndindex(2, 3):
        ref = fx[s].copy()
        O.share_halos(ref, h)
        assert np.array_equal(got_f[s], ref)       # halos shared in place, bit-exact
        assert scale_err(lap[s].cpu().numpy(), O.laplacian(ref, h, dx)) < 1e-13
        assert scale_err(grd[s].cpu().numpy(), O.gradient(ref, h, dx)) < 1e-13
    only = torch.zeros((2, 3, 3) + N, dtype=torch.float64, device="cuda")
    derivs(None, fx=t, grd=only)
    assert torch.equal(only, grd)
    tup = tuple(torch.zeros((2, 3) + N, dtype=torch.float64, device="cuda") for _ in range(3))
    derivs(None, fx=t, grd=tup)
    for i in range(3):
        assert torch.equal(tup[i], grd[:, :, i])
    vec = to_dev(rng.random((3,) + pad))
    div = torch.zeros(N, dtype=torch.float64, device="cuda")
    derivs.divergence(None, vec, div)
    v = vec.cpu().numpy()
    assert scale_err(div.cpu().numpy(), O.divergence(v, h, dx)) < 1e-13


def test_share_halos_and_padding_api():
    rng = np.random.default_rng(12321)
    from helpers import periodic_global
    for h, grid in ((1, (32, 32, 32)), ((2, 0, 3), (16, 20, 24)), ((0, 2, 1), (16, 20, 24))):
        hh = (h,) * 3 if isinstance(h, int) else h
        decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
        assert decomp.rank_shape == grid and decomp.nranks == 1 and decomp.rank == 0
        data = periodic_global(rng, grid, hh)
        sub = np.zeros_like(data)
        slc = tuple(slice(x, -x) if x > 0 else slice(None) for x in hh)
        sub[slc] = data[slc]
        t = to_dev(sub)
        decomp.share_halos(None, t)
        assert np.array_equal(t.cpu().numpy(), data)
        # different-shaped input is fine (test/test_decomp.py:91-96)
        small = torch.rand(tuple(n // 2 + 2 * x for n, x in zip(grid, hh)),
                           dtype=torch.float64, device="cuda")
        decomp.share_halos(None, small)
        # remove / restore / gather / scatter, all four device/host combinations
        un_d = torch.zeros(grid, dtype=torch.float64, device="cuda")
        un_h = np.zeros(grid)
        decomp.remove_halos(None, t, un_d)
        decomp.remove_halos(None, t, un_h)
        assert np.array_equal(un_d.cpu().numpy(), data[slc]) and np.array_equal(un_h, data[slc])
        decomp.gather_array(None, t, un_h, 0)
        back_d = torch.zeros_like(t)
        back_h = np.zeros_like(data)
        decomp.scatter_array(None, un_d, back_d, 0)
        decomp.scatter_array(None, un_h, back_h, 0)
        assert np.array_equal(back_d.cpu().numpy()[slc], data[slc])
        assert np.array_equal(back_h[slc], data[slc])
        with pytest.raises(ValueError):
            decomp.remove_halos(None, t, np.zeros(grid, np.float32))
    with pytest.raises(NotImplementedError):
        ps.DomainDecomposition((1, 1, 2), 1)
    with pytest.raises(ValueError):
        ps.DomainDecomposition((2, 1, 1), 1)
    with pytest.raises(ValueError):
        ps.DomainDecomposition((1, 1, 1), 1, rank_shape=(4, 4, 4), grid_shape=(4, 4, 4))


def test_fused_stepper_on_side_stream_and_device_scalars():
    """queue = torch.cuda.Stream; a / hubble as 1-element device tensors (no sync)."""
    N, h = (16, 16, 32), 2
    dx, dt = (0.2,) * 3, 0.01
    sec = ps.ScalarSector(1, potential=lambda f: f[0]**2 / 2)
    torch.manual_seed(2)
    pad = tuple(n + 2 * h for n in N)
    f0 = torch.rand((1,) + pad, dtype=torch.float64, device="cuda")
    d0 = torch.rand((1,) + pad, dtype=torch.float64, device="cuda")
    res = []
    for mode in ("host", "device"):
        st = ps.LowStorageRK54(sec, halo_shape=h, dt=dt, MapKernel=ps.FusedStencilStage,
                               dx=dx)
        f, d = f0.clone(), d0.clone()
        q = torch.cuda.Stream() if mode == "device" else None
        if q is not None:
            q.wait_stream(torch.cuda.current_stream())
        a = 1.1 if mode == "host" else torch.tensor([1.1], dtype=torch.float64, device="cuda")
        H = 0.2 if mode == "host" else torch.tensor([0.2], dtype=torch.float64, device="cuda")
        for s in range(5):
            st(s, queue=q, f=f, dfdt=d, a=a, hubble=H)
        if q is not None:
            q.synchronize()
        res.append((f.clone(), d.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


def _oracle_energy(f, dfdt, h, dx, a, terms):
    n = f.shape[0]
    fi, di, laps = [], [], []
    for i in range(n):
        g = f[i].copy()
        O.share_halos(g, h)
        laps.append(O.laplacian(g, h, dx))
        fi.append(O.interior(f[i], h))
        di.append(O.interior(dfdt[i], h))
    return O.get_rho_and_p(O.scalar_energy(fi, di, laps, a, terms))


@pytest.mark.parametrize("nscalars", [1, 2])
def test_device_resident_preheating_loop(nscalars):
    """
    Fused stage with in-kernel energy sums + DeviceExpansion (no host sync) against
    the reference's loop order stepper -> expand.step -> compute_energy
    (examples/scalar_preheating.py:258-264) run with the oracle.
    """
    rng = np.random.default_rng(49279)
    N, h = (16, 24, 32), 2
    pad = tuple(n + 2 * h for n in N)
    dx = tuple(5 / n for n in N)
    dt = 0.1 * min(dx)
    mpl = 1.0
    if nscalars == 1:
        def potential(f):
            return f[0]**2 / 2
        terms = [(0.5, (2,))]
        f0 = (.193 + 1e-3 * rng.standard_normal((1,) + pad))
        d0 = (-.142231 + 1e-3 * rng.standard_normal((1,) + pad))
    else:
        potential, terms = _example_potential()
        f0 = np.array([.193, 0.])[:, None, None, None] + 1e-3 * rng.standard_normal((2,) + pad)
        d0 = np.array([-.142231, 0.])[:, None, None, None] + 1e-3 * rng.standard_normal((2,) + pad)
    gs = float(np.prod(N))

    # --- oracle, reference order
    fr, dr = f0.copy(), d0.copy()
    kfr, kdr = np.zeros_like(fr), np.zeros_like(fr)
    lap = np.zeros((nscalars,) + N)
    energy = _oracle_energy(fr, dr, h, dx, 1., terms)
    e0 = float(energy["total"])
    expand = O.Expansion(energy["total"], "LowStorageRK54", mpl=mpl)
    stage_energies = []
    nsteps = 2
    for _ in range(nsteps):
        for s in range(5):
            stage_energies.append((float(energy["total"]), float(energy["pressure"])))
            O.reference_schedule_stage(s, fr, dr, kfr, kdr, lap, h, dx, dt,
                                       a=expand.a[0], hubble=expand.hubble[0],
                                       potential_terms=terms)
            expand.step(s, energy["total"], energy["pressure"], dt)
            energy = _oracle_energy(fr, dr, h, dx, expand.a[0], terms)

    # --- device-resident
    sector = ps.ScalarSector(nscalars, potential=potential)
    stepper = ps.LowStorageRK54(sector, halo_shape=h, dt=dt, dx=dx,
                                MapKernel=ps.FusedStencilStage, energy=True)
    dexp = ps.DeviceExpansion(e0, ps.LowStorageRK54, mpl=mpl, grid_size=gs,
                              nscalars=nscalars)
    f, dfdt = to_dev(f0), to_dev(d0)
    got_energies = []
    for _ in range(nsteps):
        for s in range(5):
            stepper(s, f=f, dfdt=dfdt, a=dexp.a, hubble=dexp.hubble)
            dexp.step(s, stepper, dt)
            g = dexp.get()
            got_energies.append((g["total"], g["pressure"]))
    for (t0, p0), (t1, p1) in zip(stage_energies, got_energies):
        assert np.isclose(t1, t0, rtol=1e-12) and np.isclose(p1, p0, rtol=1e-11, atol=1e-16)
    g = dexp.get()
    assert np.isclose(g["a"], expand.a[0], rtol=1e-13)
    assert np.isclose(g["adot"], expand.adot[0], rtol=1e-12)
    assert np.isclose(g["hubble"], expand.hubble[0], rtol=1e-12)
    for i in range(nscalars):
        assert scale_err(O.interior(f[i].cpu().numpy(), h), O.interior(fr[i], h)) < 1e-12
        assert scale_err(O.interior(dfdt[i].cpu().numpy(), h), O.interior(dr[i], h)) < 1e-12


@pytest.mark.parametrize("fused_scalars", [False, True])
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 3e-5)])
def test_two_scalars_plus_gw_sector(fused_scalars, dtype, tol):
    """
    Config 4 structure: ScalarSector(2) + TensorPerturbationSector in one stepper
    (examples/scalar_preheating.py:107-114, 258-266), derivatives by
    FiniteDifferencer, against the oracle's scalar_rhs + gw_rhs + lsrk_stage.
    """
    rng = np.random.default_rng(5)
    N, h = (16, 12, 20), 2
    pad = tuple(n + 2 * h for n in N)
    dx = tuple(5 / n for n in N)
    dt = 0.1 * min(dx)
    potential, terms = _example_potential(mphi=1.0, gsq=0.3)
    f0 = (0.2 * rng.standard_normal((2,) + pad)).astype(dtype)
    d0 = (0.2 * rng.standard_normal((2,) + pad)).astype(dtype)
    h0 = (1e-2 * rng.standard_normal((6,) + pad)).astype(dtype)
    dh0 = (1e-2 * rng.standard_normal((6,) + pad)).astype(dtype)
    a, hubble = 1.1, 0.3

    # ---- oracle
    F, D, Hh, DH = (x.astype(np.float64) for x in (f0, d0, h0, dh0))
    kF, kD, kH, kDH = (np.zeros_like(x) for x in (F, D, Hh, DH))
    A_, B_, _ = O.LSRK_TABLES["LowStorageRK54"]

    def derivs_all():
        lapf, grd, laph = [], [], []
        for n in range(2):
            O.share_halos(F[n], h)
            lapf.append(O.laplacian(F[n], h, dx))
            grd.append(O.gradient(F[n], h, dx))
        for c in range(6):
            O.share_halos(Hh[c], h)
            laph.append(O.laplacian(Hh[c], h, dx))
        return lapf, np.stack(grd), laph

    for s in range(5):
        lapf, grd, laph = derivs_all()
        Fi = [O.interior(F[n], h) for n in range(2)]
        Di = [O.interior(D[n], h) for n in range(2)]
        Hi = [O.interior(Hh[c], h) for c in range(6)]
        DHi = [O.interior(DH[c], h) for c in range(6)]
        rhs = [np.array(r) for r in O.scalar_rhs(Fi, Di, lapf, a, hubble, terms)]
        rhs += [np.array(r) for r in O.gw_rhs(DHi, laph, grd, hubble)]
        unk, tmp = [], []
        for n in range(2):
            unk += [Fi[n], Di[n]]
            tmp += [O.interior(kF[n], h), O.interior(kD[n], h)]
        for c in range(6):
            unk += [Hi[c], DHi[c]]
            tmp += [O.interior(kH[c], h), O.interior(kDH[c], h)]
        O.lsrk_stage(A_[s], B_[s], dt, unk, tmp, rhs)

    # ---- device, the reference's loop
    decomp = ps.DomainDecomposition((1, 1, 1), h, N)
    scalar_sector = ps.ScalarSector(2, potential=potential)
    gw_sector = ps.TensorPerturbationSector([scalar_sector])
    kw = dict(halo_shape=h, rank_shape=N, dt=dt)
    if fused_scalars:
        kw.update(MapKernel=ps.FusedStencilStage, dx=dx, decomp=decomp)
    stepper = ps.LowStorageRK54([scalar_sector, gw_sector], **kw)
    assert stepper.num_unknowns == 16
    derivs = ps.FiniteDifferencer(decomp, h, dx, rank_shape=N)
    f, dfdt, hij, dhijdt = (to_dev(x) for x in (f0, d0, h0, dh0))
    tdt = f.dtype
    lap_f = torch.zeros((2,) + N, dtype=tdt, device="cuda")
    dfdx = torch.zeros((2, 3) + N, dtype=tdt, device="cuda")
    lap_hij = torch.zeros((6,) + N, dtype=tdt, device="cuda")
    for s in range(5):
        derivs(None, fx=f, lap=lap_f, grd=dfdx)
        derivs(None, fx=hij, lap=lap_hij)
        stepper(s, queue=None, a=np.array(a), hubble=np.array(hubble), f=f, dfdt=dfdt,
                dfdx=dfdx, lap_f=lap_f, hij=hij, dhijdt=dhijdt, lap_hij=lap_hij,
                filter_args=True)
    assert set(stepper.tmp_arrays) == {"_f_tmp", "_dfdt_tmp", "_hij_tmp", "_dhijdt_tmp"}
    for got, ref in ((f, F), (dfdt, D), (hij, Hh), (dhijdt, DH)):
        for c in range(ref.shape[0]):
            assert scale_err(O.interior(got[c].cpu().numpy(), h),
                             O.interior(ref[c], h)) < tol


@pytest.mark.parametrize("name", ["RungeKutta4", "RungeKutta3SSP", "RungeKutta2Midpoint"])
def test_classical_rk_wave_equation_device(name):
    """
    BASELINE config 1 wording: wave equation, RK4 + 2nd-order Laplacian (h = 1), 64^3.
    Classical RK with storage copies on device arrays vs the oracle's rk_stage.
    """
    rng = np.random.default_rng(12321)
    N, h = ((64, 64, 64) if name == "RungeKutta4" else (24, 24, 24)), 1
    dx = tuple(10 / n for n in N)
    dt = min(dx) / 10
    pad = tuple(n + 2 * h for n in N)
    Stepper = getattr(ps, name)
    nc = Stepper.num_copies
    fq = np.zeros((nc,) + pad); dq = np.zeros((nc,) + pad)
    fq[0], dq[0] = rng.random(pad), rng.random(pad)
    f0, d0 = fq.copy(), dq.copy()

    nsteps = 3
    for _ in range(nsteps):
        for s in range(Stepper.num_stages):
            q = O.rk_rhs_copy(s)
            O.share_halos(fq[q], h)
            lap = O.laplacian(fq[q], h, dx)
            rhs_f = np.array(O.interior(dq[q], h))
            fi = fq[(slice(None),) + (slice(h, -h),) * 3]
            di = dq[(slice(None),) + (slice(h, -h),) * 3]
            O.rk_stage(name, s, dt, fi, rhs_f)
            O.rk_stage(name, s, dt, di, lap)

    decomp = ps.DomainDecomposition((1, 1, 1), h, N)
    derivs = ps.FiniteDifferencer(decomp, h, dx)
    f_ = ps.DynamicField("f", offset="h")
    stepper = Stepper({f_: f_.dot, f_.dot: f_.lap}, dt=dt, halo_shape=h)
    f, dfdt = to_dev(f0), to_dev(d0)
    lap_f = torch.zeros(N, dtype=torch.float64, device="cuda")
    for _ in range(nsteps):
        for s in range(stepper.num_stages):
            q = 0 if s == 0 else 1
            derivs(None, fx=f[q], lap=lap_f)
            stepper(s, queue=None, f=f, dfdt=dfdt, lap_f=lap_f)
    for got, ref in ((f, fq), (dfdt, dq)):
        assert scale_err(O.interior(got[0].cpu().numpy(), h), O.interior(ref[0], h)) < 1e-13


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 3e-5)])
def test_unfused_gw_sector_with_fused_scalars(dtype, tol, monkeypatch):
    """test_two_scalars_plus_gw_sector(fused_scalars=True) runs the fused tensor stage (the
    default under FusedStencilStage: lap_hij / dfdx are never read); PSB_FUSED_GW=0 forces the
    element-wise tensor stage fed by FiniteDifferencer next to the fused scalar stage."""
    monkeypatch.setenv("PSB_FUSED_GW", "0")
    test_two_scalars_plus_gw_sector(True, dtype, tol)


def test_fused_gw_sector_contract():
    """Sector order and the hubble argument are checked: the fused tensor stage
    reads the OLD scalar fields from the scalar stage of the same call."""
    rng = np.random.default_rng(6)
    N, h = (16, 12, 20), 2
    pad = tuple(n + 2 * h for n in N)
    dx = tuple(5 / n for n in N)
    potential, _ = _example_potential(mphi=1.0, gsq=0.3)
    scalar_sector = ps.ScalarSector(2, potential=potential)
    gw_sector = ps.TensorPerturbationSector([scalar_sector])
    arrays = dict(f=to_dev(rng.standard_normal((2,) + pad)),
                  dfdt=to_dev(rng.standard_normal((2,) + pad)),
                  hij=to_dev(rng.standard_normal((6,) + pad)),
                  dhijdt=to_dev(rng.standard_normal((6,) + pad)))
    kw = dict(halo_shape=h, dt=1e-3, dx=dx, MapKernel=ps.FusedStencilStage)
    st = ps.LowStorageRK54([scalar_sector, gw_sector], **kw)
    with pytest.raises(TypeError, match="hubble"):
        st(0, a=np.array(1.0), **arrays)
    st = ps.LowStorageRK54([gw_sector, scalar_sector], **kw)
    with pytest.raises(ValueError, match="ScalarSector before"):
        st(0, a=np.array(1.0), hubble=np.array(0.1), **arrays)


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 5e-14), (np.float32, 1e-5)])
def test_generic_elementwise_map_experimental(dtype, rtol):
    """test/test_elementwise.py:36-90: temporaries first, then statements in order."""
    from pystella_b200.field import var
    rng = np.random.default_rng(1)
    N = (16, 12, 40)
    a, b = var("a"), var("b")
    x, y, z = ps.Field("x"), ps.Field("y"), ps.Field("z")
    tmp_dict = {a[0]: x + 2, a[1]: 2 + x * y, b: x + y / 2}
    map_dict = {x: a[0] * y**2 * x + a[1] * b, z: z + a[1] * b}
    ew_map = ps.ElementWiseMap(map_dict, tmp_instructions=tmp_dict, dtype=dtype, jit=True)
    xh, yh, zh = (rng.random(N).astype(dtype) for _ in range(3))
    xd, yd, zd = to_dev(xh), to_dev(yh), to_dev(zh)
    ew_map(None, x=xd, y=yd, z=zd)
    a0, a1, bb = xh + 2, 2 + xh * yh, xh + yh / 2
    x_true, z_true = a0 * yh**2 * xh + a1 * bb, zh + a1 * bb
    assert O.get_errs(x_true, xd.cpu().numpy())[0] < rtol
    assert O.get_errs(z_true, zd.cpu().numpy())[0] < rtol
    assert torch.equal(yd, to_dev(yh))
    single = ps.ElementWiseMap({x: y + z}, dtype=dtype, jit=True)
    single(None, x=xd, y=yd, z=zd)
    assert O.get_errs(yh + z_true, xd.cpu().numpy())[0] < rtol


def test_jit_stepper_experimental():
    """A system outside the menu (sine-Gordon force) through the generated stage, and
    the wave equation through it == the hand-written element-wise kernel."""
    from pystella_b200.field import var
    rng = np.random.default_rng(2)
    N, h = (16, 16, 32), 2
    dx = (0.2,) * 3
    dt = 0.01
    pad = tuple(n + 2 * h for n in N)
    f0, d0 = rng.random(pad), rng.random(pad)
    decomp = ps.DomainDecomposition((1, 1, 1), h, N)
    derivs = ps.FiniteDifferencer(decomp, h, dx)
    f_ = ps.DynamicField("f", offset="h")
    sin = var("sin")
    stepper = ps.LowStorageRK54({f_: f_.dot, f_.dot: f_.lap - sin(f_)}, dt=dt, halo_shape=h,
                                jit=True)
    f, dfdt = to_dev(f0), to_dev(d0)
    lap_f = torch.zeros(N, dtype=torch.float64, device="cuda")
    fr, dr = f0.copy(), d0.copy()
    kf, kd = np.zeros_like(fr), np.zeros_like(fr)
    A, B, _ = O.LSRK_TABLES["LowStorageRK54"]
    for s in range(5):
        derivs(None, fx=f, lap=lap_f)
        stepper(s, f=f, dfdt=dfdt, lap_f=lap_f)
        O.share_halos(fr, h)
        lap = O.laplacian(fr, h, dx)
        fi, di = O.interior(fr, h), O.interior(dr, h)
        rhs = [np.array(di), lap - np.sin(fi)]
        O.lsrk_stage(A[s], B[s], dt, [fi, di], [O.interior(kf, h), O.interior(kd, h)], rhs)
    assert scale_err(O.interior(f.cpu().numpy(), h), O.interior(fr, h)) < 1e-12
    assert scale_err(O.interior(dfdt.cpu().numpy(), h), O.interior(dr, h)) < 1e-12
