"""
Parity of the CUDA kernels (called through the C ABI) against the oracle.
fp64 bar: bit-exact with the strict (-fmad=false) build, <= 1e-12 (array-scale
relative) with the default FMA build; fp32: 1e-5.
"""
import numpy as np
import pytest
import torch

import oracle as O
from pystella_b200 import _kernels as K
from helpers import periodic_global, scale_err, to_dev

pytestmark = pytest.mark.gpu

RK54 = O.LSRK_TABLES["LowStorageRK54"]


def _state(rng, n, N, h, dtype):
    shape = (n,) + tuple(x + 2 * h for x in N)
    f = rng.random(shape).astype(dtype)
    for i in range(n):
        O.share_halos(f[i], h)
    return (f, rng.random(shape).astype(dtype), rng.random(shape).astype(dtype),
            rng.random(shape).astype(dtype))


POT1 = [(0.5, (2,)), (0.25, (4,))]
POT2 = [(0.5, (2, 0)), (0.5, (0, 2)), (0.5, (2, 2))]


def _dV(pot, n):
    return [O.poly_diff(pot, i) for i in range(n)] if pot else None


CASES = [
    # N, h, nscalars, pot, variant
    ((32, 32, 32), 2, 1, POT1, 0),
    ((20, 24, 36), 2, 1, POT1, 0),
    ((20, 24, 36), 2, 1, POT1, 2),     # cooperative (non-TMA) loader
    ((20, 24, 36), 2, 1, POT1, 3),     # scalar (VEC=1) path
    ((17, 9, 33), 1, 1, None, 0),      # ragged, odd sizes (TMA impossible -> fallback)
    ((16, 16, 32), 1, 1, None, 0),
    ((16, 12, 70), 3, 1, POT1, 0),
    ((16, 12, 70), 4, 1, POT1, 0),
    ((4, 4, 4), 2, 1, None, 0),        # smallest grid that still holds the stencil
    ((24, 20, 64), 2, 2, POT2, 0),
    ((12, 10, 16), 2, 3, None, 0),
    ((12, 10, 16), 1, 4, None, 0),
    ((70, 10, 130), 2, 1, POT1, 0),    # more than one x-chunk / z tile
]


@pytest.mark.parametrize("stage", [0, 2])
@pytest.mark.parametrize("N,h,n,pot,variant", CASES)
@pytest.mark.parametrize("strict", [True, False])
def test_fused_stage_fp64(N, h, n, pot, variant, strict, stage):
    rng = np.random.default_rng(7 + h)
    f, d, kf, kd = _state(rng, n, N, h, np.float64)
    dx, dt, a, hub = (0.1, 0.13, 0.07), 0.01, 1.3, 0.2
    A, B = RK54[0][stage], RK54[1][stage]
    fo, do, kfo, kdo = f.copy(), d.copy(), kf.copy(), kd.copy()
    lap = np.zeros((n,) + tuple(N))
    O.reference_schedule_stage(stage, fo, do, kfo, kdo, lap, h, dx, dt, a=a, hubble=hub,
                               potential_terms=pot)
    tf, td, tkf, tkd = (to_dev(x) for x in (f, d, kf, kd))
    tfo = torch.full_like(tf, float("nan"))
    K.scalar_stage(fused=True, f_in=tf, f_out=tfo, dfdt=td, kf=tkf, kdfdt=tkd, h=h, dx=dx,
                   dt=dt, A=A, B=B, a=a, hubble=hub, dV_terms=_dV(pot, n), strict=strict,
                   variant=variant, xchunk=32)
    torch.cuda.synchronize()
    got = {"f": tfo, "dfdt": td, "kf": tkf, "kdfdt": tkd}
    ref = {"f": fo, "dfdt": do, "kf": kfo, "kdfdt": kdo}
    for name in got:
        g = O.interior(got[name].cpu().numpy(), h)
        r = O.interior(ref[name], h)
        if strict:
            assert np.array_equal(g, r), f"{name} not bit-exact"
        else:
            assert scale_err(g, r) < 1e-12, name
    # input f untouched; pointwise arrays' halos untouched
    assert torch.equal(tf, to_dev(f))
    for name, orig in (("dfdt", d), ("kf", kf), ("kdfdt", kd)):
        g = got[name].cpu().numpy()
        mask = np.ones(g.shape, bool)
        O.interior(mask, h)[...] = False
        assert np.array_equal(g[mask], orig[mask]), f"{name} halo clobbered"
    # write-through halos: the six faces of f_out equal a periodic fill
    g = tfo.cpu().numpy()
    full = fo.copy()
    for i in range(n):
        O.share_halos(full[i], h)
    s = slice(h, -h)
    for i in range(n):
        for sl in ((slice(0, h), s, s), (slice(-h, None), s, s), (s, slice(0, h), s),
                   (s, slice(-h, None), s), (s, s, slice(0, h)), (s, s, slice(-h, None))):
            if strict:
                assert np.array_equal(g[i][sl], full[i][sl])
            else:
                assert scale_err(g[i][sl], full[i][sl]) < 1e-12


@pytest.mark.parametrize("N,h,n,pot", [((32, 32, 32), 2, 1, POT1), ((17, 9, 33), 1, 1, None),
                                       ((24, 20, 64), 2, 2, POT2)])
def test_fused_stage_fp32(N, h, n, pot):
    rng = np.random.default_rng(3)
    f, d, kf, kd = _state(rng, n, N, h, np.float32)
    dx, dt, a, hub = (0.1, 0.13, 0.07), 0.01, 1.3, 0.2
    A, B = RK54[0][1], RK54[1][1]
    fo, do, kfo, kdo = (x.astype(np.float64) for x in (f, d, kf, kd))
    lap = np.zeros((n,) + tuple(N))
    O.reference_schedule_stage(1, fo, do, kfo, kdo, lap, h, dx, dt, a=a, hubble=hub,
                               potential_terms=pot)
    tf, td, tkf, tkd = (to_dev(x) for x in (f, d, kf, kd))
    tfo = torch.zeros_like(tf)
    K.scalar_stage(fused=True, f_in=tf, f_out=tfo, dfdt=td, kf=tkf, kdfdt=tkd, h=h, dx=dx,
                   dt=dt, A=A, B=B, a=a, hubble=hub, dV_terms=_dV(pot, n))
    for g, r in ((tfo, fo), (td, do), (tkf, kfo), (tkd, kdo)):
        assert scale_err(O.interior(g.cpu().numpy(), h), O.interior(r, h)) < 1e-5


def test_fused_x_range_split_equals_whole():
    """Boundary/interior launches (multi-rank schedule) == one launch."""
    rng = np.random.default_rng(11)
    N, h = (40, 16, 32), 2
    f, d, kf, kd = _state(rng, 1, N, h, np.float64)
    args = dict(h=h, dx=(0.1,) * 3, dt=0.01, A=RK54[0][1], B=RK54[1][1],
                dV_terms=_dV(POT1, 1), wrap=(False, True, True))
    outs = []
    for ranges in ([(0, 40)], [(0, 2), (38, 40), (2, 38)]):
        tf, td, tkf, tkd = (to_dev(x) for x in (f, d, kf, kd))
        tfo = torch.zeros_like(tf)
        for r in ranges:
            K.scalar_stage(fused=True, f_in=tf, f_out=tfo, dfdt=td, kf=tkf, kdfdt=tkd,
                           x_range=r, **args)
        outs.append([x.cpu() for x in (tfo, td, tkf, tkd)])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("strict", [True, False])
def test_unfused_stage(strict):
    rng = np.random.default_rng(2)
    N, h, n = (16, 12, 20), 2, 2
    f, d, kf, kd = _state(rng, n, N, h, np.float64)
    lap = rng.random((n,) + N)
    A, B = RK54[0][3], RK54[1][3]
    fo, do, kfo, kdo = f.copy(), d.copy(), kf.copy(), kd.copy()
    fi = [O.interior(fo[i], h) for i in range(n)]
    di = [O.interior(do[i], h) for i in range(n)]
    rhs = [np.array(r) for r in O.scalar_rhs(fi, di, list(lap), 1.1, 0.3, POT2)]
    unk, tmp = [], []
    for i in range(n):
        unk += [fi[i], di[i]]
        tmp += [O.interior(kfo[i], h), O.interior(kdo[i], h)]
    O.lsrk_stage(A, B, 0.02, unk, tmp, rhs)
    tf, td, tkf, tkd, tl = (to_dev(x) for x in (f, d, kf, kd, lap))
    K.scalar_stage(fused=False, f_in=tf, f_out=tf, dfdt=td, kf=tkf, kdfdt=tkd, lap_f=tl,
                   h=h, dt=0.02, A=A, B=B, a=1.1, hubble=0.3, dV_terms=_dV(POT2, n),
                   strict=strict)
    for g, r in ((tf, fo), (td, do), (tkf, kfo), (tkd, kdo)):
        assert scale_err(g.cpu().numpy(), r) < 1e-13


@pytest.mark.parametrize("h", [1, 2, 3, 4])
@pytest.mark.parametrize("halo_extra", [0, 1])
@pytest.mark.parametrize("strict", [True, False])
def test_finite_difference_vs_oracle(h, halo_extra, strict):
    rng = np.random.default_rng(h)
    N = (18, 11, 37)
    halo = h + halo_extra
    fx = rng.random(tuple(n + 2 * halo for n in N))
    O.share_halos(fx, halo)
    dx = (0.1, 0.2, 0.3)
    t = to_dev(fx)
    lap = torch.zeros(N, dtype=torch.float64, device="cuda")
    g = [torch.zeros_like(lap) for _ in range(3)]
    K.finite_difference(t, lap=lap, pdx=g[0], pdy=g[1], pdz=g[2], h=h, halo=halo, dx=dx,
                        strict=strict)
    ref_l = O.laplacian(fx, h, dx, halo=halo)
    inner = fx[tuple(slice(halo - h, n - (halo - h)) for n in fx.shape)]
    ref_g = O.gradient(inner, h, dx)
    if strict:
        assert np.array_equal(lap.cpu().numpy(), ref_l)
        for i in range(3):
            assert np.array_equal(g[i].cpu().numpy(), ref_g[i])
    else:
        assert scale_err(lap.cpu().numpy(), ref_l) < 1e-13
        for i in range(3):
            assert scale_err(g[i].cpu().numpy(), ref_g[i]) < 1e-13
    # single outputs and the accumulate (divergence) kernels
    one = torch.zeros_like(lap)
    K.finite_difference(t, lap=one, h=h, halo=halo, dx=dx, strict=strict)
    assert torch.equal(one, lap)
    for name, i in (("pdx", 0), ("pdy", 1), ("pdz", 2)):
        one = torch.zeros_like(lap)
        K.finite_difference(t, h=h, halo=halo, dx=dx, strict=strict, **{name: one})
        assert torch.equal(one, g[i])
    div = g[0].clone()
    K.finite_difference(t, pdy=div, h=h, halo=halo, dx=dx, accumulate=True, strict=strict)
    K.finite_difference(t, pdz=div, h=h, halo=halo, dx=dx, accumulate=True, strict=strict)
    ref_d = O.divergence(np.stack([inner] * 3), h, dx)
    if strict:
        assert np.array_equal(div.cpu().numpy(), ref_d)
    else:
        assert scale_err(div.cpu().numpy(), ref_d) < 1e-13


@pytest.mark.parametrize("h", [1, 2, 3, 4])
def test_plane_wave_eigenvalues_cuda(h):
    """test/test_derivs.py:38-135 against the CUDA kernels."""
    grid_shape = (32, 32, 32)
    L = (3, 5, 7)
    dx = tuple(Li / Ni for Li, Ni in zip(L, grid_shape))
    dk = tuple(2 * np.pi / Li for Li in L)
    kvec = np.array(dk) * np.array([-5, 4, -3]).astype(np.float64)
    xvec = np.meshgrid(*[dxi * np.arange(0, ni) for dxi, ni in zip(dx, grid_shape)],
                       indexing="ij")
    phases = sum(ki * xi for ki, xi in zip(kvec, xvec))
    fx = np.zeros(tuple(n + 2 * h for n in grid_shape))
    fx[h:-h, h:-h, h:-h] = np.sin(phases)
    t = to_dev(fx)
    K.halo_wrap(t, (h, h, h))
    lap = torch.zeros(grid_shape, dtype=torch.float64, device="cuda")
    g = [torch.zeros_like(lap) for _ in range(3)]
    K.finite_difference(t, lap=lap, pdx=g[0], pdy=g[1], pdz=g[2], h=h, dx=dx)
    lap_true = sum(O.second_eigenvalue(h, k, d) for k, d in zip(kvec, dx)) * np.sin(phases)
    mask = np.abs(lap_true) > 1e-11
    max_err, avg_err = O.get_errs(lap_true[mask], lap.cpu().numpy()[mask])
    assert max_err < 1e-9 and avg_err < 1e-11
    for i in range(3):
        true = O.first_eigenvalue(h, kvec[i], dx[i]) * np.cos(phases)
        mask = np.abs(true) > 1e-11
        max_err, avg_err = O.get_errs(true[mask], g[i].cpu().numpy()[mask])
        assert max_err < 1e-9 and avg_err < 1e-11


@pytest.mark.parametrize("h, grid_shape", [
    (1, (32, 32, 32)), (1, (128, 64, 32)), (1, (109, 67, 32)),
    ((2, 0, 3), (32, 32, 32)), ((0, 2, 1), (32, 32, 32)), (2, (16, 12, 20)),
    (4, (9, 8, 11))])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_halo_wrap_bit_exact(h, grid_shape, dtype):
    """test/test_decomp.py:47-96 (single rank): equals a globally periodic array."""
    rng = np.random.default_rng(12321)
    hh = (h,) * 3 if isinstance(h, int) else h
    data = periodic_global(rng, grid_shape, hh, dtype)
    sub = np.zeros_like(data)
    slc = tuple(slice(hi, -hi) if hi > 0 else slice(None) for hi in hh)
    sub[slc] = data[slc]
    t = to_dev(sub)
    K.halo_wrap(t, hh)
    assert np.array_equal(t.cpu().numpy(), data)


def test_pack_unpack_and_padding():
    rng = np.random.default_rng(1)
    N, h = (10, 12, 14), (2, 3, 1)
    arr = rng.random(tuple(n + 2 * x for n, x in zip(N, h)))
    t = to_dev(arr)
    for axis, shape in ((0, (2, h[0], N[1] + 2 * h[1], N[2] + 2 * h[2])),
                        (1, (2, N[0] + 2 * h[0], h[1], N[2] + 2 * h[2]))):
        buf = torch.zeros(shape, dtype=torch.float64, device="cuda")
        K.halo_pack(t, buf, h, axis)
        b = buf.cpu().numpy()
        if axis == 0:
            assert np.array_equal(b[1], arr[N[0]:N[0] + h[0]])
            assert np.array_equal(b[0], arr[h[0]:2 * h[0]])
        else:
            assert np.array_equal(b[1], arr[:, N[1]:N[1] + h[1]])
            assert np.array_equal(b[0], arr[:, h[1]:2 * h[1]])
        t2 = t.clone()
        K.halo_pack(t2, buf, h, axis, unpack=True)
        ref = arr.copy()
        if axis == 0:
            ref[:h[0]] = b[1]; ref[N[0] + h[0]:] = b[0]
        else:
            ref[:, :h[1]] = b[1]; ref[:, N[1] + h[1]:] = b[0]
        assert np.array_equal(t2.cpu().numpy(), ref)
    un = torch.zeros(N, dtype=torch.float64, device="cuda")
    K.remove_halos(t, un, h)
    assert np.array_equal(un.cpu().numpy(), O.interior(arr, h))
    back = torch.zeros_like(t)
    K.remove_halos(back, un, h, restore=True)
    assert np.array_equal(O.interior(back.cpu().numpy(), h), O.interior(arr, h))


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 5e-14), (np.float32, 1e-5)])
def test_reduce_poly(dtype, rtol):
    rng = np.random.default_rng(4)
    N, h = (20, 18, 40), 2
    a = rng.random(tuple(n + 2 * h for n in N)).astype(dtype)
    b = rng.random(N).astype(dtype) - 0.5
    ta, tb = to_dev(a), to_dev(b)
    ai = O.interior(a, h).astype(np.float64)
    bb = b.astype(np.float64)
    terms = [(0, 0.5, [2, 0]), (1, -0.5, [1, 1]), (1, 0.31, [0, 0]), (2, 1.0, [0, 1]),
             (3, 1.0, [0, 1]), (4, 1.0, [0, 1]), (5, 1.0, [0, 1])]
    out = K.reduce_poly([ta, tb], [h, 0], terms, [0, 0, 1, 2, 1, 2], [0, 0, 0, 0, 1, 1], N)
    got = out.cpu().numpy()
    ref = [np.sum(0.5 * ai**2), np.sum(-0.5 * ai * bb + 0.31), bb.max(), bb.min(),
           np.abs(bb).max(), np.abs(bb).min()]
    assert np.allclose(got, ref, rtol=rtol, atol=0)


def test_full_size_properties():
    """
    Size-independent properties at the benchmark size (512^3 fp64, h=2):
    (1) a constant field has zero Laplacian: f, dfdt stay constant under the wave
        equation; (2) linearity of a stage in (f, dfdt, k).
    """
    N, h = (512, 512, 512), 2
    shape = tuple(n + 2 * h for n in N)
    dev = "cuda"
    f = torch.full(shape, 0.75, dtype=torch.float64, device=dev)
    d = torch.zeros(shape, dtype=torch.float64, device=dev)
    kf = torch.zeros_like(f); kd = torch.zeros_like(f); fo = torch.zeros_like(f)
    K.scalar_stage(fused=True, f_in=f, f_out=fo, dfdt=d, kf=kf, kdfdt=kd, h=h,
                   dx=(0.01,) * 3, dt=1e-3, A=RK54[0][1], B=RK54[1][1])
    s_ = slice(h, -h)
    for sl in ((s_, s_, s_), (slice(0, h), s_, s_), (slice(-h, None), s_, s_),
               (s_, slice(0, h), s_), (s_, slice(-h, None), s_), (s_, s_, slice(0, h)),
               (s_, s_, slice(-h, None))):          # interior + the six face halos
        assert float(fo[sl].min()) == 0.75 and float(fo[sl].max()) == 0.75
    assert float(d.abs().max()) < 1e-9   # |dt * lap| round-off only
    del f, d, kf, kd, fo
    torch.manual_seed(0)
    M = (128, 128, 128)
    shp = tuple(n + 2 * h for n in M)

    def run(fv, dv):
        f = fv.clone(); K.halo_wrap(f, (h,) * 3)
        d = dv.clone(); kf = torch.zeros_like(f); kd = torch.zeros_like(f)
        fo = torch.zeros_like(f)
        K.scalar_stage(fused=True, f_in=f, f_out=fo, dfdt=d, kf=kf, kdfdt=kd, h=h,
                       dx=(0.01,) * 3, dt=1e-3, A=0., B=RK54[1][0])
        return fo, d

    f1, d1 = (torch.rand(shp, dtype=torch.float64, device=dev) for _ in range(2))
    f2, d2 = (torch.rand(shp, dtype=torch.float64, device=dev) for _ in range(2))
    o1, e1 = run(f1, d1)
    o2, e2 = run(f2, d2)
    o3, e3 = run(f1 + 2 * f2, d1 + 2 * d2)
    s = (slice(h, -h),) * 3
    assert float((o3 - (o1 + 2 * o2))[s].abs().max()) < 1e-12
    assert float((e3 - (e1 + 2 * e2))[s].abs().max()) < 1e-9
