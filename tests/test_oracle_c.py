"""The C/OpenMP restatement (CPU baseline) agrees with the NumPy oracle."""
import numpy as np

import oracle as O
from oracle import c_oracle as CO


def test_c_oracle_matches_numpy_oracle():
    rng = np.random.default_rng(0)
    N, h = (16, 12, 20), 2
    pad = tuple(n + 2 * h for n in N)
    terms = [(0.5, (2, 0)), (0.5, (0, 2)), (0.5, (2, 2))]
    f = rng.random((2,) + pad)
    d = rng.random((2,) + pad)
    dx, dt = (0.1, 0.2, 0.3), 0.01
    pot = CO.make_potential(2, terms)
    gs = float(np.prod(N))

    def energy(fa, da, a):
        fi, di, laps = [], [], []
        for i in range(2):
            g = fa[i].copy()
            O.share_halos(g, h)
            laps.append(O.laplacian(g, h, dx))
            fi.append(O.interior(fa[i], h))
            di.append(O.interior(da[i], h))
        return O.get_rho_and_p(O.scalar_energy(fi, di, laps, a, terms))

    fr, dr = f.copy(), d.copy()
    kfr, kdr = np.zeros_like(f), np.zeros_like(f)
    lapr = np.zeros((2,) + N)
    e = energy(fr, dr, 1.0)
    ex = O.Expansion(e["total"], "LowStorageRK54", mpl=1.0)
    A, B, _ = O.LSRK_TABLES["LowStorageRK54"]

    fc, dc = f.copy(), d.copy()
    kfc, kdc = np.zeros_like(f), np.zeros_like(f)
    lapc = np.zeros((2,) + N)
    for i in range(2):
        CO.share_halos(fc[i], h)
        CO.laplacian(fc[i], lapc[i], h, dx)
        ref = f[i].copy()
        O.share_halos(ref, h)
        assert np.array_equal(fc[i], ref)
        assert np.allclose(lapc[i], O.laplacian(ref, h, dx), rtol=1e-13, atol=1e-11)
    en_c = CO.scalar_energy(fc, dc, lapc, h, 1.0, gs, pot)
    assert np.isclose(en_c.sum(), e["total"], rtol=1e-13)
    adot = np.sqrt(8 * np.pi / 3 * en_c.sum())
    state = np.array([1.0, adot, adot, 0., 0.])
    for _ in range(2):
        for s in range(5):
            O.reference_schedule_stage(s, fr, dr, kfr, kdr, lapr, h, dx, dt, a=ex.a[0],
                                       hubble=ex.hubble[0], potential_terms=terms)
            ex.step(s, e["total"], e["pressure"], dt)
            e = energy(fr, dr, ex.a[0])
    CO.preheating_steps(2, A, B, fc, dc, kfc, kdc, lapc, h, dx, dt, 1.0, gs, pot, state, en_c)
    assert np.max(np.abs(O.interior(fc, h) - O.interior(fr, h))) < 1e-13
    assert np.max(np.abs(O.interior(dc, h) - O.interior(dr, h))) < 1e-13
    assert np.isclose(state[0], ex.a[0], rtol=1e-14)
    assert np.isclose(state[1], ex.adot[0], rtol=1e-13)
    assert CO.num_threads() >= 1


def test_c_oracle_tensor_sector_matches_numpy_oracle():
    """Config-4 restatement (scalars + tensor perturbations): C vs NumPy oracle."""
    rng = np.random.default_rng(1)
    N, h = (12, 10, 16), 2
    pad = tuple(n + 2 * h for n in N)
    terms = [(0.5, (2, 0)), (0.15, (2, 2))]
    dx, dt = (0.1, 0.2, 0.3), 0.01
    a, hubble = 1.0, 0.25
    f, d = 0.2 * rng.standard_normal((2,) + pad), 0.2 * rng.standard_normal((2,) + pad)
    hh, dh = 1e-2 * rng.standard_normal((6,) + pad), 1e-2 * rng.standard_normal((6,) + pad)
    pot = CO.make_potential(2, terms)
    A, B, _ = O.LSRK_TABLES["LowStorageRK54"]

    # NumPy oracle: the loop of tests/test_gpu_api.py::test_two_scalars_plus_gw_sector
    F, D, H, DH = f.copy(), d.copy(), hh.copy(), dh.copy()
    kF, kD, kH, kDH = (np.zeros_like(x) for x in (F, D, H, DH))
    for s in range(5):
        lapf, grd, laph = [], [], []
        for n in range(2):
            O.share_halos(F[n], h)
            lapf.append(O.laplacian(F[n], h, dx))
            grd.append(O.gradient(F[n], h, dx))
        for c in range(6):
            O.share_halos(H[c], h)
            laph.append(O.laplacian(H[c], h, dx))
        Fi = [O.interior(F[n], h) for n in range(2)]
        Di = [O.interior(D[n], h) for n in range(2)]
        Hi = [O.interior(H[c], h) for c in range(6)]
        DHi = [O.interior(DH[c], h) for c in range(6)]
        rhs = [np.array(r) for r in O.scalar_rhs(Fi, Di, lapf, a, hubble, terms)]
        rhs += [np.array(r) for r in O.gw_rhs(DHi, laph, np.stack(grd), hubble)]
        unk, tmp = [], []
        for n in range(2):
            unk += [Fi[n], Di[n]]
            tmp += [O.interior(kF[n], h), O.interior(kD[n], h)]
        for c in range(6):
            unk += [Hi[c], DHi[c]]
            tmp += [O.interior(kH[c], h), O.interior(kDH[c], h)]
        O.lsrk_stage(A[s], B[s], dt, unk, tmp, rhs)

    fc, dc, hc, dhc = f.copy(), d.copy(), hh.copy(), dh.copy()
    kf, kd, kh, kdh = (np.zeros_like(x) for x in (fc, dc, hc, dhc))
    lap_f, dfdx, lap_hij = CO.gw_initial_derivatives(fc, hc, h, dx)
    g0 = f[0].copy()
    O.share_halos(g0, h)
    assert np.allclose(dfdx[0], O.gradient(g0, h, dx), rtol=1e-13, atol=1e-12)
    state = np.array([a, hubble * a, hubble, 0., 0.])
    en = np.zeros(5)
    CO.preheating_gw_steps(1, A, B, fc, dc, kf, kd, lap_f, dfdx, hc, dhc, kh, kdh, lap_hij, h,
                           dx, dt, 1.0, float(np.prod(N)), pot, state, en,
                           with_expansion=False)
    for got, ref in ((fc, F), (dc, D), (hc, H), (dhc, DH)):
        assert np.max(np.abs(O.interior(got, h) - O.interior(ref, h))) < 1e-13 * max(
            1.0, np.max(np.abs(ref)))
