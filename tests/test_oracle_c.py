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
