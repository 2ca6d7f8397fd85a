"""
Pin the oracle AND the product's tables against numbers extracted from the
reference sources (tests/golden/reference_tables.json, made by
tests/golden/make_golden.py which imports the unmodified reference modules).
"""
import numpy as np
import pytest

import oracle as O
import pystella_b200 as ps
from pystella_b200 import derivs as pd, step as pstep


def test_fd_coefficient_tables(golden):
    for h in (1, 2, 3, 4):
        g = {int(s): c for s, c in golden["grad_coefs"][str(h)].items()}
        l = {int(s): c for s, c in golden["lap_coefs"][str(h)].items()}
        assert O.GRAD_COEFS[h] == g and pd._GRAD[h] == g
        assert O.LAP_COEFS[h] == l and pd._LAP[h] == l


def test_stencil_eigenvalues(golden):
    e = golden["eigenvalues"]
    L, grid, modes = e["L"], e["grid"], e["modes"]
    dx = [Li / Ni for Li, Ni in zip(L, grid)]
    k = [2 * np.pi / Li * m for Li, m in zip(L, modes)]
    for h in (1, 2, 3, 4):
        for ax in range(3):
            ref1 = e["values"][str(h)]["first"][ax]
            ref2 = e["values"][str(h)]["second"][ax]
            assert O.first_eigenvalue(h, k[ax], dx[ax]) == pytest.approx(ref1, rel=1e-15)
            assert O.second_eigenvalue(h, k[ax], dx[ax]) == pytest.approx(ref2, rel=1e-15)
            assert ps.FirstCenteredDifference(h).get_eigenvalues(k[ax], dx[ax]) == \
                pytest.approx(ref1, rel=1e-13)
            assert ps.SecondCenteredDifference(h).get_eigenvalues(k[ax], dx[ax]) == \
                pytest.approx(ref2, rel=1e-13)


def test_lsrk_tables(golden):
    for name, tab in golden["lsrk"].items():
        cls = getattr(ps, name)
        assert [float(x) for x in cls._A] == tab["A"], name
        assert [float(x) for x in cls._B] == tab["B"], name
        assert [float(x) for x in cls._C] == tab["C"], name
        assert cls.num_stages == tab["num_stages"]
        assert cls.expected_order == tab["expected_order"]
        if name in O.LSRK_TABLES:
            A, B, order = O.LSRK_TABLES[name]
            assert [float(x) for x in A] == tab["A"] and [float(x) for x in B] == tab["B"]
            assert order == tab["expected_order"]


def test_rk_tables(golden):
    for name, tab in golden["rk"].items():
        cls = getattr(ps, name)
        assert (cls.num_stages, cls.num_copies, cls.expected_order) == \
            (tab["num_stages"], tab["num_copies"], tab["expected_order"])
        assert O.RK_TABLES[name] == (tab["num_stages"], tab["num_copies"],
                                     tab["expected_order"])
    assert [c.__name__ for c in ps.all_steppers] == golden["all_steppers"]
    assert [c.__name__ for c in pstep.all_steppers] == golden["all_steppers"]


def test_tensor_index(golden):
    for key, val in golden["tensor_index"].items():
        i, j = (int(x) for x in key.split(","))
        assert O.tensor_index(i, j) == val
        assert ps.tensor_index(i, j) == val


def test_partition(golden):
    for case in golden["partition"]:
        P, G, rt = tuple(case["proc_shape"]), tuple(case["grid_shape"]), \
            tuple(case["rank_tuple"])
        shape, start = O.rank_shape_start(P, G, rt)
        assert list(shape) == case["rank_shape"] and list(start) == case["rank_start"]
        assert O.rank_id(P, *rt) == case["rankID"]
        assert O.rank_id(P, rt[0] - 1, rt[1] - 1, 0) == case["rankID_minus"]
        assert O.rank_tuple_of(P, case["rankID"]) == rt
        # product arithmetic without a process group
        d = object.__new__(ps.DomainDecomposition)
        d.proc_shape = P
        d.rx, d.ry, d.rz = rt
        s2, st2 = d.get_rank_shape_start(G)
        assert list(s2) == case["rank_shape"] and list(st2) == case["rank_start"]
        assert d.rankID(*rt) == case["rankID"]
        assert d.rankID(rt[0] - 1, rt[1] - 1, 0) == case["rankID_minus"]


def _cplx(pair):
    return np.array(pair[0]) + 1j * np.array(pair[1])


def test_fourier_helpers(golden):
    """The oracle's restatement of the NumPy helpers behind RayleighGenerator ==
    the reference functions themselves (run by make_golden.py)."""
    from oracle import fourier_oracle as F
    for n, ref in golden["fftfreq"].items():
        assert F.fftfreq(int(n)).tolist() == ref
    g = golden["fourier"]
    grid = tuple(g["grid_shape"])
    fk = _cplx(g["fk"])
    herm = F.make_hermitian(fk.copy())
    assert np.array_equal(herm, _cplx(g["make_hermitian"]))
    assert np.array_equal(F.zero_corner_modes(herm.copy(), grid, only_imag=True),
                          _cplx(g["zero_corner_modes_only_imag"]))
    assert np.array_equal(F.zero_corner_modes(fk.copy(), grid), _cplx(g["zero_corner_modes"]))
    wk = np.array(g["ps_wrapper"]["wk"])
    km = wk.copy()
    km[0, 0, 0] = 0.
    power = F.ps_wrapper(lambda w: 1 / 2 / w, wk, km)
    assert np.array_equal(power, np.array(g["ps_wrapper"]["power"]))
    assert np.array_equal(wk, np.array(g["ps_wrapper"]["wk_after"]))
    # a Hermitian half-spectrum transforms to a real field and back
    fx = F.idft(F.zero_corner_modes(herm.copy(), grid, only_imag=True), grid)
    assert np.allclose(F.dft(fx) / np.prod(grid), herm, atol=1e-13)


def test_example_constraint_golden(golden):
    """
    THE stored golden number of the reference (test/test_examples.py:33): the Friedmann
    constraint after `scalar_preheating.py --grid-shape 32 32 32 --end-time 1`.  The whole
    pipeline on the oracle -- WKB Rayleigh initial conditions, 4th-order stencils,
    LowStorageRK54 stages of the two-field sector, energies, FLRW expansion -- lands on it.
    The reference's rtol 1e-3 presumes its own random stream (pyopencl Threefry); with
    NumPy's the value scatters by ~0.5% from seed to seed around the stored number
    (5.549e-8 .. 5.610e-8 for six seeds), hence 2e-2 here.
    """
    from helpers import example_initial_conditions, example_parameters, example_run_oracle
    p = example_parameters(golden)
    assert golden["example"]["argv"] == ["--grid-shape", "32", "32", "32", "--end-time", "1"]
    ic = example_initial_conditions(p, (32, 32, 32), seed=49279)
    nsteps, a, constraint, _ = example_run_oracle(p, ic, end_time=1.)
    assert nsteps == 64
    assert abs(constraint / golden["example"]["constraint"] - 1) < 2e-2, constraint
    # and it is the fluctuations' spectrum that sets it: a homogeneous run misses it
    ic0 = example_initial_conditions(p, (32, 32, 32), seed=49279)
    for i in range(2):
        ic0["f"][i], ic0["dfdt"][i] = p["f0"][i], p["df0"][i]
    _, _, c0, _ = example_run_oracle(p, ic0, end_time=1.)
    assert abs(c0 / golden["example"]["constraint"] - 1) > 0.2
