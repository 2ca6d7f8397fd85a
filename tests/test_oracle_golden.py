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
