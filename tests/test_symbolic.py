"""
The symbolic front-end: indexing, shifting, differentiation and argument
inference follow test/test_field.py:30-242 of the reference (compared through the
printed form / numeric evaluation, since expression trees are our own).
"""
import numpy as np
import pytest

import pystella_b200 as ps
from pystella_b200.field import (evaluate, index_fields, parse, shift_fields,
                                 to_polynomial, var)


def same(a, b):
    return str(a).replace(" ", "") == str(b).replace(" ", "")


# (child, constructor kwargs, subscript, prepend_with, shift) -> printed indexed form;
# the cases of test/test_field.py:30-82
INDEXING_CASES = [
    ("y", dict(offset="h"), None, None, None, "y[i+h,j+h,k+h]"),
    ("y", dict(offset="h", indices=("a", "b", "c")), None, None, None, "y[a+h,b+h,c+h]"),
    ("y", dict(ignore_prepends=True), None, (0, 1), None, "y[i,j,k]"),
    ("y[4, 5]", dict(ignore_prepends=True), None, (0, 1), None, "y[4,5,i,j,k]"),
    ("y", dict(ignore_prepends=True), (2, 3), (0, 1), None, "y[2,3,i,j,k]"),
    ("y[4, 5]", dict(ignore_prepends=True), (2, 3), (0, 1), None, "y[2,3,4,5,i,j,k]"),
    ("y", dict(ignore_prepends=False), None, (0, 1), None, "y[0,1,i,j,k]"),
    ("y[4, 5]", dict(ignore_prepends=False), None, (0, 1), None, "y[0,1,4,5,i,j,k]"),
    ("y", dict(ignore_prepends=False), (2, 3), (0, 1), None, "y[0,1,2,3,i,j,k]"),
    ("y[4, 5]", dict(ignore_prepends=False), (2, 3), (0, 1), None,
     "y[0,1,2,3,4,5,i,j,k]"),
    ("y", dict(offset=("hx", "hy", "hz")), None, None, (1, 2, 3),
     "y[i+hx+1,j+hy+2,k+hz+3]"),
    ("y", dict(offset=("hx", var("hy"), "hz")), None, None, (1, 2, var("a")),
     "y[i+hx+1,j+hy+2,k+hz+a]"),
]


@pytest.mark.parametrize("child,kw,sub,prepend,shift,expected", INDEXING_CASES)
def test_field_indexing(child, kw, sub, prepend, shift, expected):
    expr = ps.Field(child, **kw)
    if sub is not None:
        expr = expr[sub]
    if shift is not None:
        expr = shift_fields(expr, shift)
    assert same(index_fields(expr, prepend_with=prepend), expected)


def test_field_offset_length_is_checked():
    with pytest.raises(ValueError):
        ps.Field("y", offset=(1, 2))
    assert parse("f[0]") == var("f")[0]


def test_dynamic_field():
    y = ps.DynamicField("y", offset="h")
    assert same(index_fields(y), "y[i+h,j+h,k+h]")
    assert same(index_fields(y.lap), "lap_y[i,j,k]")
    assert same(index_fields(y.dot), "dydt[i+h,j+h,k+h]")
    assert same(index_fields(y.pd[var("x")]), "dydx[x,i,j,k]")
    assert same(index_fields(y.d(1, 0)), "dydt[1,i+h,j+h,k+h]")
    assert same(index_fields(y.d(1, 1)), "dydx[1,0,i,j,k]")
    f = ps.DynamicField("f")
    assert same(f.d(0), "dfdt") and same(f.d(1, 2, 0), "dfdt[1,2]")
    assert same(f.d(2), "dfdx[1]") and same(f.d(0, 1, 3), "dfdx[0,1,2]")


def test_field_diff():
    from pystella_b200 import diff
    y = ps.Field("y")
    assert diff(y, y) == 1 and diff(y[0], y[0]) == 1 and diff(y[0], y[1]) == 0
    y = ps.DynamicField("y")
    assert diff(y, y) == 1 and diff(y[0], y[0]) == 1 and diff(y[0], y[1]) == 0
    ctx = dict(y=1.7, dydt=0.3)
    assert np.isclose(evaluate(diff(y**3, y, "t"), ctx), 3 * 2 * 1.7 * 0.3)
    assert np.isclose(evaluate(diff(y**3, "t", y), ctx), 3 * 2 * 1.7 * 0.3)
    for i, x in enumerate(["t", "x", "y", "z"]):
        assert diff(y, x) == y.d(i)
        assert diff(y[1, 3], x) == y.d(1, 3, i)
    # chain / product / quotient rules, numerically
    a, b = var("a"), var("b")
    expr = (a**2 * b + 3) / (1 + b**2)
    ctx = dict(a=0.7, b=1.3)
    eps = 1e-6
    num = (evaluate(expr, dict(a=0.7, b=1.3 + eps))
           - evaluate(expr, dict(a=0.7, b=1.3 - eps))) / (2 * eps)
    assert np.isclose(evaluate(diff(expr, b), ctx), num, rtol=1e-8)


def test_get_field_args():
    f = ps.DynamicField("f", offset="h", shape=(2,))
    H = ps.Field("hubble", indices=[])
    args = ps.get_field_args({f[0]: f.dot[0], f.dot[0]: f.lap[0] - 2 * H * f.dot[0]})
    names = [a.name for a in args]
    assert names == sorted(names) == ["dfdt", "f", "hubble", "lap_f"]
    by = {a.name: a for a in args}
    assert by["hubble"].is_value and not by["f"].is_value
    assert len(by["f"].shape) == 4 and by["f"].shape[0] == 2
    assert by["lap_f"].shape == (2, "Nx", "Ny", "Nz")


def test_polynomial_normal_form_and_recognition():
    f = ps.DynamicField("f", offset="h", shape=(2,))
    a = ps.Field("a", indices=[])
    poly = to_polynomial(a**2 * (f[0]**2 * f[1] + 3 * f[1]) / 2 - f.lap[0])
    keys = {tuple((r.name, r.subscript, p) for r, p in m) for m in poly}
    assert (("f", (0,), 2), ("f", (1,), 1)) in keys
    assert (("f", (1,), 1),) in keys and (("lap_f", (0,), 1),) in keys
    with pytest.raises(NotImplementedError):
        to_polynomial(1 / f[0])
    with pytest.raises(NotImplementedError):
        to_polynomial(f[0]**0.5)

    from pystella_b200.step import recognise_scalar_system, scalar_system_from_sector
    sec = ps.ScalarSector(2, potential=lambda g: g[0]**2 / 2 + 0.25 * g[0]**2 * g[1]**2)
    sysm = recognise_scalar_system(sec.rhs_dict)
    assert sysm.nscalars == 2 and sysm.f_name == "f" and sysm.lap_name == "lap_f"
    exact = scalar_system_from_sector(sec)
    ctx = dict(a=1.3, hubble=0.2)
    assert np.isclose(evaluate(sysm.hubble, ctx), 0.2)
    for n in range(2):
        got = sorted((p, float(evaluate(c, ctx))) for c, p in sysm.dV[n])
        ref = sorted((p, 1.3**2 * float(evaluate(c, ctx))) for c, p in exact.dV[n])
        assert [g[0] for g in got] == [r[0] for r in ref]
        assert np.allclose([g[1] for g in got], [r[1] for r in ref])
    # things outside the kernel menu are refused with a reason
    bad = dict(sec.rhs_dict)
    bad[f.dot[0]] = bad[f.dot[0]] + f.dot[1]
    with pytest.raises(NotImplementedError):
        recognise_scalar_system(bad)
    g = ps.DynamicField("g", offset="h")
    with pytest.raises(NotImplementedError):
        recognise_scalar_system({g: g.dot, g.dot: g.lap * 2})


def test_sector_dictionaries():
    """Key order and contents of rhs_dict / reducers (sectors.py:117-144, 187-204)."""
    sec = ps.ScalarSector(2, potential=lambda f: f[0]**2 / 2)
    keys = [str(k) for k in sec.rhs_dict]
    assert keys == ["f[0]", "dfdt[0]", "f[1]", "dfdt[1]"]
    assert list(sec.reducers) == ["kinetic", "potential", "gradient"]
    assert [len(v) for v in sec.reducers.values()] == [2, 1, 2]
    gw = ps.TensorPerturbationSector([sec])
    keys = [str(k) for k in gw.rhs_dict]
    assert keys == [x for c in range(6) for x in (f"hij[{c}]", f"dhijdt[{c}]")]
    assert gw.reducers == {}
    # stress tensor: T_ij (trace dropped) = sum_n d_i f d_j f
    T12 = sec.stress_tensor(1, 2, drop_trace=True)
    ctx = {"dfdx": np.arange(6.).reshape(2, 3) + 1}
    assert np.isclose(evaluate(T12, ctx), 1 * 2 + 4 * 5)
    ctx.update(dfdt=np.array([0.3, 0.4]), f=np.array([0.5, 0.6]), a=1.2)
    T00 = evaluate(sec.stress_tensor(0, 0), ctx)
    kin = 0.3**2 + 0.4**2
    grad2 = sum((ctx["dfdx"][n]**2).sum() for n in range(2))
    lag = -(-kin / 1.2**2 + grad2 / 1.2**2) / 2 - 0.5**2 / 2
    assert np.isclose(T00, kin + (-1.2**2) * lag)


# ---------------------------------------------------------------------------------------
# property test: what the stage kernels are told (coefficients, powers) == the symbolic RHS
# ---------------------------------------------------------------------------------------

from hypothesis import given, settings, strategies as st  # noqa: E402
from pystella_b200.field import is_constant  # noqa: E402


@st.composite
def _potentials(draw):
    n = draw(st.integers(1, 3))
    nterms = draw(st.integers(1, 5))
    terms = []
    for _ in range(nterms):
        powers = tuple(draw(st.integers(0, 4)) for _ in range(n))
        if sum(powers) == 0:
            powers = (2,) + powers[1:]
        coef = draw(st.floats(-3, 3).filter(lambda c: abs(c) > 1e-3))
        terms.append((coef, powers))
    return n, terms


@settings(max_examples=40, deadline=None)
@given(_potentials(), st.integers(0, 2**31 - 1))
def test_recognised_system_equals_symbolic_rhs(pot, seed):
    """For random polynomial potentials: the (coefficient, powers) tables handed to the
    C ABI, evaluated the way the kernels do (oracle.poly_eval), reproduce the sector's
    symbolic right-hand sides and potential at random field values -- for both front ends
    (the sector shortcut and the generic rhs_dict recogniser)."""
    import oracle as O
    from pystella_b200.step import recognise_scalar_system, scalar_system_from_sector
    n, terms = pot

    def potential(f):
        total = 0
        for c, pw in terms:
            mono = c
            for m, p in enumerate(pw):
                if p:
                    mono = mono * f[m]**p
            total = total + mono
        return total

    sector = ps.ScalarSector(n, potential=potential)
    rng = np.random.default_rng(seed)
    fv, dv, lv = rng.random(n) + 0.5, rng.random(n), rng.random(n)
    a, hubble = 1.3, 0.2
    ctx = {"f": fv, "dfdt": dv, "lap_f": lv, "a": a, "hubble": hubble}
    f = ps.DynamicField("f", offset="h", shape=(n,))
    want_rhs = [float(evaluate(sector.rhs_dict[f.dot[i]], ctx)) for i in range(n)]
    want_V = float(sum(c * np.prod(fv**np.array(pw)) for c, pw in terms))

    for system, a_factor in ((scalar_system_from_sector(sector), a**2),
                             (recognise_scalar_system(sector.rhs_dict), 1.0)):
        assert system.nscalars == n
        hub = float(evaluate(system.hubble, ctx))
        assert np.isclose(hub, hubble, rtol=1e-14)
        for i in range(n):
            dV = [(float(c) if is_constant(c) else float(evaluate(c, ctx)), p)
                  for c, p in system.dV[i]]
            got = (lv[i] - 2 * hub * dv[i]) - a_factor * O.poly_eval(dV, list(fv))
            assert np.isclose(got, want_rhs[i], rtol=1e-11, atol=1e-12), (i, terms)
        if system.V is not None:
            V = [(float(c) if is_constant(c) else float(evaluate(c, ctx)), p)
                 for c, p in system.V]
            assert np.isclose(O.poly_eval(V, list(fv)), want_V, rtol=1e-12, atol=1e-13)


def test_collect_field_indices():
    """test/test_field.py:177-199."""
    from pystella_b200.field import collect_field_indices, infer_field_domains
    x = ps.Field("x", offset=(1, 2, 3))
    y = ps.Field("y", indices=("i", "x"), offset="h")
    z = ps.DynamicField("z", shape=(2, "a"))
    assert collect_field_indices({x: y, y: x * z + z.pd[0]}) == {"i", "j", "k", "x"}
    assert collect_field_indices([x, z]) == {"i", "j", "k"}
    assert collect_field_indices([shift_fields(x, (1, 2, 3)), y + z.pd[0], y * z**2]) \
        == {"i", "j", "k", "x"}
    assert infer_field_domains([x, z]) == \
        "{[i,j,k]: 0 <= i < Ni and 0 <= j < Nj and 0 <= k < Nk}"
