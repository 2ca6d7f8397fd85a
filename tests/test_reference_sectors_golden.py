"""
Oracle and host-side sector algebra against ``tests/golden/reference_sectors.json``: the
right-hand sides, reducers, dV/df, stress tensor and rho / P produced by the reference's OWN
``pystella/sectors.py`` + ``pystella/field`` (imported unmodified by
``tests/golden/make_golden_sectors.py`` on top of a stand-in for pymbolic's expression nodes)
on seeded random data.  The test regenerates the data from the seeds.
"""

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
import pystella_b200 as ps  # noqa: E402
from pystella_b200.field import evaluate, is_constant  # noqa: E402

RTOL = 2e-13      # different association orders of the same fp64 expression


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_sectors.json")) as fh:
        return json.load(fh)


def _data(golden, ns, seed):
    """Must mirror make_golden_sectors.random_data (same generator, same draw order)."""
    lattice = tuple(golden["lattice"])
    rng = np.random.default_rng(seed)
    data = {
        "f": rng.standard_normal((ns,) + lattice),
        "dfdt": rng.standard_normal((ns,) + lattice),
        "lap_f": rng.standard_normal((ns,) + lattice),
        "dfdx": rng.standard_normal((ns, 3) + lattice),
        "hij": rng.standard_normal((6,) + lattice),
        "dhijdt": rng.standard_normal((6,) + lattice),
        "lap_hij": rng.standard_normal((6,) + lattice),
        "a": float(1.0 + rng.random()),
        "hubble": float(0.1 + rng.random()),
    }
    return data


def _terms(case):
    return [(c, tuple(p)) for c, p in case["potential_terms"]]


def _potential(terms):
    def V(f):
        out = 0
        for c, powers in terms:
            t = c
            for n, p in enumerate(powers):
                if p:
                    t = t * f[n]**p
            out = out + t
        return out
    return V


def _close(got, want, what):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(1.0, float(np.max(np.abs(want))))
    err = float(np.max(np.abs(got - want))) / scale
    assert err < RTOL, (what, err)


CASES = ["massive_1field", "mixed_3field", "preheating_2field", "test_energy_2field"]


def test_fixture_has_the_expected_cases(golden):
    assert sorted(golden["cases"]) == CASES
    for name, case in golden["cases"].items():
        ns = case["nscalars"]
        # rhs_dict order of the reference: f[0], dfdt[0], f[1], dfdt[1], ...
        assert case["rhs_keys"] == [k for n in range(ns) for k in (f"f[{n}]", f"dfdt[{n}]")]
        d = _data(golden, ns, case["seed"])
        assert d["a"] == case["a"] and d["hubble"] == case["hubble"], name


@pytest.mark.parametrize("name", CASES)
def test_oracle_scalar_rhs_matches_reference_sector(golden, name):
    """oracle.scalar_rhs / poly_diff against ScalarSector.rhs_dict (sectors.py:117-131)."""
    case = golden["cases"][name]
    ns, terms = case["nscalars"], _terms(case)
    d = _data(golden, ns, case["seed"])
    rhs = O.scalar_rhs(list(d["f"]), list(d["dfdt"]), list(d["lap_f"]), d["a"], d["hubble"],
                       terms)
    assert len(rhs) == len(case["rhs"])
    for k, got, want in zip(case["rhs_keys"], rhs, case["rhs"]):
        _close(got, want, (name, k))
    for n in range(ns):
        _close(O.poly_eval(O.poly_diff(terms, n), list(d["f"])) + 0 * d["f"][0],
               case["dV"][n], (name, "dV", n))


@pytest.mark.parametrize("name", CASES)
def test_oracle_energies_match_reference_reducers(golden, name):
    """oracle.scalar_energy / get_rho_and_p against ScalarSector.reducers averaged over the
    lattice and sectors.get_rho_and_p (sectors.py:133-144, 211-229)."""
    case = golden["cases"][name]
    ns, terms = case["nscalars"], _terms(case)
    d = _data(golden, ns, case["seed"])
    e = O.scalar_energy(list(d["f"]), list(d["dfdt"]), list(d["lap_f"]), d["a"], terms)
    for key in ("kinetic", "potential", "gradient"):
        want = [np.mean(np.array(x)) for x in case["reducers"][key]]
        _close(e[key], want, (name, key))
    full = O.get_rho_and_p(dict(e))
    _close(full["total"], case["rho_and_p"]["total"], (name, "total"))
    _close(full["pressure"], case["rho_and_p"]["pressure"], (name, "pressure"))


@pytest.mark.parametrize("name", ["massive_1field", "preheating_2field", "test_energy_2field"])
def test_oracle_gw_rhs_matches_reference_tensor_sector(golden, name):
    """oracle.gw_rhs against TensorPerturbationSector.rhs_dict sourced by
    ScalarSector.stress_tensor(i, j, drop_trace=True) (sectors.py:146-152, 187-204)."""
    case = golden["cases"][name]
    ns = case["nscalars"]
    d = _data(golden, ns, case["seed"])
    rhs = O.gw_rhs(d["dhijdt"], d["lap_hij"], d["dfdx"], d["hubble"])
    keys = []
    for i in range(1, 4):
        for j in range(i, 4):
            c = O.tensor_index(i, j)
            keys += [f"hij[{c}]", f"dhijdt[{c}]"]
            Sij = sum(d["dfdx"][n][i - 1] * d["dfdx"][n][j - 1] for n in range(ns))
            _close(Sij, case["stress_tensor_drop_trace"][f"{i}{j}"], (name, "Sij", i, j))
    assert case["gw_rhs_keys"] == keys
    for k, got, want in zip(keys, rhs, case["gw_rhs"]):
        _close(got, want, (name, k))
    for ij, c in golden["tensor_index"].items():
        assert O.tensor_index(int(ij[0]), int(ij[1])) == c


@pytest.mark.parametrize("name", CASES)
def test_product_sector_algebra_matches_reference(golden, name):
    """The host side of the product: ``ps.ScalarSector`` / ``ps.TensorPerturbationSector``
    build the same equations (evaluated with the package's own evaluator), and the
    (coefficient, powers) tables that ``scalar_system_from_sector`` hands to the C ABI,
    evaluated the way the kernels evaluate them, reproduce the reference's right-hand sides,
    potential and energies."""
    from pystella_b200.step import recognise_scalar_system, scalar_system_from_sector
    case = golden["cases"][name]
    ns, terms = case["nscalars"], _terms(case)
    d = _data(golden, ns, case["seed"])
    sector = ps.ScalarSector(ns, potential=_potential(terms))
    rhs = sector.rhs_dict
    assert len(rhs) == 2 * ns
    for (key, expr), want in zip(rhs.items(), case["rhs"]):
        _close(evaluate(expr, d) + 0 * d["f"][0], want, (name, str(key)))
    red = sector.reducers
    for kind in ("kinetic", "potential", "gradient"):
        for expr, want in zip(red[kind], case["reducers"][kind]):
            _close(evaluate(expr, d) + 0 * d["f"][0], want, (name, kind))
    if "gw_rhs" in case:
        gw = ps.TensorPerturbationSector([sector])
        for (key, expr), want in zip(gw.rhs_dict.items(), case["gw_rhs"]):
            _close(evaluate(expr, d), want, (name, "gw", str(key)))

    a, hubble = d["a"], d["hubble"]
    fields = list(d["f"])
    for system, a_factor in ((scalar_system_from_sector(sector), a**2),
                             (recognise_scalar_system(sector.rhs_dict), 1.0)):
        hub = float(evaluate(system.hubble, d))
        for n in range(ns):
            dV = [(float(c) if is_constant(c) else float(evaluate(c, d)), p)
                  for c, p in system.dV[n]]
            got = (d["lap_f"][n] - 2 * hub * d["dfdt"][n]) - a_factor * O.poly_eval(dV, fields)
            _close(got, case["rhs"][2 * n + 1], (name, "tables", n))
        if system.V is not None:
            V = [(float(c) if is_constant(c) else float(evaluate(c, d)), p)
                 for c, p in system.V]
            _close(O.poly_eval(V, fields) + 0 * fields[0], case["reducers"]["potential"][0],
                   (name, "V tables"))


def test_product_generated_rhs_for_a_cosine_potential(golden):
    """A potential outside the polynomial menu (the generated-kernel route): the package's
    symbolic derivative against the reference's ``diff`` (field/diff.py:29-94)."""
    case = golden["cosine_case"]
    d = _data(golden, 1, case["seed"])
    cos = ps.var("cos") if hasattr(ps, "var") else None
    if cos is None:
        from pystella_b200.field import Variable
        cos = Variable("cos")
    sector = ps.ScalarSector(
        1, potential=lambda f: case["Lambda4"] * (1 - cos(f[0] / case["f_a"])))
    rhs = sector.rhs_dict
    assert case["rhs_keys"] == ["f[0]", "dfdt[0]"]
    for (key, expr), want in zip(rhs.items(), case["rhs"]):
        _close(evaluate(expr, d) + 0 * d["f"][0], want, str(key))
    # and against the closed form
    f0 = d["f"][0]
    want = d["lap_f"][0] - 2 * d["hubble"] * d["dfdt"][0] \
        - d["a"]**2 * case["Lambda4"] / case["f_a"] * np.sin(f0 / case["f_a"])
    _close(case["rhs"][1], want, "closed form")
