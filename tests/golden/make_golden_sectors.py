"""
Generates ``tests/golden/reference_sectors.json``: the right-hand sides, reducers and
``get_rho_and_p`` of the reference's OWN sector classes, evaluated on seeded random data.

``/root/reference/pystella/sectors.py`` (``ScalarSector.rhs_dict``, ``.reducers``,
``.stress_tensor``, ``TensorPerturbationSector.rhs_dict``, ``get_rho_and_p``) and
``pystella/field/__init__.py`` / ``field/diff.py`` (``Field``, ``DynamicField``, ``diff``) are
imported UNMODIFIED.  They are pure Python on top of ``pymbolic``, which is not installed in
this image (nor are loopy / pyopencl); so, exactly like ``make_golden.py`` stubs the packages
whose *behaviour* is not needed, this script supplies a small stand-in for the part of
``pymbolic`` those modules drive: expression nodes with operator overloading, structural
equality / hashing, and a differentiation mapper with the method names the reference's
``FieldDifferentiationMapper`` subclasses.  What the reference's code decides -- which terms
enter each equation, their signs and factors (``-2 H f'``, ``- a^2 dV/df``, ``16 pi S_ij``,
``tensor_index``, the kinetic / potential / gradient energy definitions, rho and P) -- is
therefore evaluated from the reference's source, not from a restatement; the stand-in's own
derivative rules are checked here against central finite differences before anything is
written.

The resulting fixture pins ``oracle.scalar_rhs / gw_rhs / scalar_energy / get_rho_and_p``
and the product's host-side sector algebra (``tests/test_reference_sectors_golden.py``).  Only run in
the build container (the reference is absent on the GPU box):

    python tests/golden/make_golden_sectors.py
"""

import json
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (stub machinery for loopy / pyopencl / mpi4py ...)

REFERENCE = MG.REFERENCE


# ---------------------------------------------------------------------------------------
# stand-in for the subset of pymbolic that pystella/field and pystella/sectors drive
# ---------------------------------------------------------------------------------------

def _is_number(x):
    return isinstance(x, (int, float, complex, np.number))


class Expression:
    """Immutable node; equality and hash are structural over ``__getinitargs__()``."""
    init_arg_names = ()
    mapper_method = None

    def __getinitargs__(self):
        return tuple(getattr(self, n) for n in self.init_arg_names)

    def __eq__(self, other):
        return type(self) is type(other) and \
            self.__getinitargs__() == other.__getinitargs__()

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((type(self).__name__,) + tuple(
            tuple(a) if isinstance(a, list) else a for a in self.__getinitargs__()))

    def __iter__(self):
        raise TypeError("expression nodes are not iterable")   # keeps numpy from unpacking

    # arithmetic ------------------------------------------------------------------------
    def __add__(self, other):
        return _sum(self, other)

    def __radd__(self, other):
        return _sum(other, self)

    def __sub__(self, other):
        return _sum(self, _product(-1, other))

    def __rsub__(self, other):
        return _sum(other, _product(-1, self))

    def __mul__(self, other):
        return _product(self, other)

    def __rmul__(self, other):
        return _product(other, self)

    def __truediv__(self, other):
        return Quotient(self, other)

    def __rtruediv__(self, other):
        return Quotient(other, self)

    def __pow__(self, other):
        return Power(self, other)

    def __rpow__(self, other):
        return Power(other, self)

    def __neg__(self):
        return _product(-1, self)

    def __pos__(self):
        return self

    def __getitem__(self, subscript):
        if isinstance(subscript, tuple) and len(subscript) == 0:
            return self
        return Subscript(self, subscript)

    def __call__(self, *parameters):
        return Call(self, tuple(parameters))

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__,
                           ", ".join(repr(a) for a in self.__getinitargs__()))


def _is_zero(x):
    return _is_number(x) and x == 0


def _sum(a, b):
    if _is_zero(a):
        return b
    if _is_zero(b):
        return a
    if _is_number(a) and _is_number(b):
        return a + b
    return Sum((a, b))


def _product(a, b):
    if _is_zero(a) or _is_zero(b):
        return 0
    if _is_number(a) and a == 1:
        return b
    if _is_number(b) and b == 1:
        return a
    if _is_number(a) and _is_number(b):
        return a * b
    return Product((a, b))


class Leaf(Expression):
    pass


class AlgebraicLeaf(Leaf):
    pass


class Variable(Leaf):
    init_arg_names = ("name",)
    mapper_method = "map_variable"

    def __init__(self, name):
        self.name = name


class Subscript(AlgebraicLeaf):
    init_arg_names = ("aggregate", "index")
    mapper_method = "map_subscript"

    def __init__(self, aggregate, index):
        self.aggregate, self.index = aggregate, index

    @property
    def index_tuple(self):
        return self.index if isinstance(self.index, tuple) else (self.index,)


class Lookup(AlgebraicLeaf):
    init_arg_names = ("aggregate", "name")
    mapper_method = "map_lookup"

    def __init__(self, aggregate, name):
        self.aggregate, self.name = aggregate, name


class Slice(Expression):
    init_arg_names = ("children",)
    mapper_method = "map_slice"

    def __init__(self, children):
        self.children = tuple(children)


class Call(AlgebraicLeaf):
    init_arg_names = ("function", "parameters")
    mapper_method = "map_call"

    def __init__(self, function, parameters):
        self.function, self.parameters = function, tuple(parameters)


class Sum(Expression):
    init_arg_names = ("children",)
    mapper_method = "map_sum"

    def __init__(self, children):
        self.children = tuple(children)


class Product(Expression):
    init_arg_names = ("children",)
    mapper_method = "map_product"

    def __init__(self, children):
        self.children = tuple(children)


class Quotient(Expression):
    init_arg_names = ("numerator", "denominator")
    mapper_method = "map_quotient"

    def __init__(self, numerator, denominator):
        self.numerator, self.denominator = numerator, denominator


class Power(Expression):
    init_arg_names = ("base", "exponent")
    mapper_method = "map_power"

    def __init__(self, base, exponent):
        self.base, self.exponent = base, exponent


class If(Expression):
    init_arg_names = ("condition", "then", "else_")
    mapper_method = "map_if"

    def __init__(self, condition, then, else_):
        self.condition, self.then, self.else_ = condition, then, else_


def make_variable(var_or_string):
    return var_or_string if isinstance(var_or_string, Expression) else Variable(var_or_string)


def parse(text):
    """Identifiers and integer literals (all that Field's offsets / indices / names use)."""
    text = text.strip()
    try:
        return int(text)
    except ValueError:
        pass
    if not text.isidentifier():
        raise ValueError(f"stand-in parser: unsupported expression {text!r}")
    return Variable(text)


class Mapper:
    """Dispatch on ``expr.mapper_method`` (numbers -> ``map_constant``, other foreign
    objects -> ``map_foreign``), the protocol the reference's mapper classes are written to."""

    def __init__(self, *args, **kwargs):
        self._cache = {}

    def get_cache_key(self, expr, *args, **kwargs):
        return (type(expr), expr, args, tuple(sorted(kwargs.items())))

    def __call__(self, expr, *args, **kwargs):
        if isinstance(expr, Expression):
            method = getattr(self, expr.mapper_method, None)
            if method is None:
                raise NotImplementedError(
                    f"{type(self).__name__} cannot handle {type(expr).__name__}")
            return method(expr, *args, **kwargs)
        return self.map_foreign(expr, *args, **kwargs)

    rec = __call__

    def map_foreign(self, expr, *args, **kwargs):
        if _is_number(expr):
            return self.map_constant(expr, *args, **kwargs)
        if isinstance(expr, (list, tuple)):
            return type(expr)(self.rec(e, *args, **kwargs) for e in expr)
        raise NotImplementedError(f"{type(self).__name__}: foreign object {expr!r}")


class Collector(Mapper):
    pass


class IdentityMapper(Mapper):
    pass


class CombineMapper(Mapper):
    pass


class SubstitutionMapper(Mapper):
    pass


class StringifyMapper(Mapper):
    pass


def make_subst_func(variable_assignments):
    return variable_assignments.get


_FUNCTION_DERIVATIVES = {
    "sin": lambda x: Variable("cos")(x),
    "cos": lambda x: -Variable("sin")(x),
    "exp": lambda x: Variable("exp")(x),
    "log": lambda x: 1 / x,
    "sinh": lambda x: Variable("cosh")(x),
    "cosh": lambda x: Variable("sinh")(x),
    "tanh": lambda x: 1 - Variable("tanh")(x)**2,
    "sqrt": lambda x: 1 / (2 * Variable("sqrt")(x)),
}


class DifferentiationMapper(Mapper):
    """d/d``variable`` with the sum / product / quotient / power / chain rules."""

    def __init__(self, variable, func_map=None, allowed_nonsmoothness="none"):
        super().__init__()
        self.variable = variable
        self.allowed_nonsmoothness = allowed_nonsmoothness

    def map_constant(self, expr, *args):
        return 0

    def map_variable(self, expr, *args):
        return 1 if expr == self.variable else 0

    def map_subscript(self, expr, *args):
        return 1 if expr == self.variable else 0

    def map_sum(self, expr, *args):
        out = 0
        for child in expr.children:
            out = _sum(out, self.rec(child, *args))
        return out

    def map_product(self, expr, *args):
        out = 0
        for i, child in enumerate(expr.children):
            d = self.rec(child, *args)
            if _is_zero(d):
                continue
            term = d
            for j, other in enumerate(expr.children):
                if j != i:
                    term = _product(term, other)
            out = _sum(out, term)
        return out

    def map_quotient(self, expr, *args):
        f, g = expr.numerator, expr.denominator
        df, dg = self.rec(f, *args), self.rec(g, *args)
        if _is_zero(dg):
            return 0 if _is_zero(df) else Quotient(df, g)
        return Quotient(_sum(_product(df, g), _product(-1, _product(f, dg))), Power(g, 2))

    def map_power(self, expr, *args):
        f, g = expr.base, expr.exponent
        df, dg = self.rec(f, *args), self.rec(g, *args)
        out = 0
        if not _is_zero(df):
            out = _product(_product(g, Power(f, _sum(g, -1))), df)
        if not _is_zero(dg):
            out = _sum(out, _product(_product(Variable("log")(f), Power(f, g)), dg))
        return out

    def map_call(self, expr, *args):
        (arg,) = expr.parameters
        darg = self.rec(arg, *args)
        if _is_zero(darg):
            return 0
        return _product(_FUNCTION_DERIVATIVES[expr.function.name](arg), darg)


def install_expression_modules():
    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    prim = module(
        "pymbolic.primitives", Expression=Expression, Leaf=Leaf, AlgebraicLeaf=AlgebraicLeaf,
        Variable=Variable, Subscript=Subscript, Lookup=Lookup, Slice=Slice, Call=Call, Sum=Sum,
        Product=Product, Quotient=Quotient, Power=Power, If=If, make_variable=make_variable,
        is_constant=_is_number, is_zero=_is_zero)
    mapper = module("pymbolic.mapper", Mapper=Mapper, Collector=Collector,
                    IdentityMapper=IdentityMapper, CombineMapper=CombineMapper)
    mapper.stringifier = module("pymbolic.mapper.stringifier", StringifyMapper=StringifyMapper)
    mapper.differentiator = module("pymbolic.mapper.differentiator",
                                   DifferentiationMapper=DifferentiationMapper)
    mapper.substitutor = module("pymbolic.mapper.substitutor",
                                SubstitutionMapper=SubstitutionMapper,
                                make_subst_func=make_subst_func)
    module("pymbolic", primitives=prim, mapper=mapper, var=Variable, parse=parse)
    # loopy itself stays a permissive stub (make_golden.py); only the three mapper base
    # classes pystella/field subclasses need to be real classes
    sys.modules["loopy.symbolic"] = types.ModuleType("loopy.symbolic")
    sys.modules["loopy.symbolic"].__dict__.update(
        IdentityMapper=IdentityMapper, CombineMapper=CombineMapper,
        SubstitutionMapper=SubstitutionMapper)


# ---------------------------------------------------------------------------------------
# numeric evaluation of the trees the reference builds
# ---------------------------------------------------------------------------------------

_NUMPY_FUNCTIONS = {n: getattr(np, n) for n in _FUNCTION_DERIVATIVES}


def evaluate(expr, data):
    """``data``: name -> scalar or array (first axes = the field's ``shape``)."""
    if _is_number(expr):
        return expr
    kind = expr.mapper_method
    if kind in ("map_field", "map_dynamic_field"):
        child = expr.child
        if isinstance(child, Subscript):
            return data[child.aggregate.name][child.index_tuple]
        return data[child.name]
    if kind == "map_variable":
        return data[expr.name]
    if kind == "map_subscript":
        return evaluate(expr.aggregate, data)[expr.index_tuple]
    if kind == "map_sum":
        out = evaluate(expr.children[0], data)
        for c in expr.children[1:]:
            out = out + evaluate(c, data)
        return out
    if kind == "map_product":
        out = evaluate(expr.children[0], data)
        for c in expr.children[1:]:
            out = out * evaluate(c, data)
        return out
    if kind == "map_quotient":
        return evaluate(expr.numerator, data) / evaluate(expr.denominator, data)
    if kind == "map_power":
        return evaluate(expr.base, data) ** evaluate(expr.exponent, data)
    if kind == "map_call":
        return _NUMPY_FUNCTIONS[expr.function.name](evaluate(expr.parameters[0], data))
    raise NotImplementedError(kind)


def key_name(expr):
    """'f[0]', 'dfdt[1]', 'dhijdt[3]' ... for the keys of an rhs_dict."""
    if isinstance(expr, Subscript):
        return "%s[%s]" % (key_name(expr.aggregate),
                           ", ".join(str(i) for i in expr.index_tuple))
    return expr.name


# ---------------------------------------------------------------------------------------
# cases
# ---------------------------------------------------------------------------------------

LATTICE = (3, 2, 4)

#: polynomial potentials as [(coefficient, (power of f_0, power of f_1, ...)), ...]
POTENTIALS = {
    # examples/scalar_preheating.py:98-105 with the bench's parameters (config 4)
    "preheating_2field": [(0.5, (2, 0)), (0.5 * 1.7e-2, (0, 2)), (0.5 * 2.5e3, (2, 2)),
                          (0.5 * 0.3, (1, 2)), (0.25 * 0.8, (0, 4))],
    # test/test_energy.py:49-51
    "test_energy_2field": [(0.5, (2, 0)), (0.5, (0, 2)), (0.5, (2, 2))],
    # single massive field (configs 2, 3, 5)
    "massive_1field": [(0.5, (2,))],
    # quartic self-interaction + mass, three fields with a cubic coupling
    "mixed_3field": [(0.5, (2, 0, 0)), (0.25 * 0.1, (4, 0, 0)), (0.7, (1, 1, 1)),
                     (1.3, (0, 3, 0)), (0.2, (0, 0, 2))],
}


def potential_from_terms(terms):
    """A callable of the kind users pass to ``ScalarSector(potential=...)``."""
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


def random_data(ns, seed):
    rng = np.random.default_rng(seed)
    data = {
        "f": rng.standard_normal((ns,) + LATTICE),
        "dfdt": rng.standard_normal((ns,) + LATTICE),
        "lap_f": rng.standard_normal((ns,) + LATTICE),
        "dfdx": rng.standard_normal((ns, 3) + LATTICE),
        "hij": rng.standard_normal((6,) + LATTICE),
        "dhijdt": rng.standard_normal((6,) + LATTICE),
        "lap_hij": rng.standard_normal((6,) + LATTICE),
        "a": float(1.0 + rng.random()),
        "hubble": float(0.1 + rng.random()),
    }
    return data


def check_derivative_rules():
    """The stand-in's differentiation against central differences (it is our code, not the
    reference's): polynomial, quotient, power and function nodes."""
    x, y = Variable("x"), Variable("y")
    exprs = [3 * x**4 * y - x * y**2 / 7 + 2, (x + y)**3 / (1 + x**2),
             Variable("cos")(x * y) + Variable("exp")(x) * Variable("tanh")(y),
             x**y + Variable("sqrt")(1 + x**2) * Variable("log")(2 + y**2)]
    pt = {"x": 0.73, "y": 1.19}
    for e in exprs:
        for v in ("x", "y"):
            d = evaluate(DifferentiationMapper(Variable(v))(e), pt)
            eps = 1e-6
            hi, lo = dict(pt), dict(pt)
            hi[v] += eps
            lo[v] -= eps
            fd = (evaluate(e, hi) - evaluate(e, lo)) / (2 * eps)
            assert abs(d - fd) < 1e-8 * max(1.0, abs(fd)), (e, v, d, fd)


def main():
    check_derivative_rules()
    install_expression_modules()
    sys.meta_path.insert(0, MG._StubFinder())
    sys.path.insert(0, REFERENCE)
    import pystella                     # the unmodified reference package
    from pystella import sectors as ref_sectors
    from pystella.field import DynamicField, Field, diff
    assert os.path.realpath(ref_sectors.__file__).startswith(os.path.realpath(REFERENCE))
    assert issubclass(DynamicField, Field) and issubclass(Field, AlgebraicLeaf)

    out = {"_comment": "generated by tests/golden/make_golden_sectors.py from the reference "
                       "sources (pystella/sectors.py, pystella/field); do not edit",
           "lattice": list(LATTICE), "cases": {}}

    for seed, (name, terms) in enumerate(sorted(POTENTIALS.items()), start=11):
        ns = len(terms[0][1])
        data = random_data(ns, seed)
        sector = pystella.ScalarSector(ns, potential=potential_from_terms(terms))
        rhs = sector.rhs_dict
        case = {"seed": seed, "nscalars": ns, "potential_terms": [[c, list(p)] for c, p in terms],
                "a": data["a"], "hubble": data["hubble"],
                "rhs_keys": [key_name(k) for k in rhs],
                "rhs": [np.broadcast_to(evaluate(v, data), LATTICE).tolist()
                        for v in rhs.values()]}
        red = sector.reducers
        case["reducers"] = {k: [np.broadcast_to(evaluate(e, data), LATTICE).tolist()
                                for e in v] for k, v in red.items()}
        # dV/df_n on its own (what the CUDA kernels' coefficient tables encode)
        f = sector.f
        case["dV"] = [np.broadcast_to(evaluate(diff(sector.potential(f), f[n]), data),
                                      LATTICE).tolist() for n in range(ns)]
        # T_ij without the trace term, all spatial pairs (source of the tensor sector)
        case["stress_tensor_drop_trace"] = {
            f"{i}{j}": evaluate(sector.stress_tensor(i, j, drop_trace=True), data).tolist()
            for i in range(1, 4) for j in range(i, 4)}
        if ns <= 2:
            gw = pystella.TensorPerturbationSector([sector])
            grhs = gw.rhs_dict
            case["gw_rhs_keys"] = [key_name(k) for k in grhs]
            case["gw_rhs"] = [evaluate(v, data).tolist() for v in grhs.values()]
        # sectors.get_rho_and_p on averaged components
        energy = {k: np.array([np.mean(np.array(x)) for x in v])
                  for k, v in case["reducers"].items()}
        full = ref_sectors.get_rho_and_p(dict(energy))
        case["rho_and_p"] = {"total": float(full["total"]), "pressure": float(full["pressure"])}
        out["cases"][name] = case

    out["tensor_index"] = {f"{i}{j}": ref_sectors.tensor_index(i, j)
                           for i in range(1, 4) for j in range(1, 4)}
    # a non-polynomial potential: dV/df of the reference's diff() for an axion-like cosine
    data = random_data(1, 99)
    cos = Variable("cos")
    sector = pystella.ScalarSector(1, potential=lambda f: 0.3 * (1 - cos(f[0] / 0.7)))
    rhs = sector.rhs_dict
    out["cosine_case"] = {
        "seed": 99, "a": data["a"], "hubble": data["hubble"], "Lambda4": 0.3, "f_a": 0.7,
        "rhs_keys": [key_name(k) for k in rhs],
        "rhs": [np.broadcast_to(evaluate(v, data), LATTICE).tolist() for v in rhs.values()]}

    path = os.path.join(HERE, "reference_sectors.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out["cases"]), "cases")
    assert math.isfinite(out["cases"]["massive_1field"]["rho_and_p"]["total"])


if __name__ == "__main__":
    main()
