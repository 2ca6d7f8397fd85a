"""
Symbolic front-end: :class:`Field`, :class:`DynamicField`, :func:`index_fields`,
:func:`shift_fields`, :func:`diff`, :func:`get_field_args`.

Mirrors the user-facing behaviour of ``pystella/field/__init__.py:52-606`` and
``pystella/field/diff.py:29-94``.  The reference builds on pymbolic (absent here
and not needed): expressions are a small self-contained tree whose only consumers
are (i) *recognition* of the systems the hand-written CUDA kernels implement
(:mod:`pystella_b200.step`, :mod:`pystella_b200.reduction`) through
:func:`to_polynomial`, and (ii) host-side evaluation of scalar ODEs
(:class:`pystella_b200.Expansion`), through :func:`evaluate`.
"""

import math
import numbers

import numpy as np

__all__ = [
    "Expr", "Variable", "var", "Sum", "Product", "Quotient", "Power", "Subscript",
    "Call", "Field", "DynamicField", "index_fields", "shift_fields", "substitute",
    "diff", "get_field_args", "evaluate", "to_polynomial", "ArrayRef", "is_constant",
    "parse",
]


def is_constant(x):
    return isinstance(x, (numbers.Number, np.number))


def _wrap(x):
    if isinstance(x, Expr) or is_constant(x):
        return x
    if isinstance(x, str):
        return parse(x)
    if isinstance(x, np.ndarray) and x.ndim == 0:
        return x.item()
    raise TypeError(f"cannot use {type(x)} in a symbolic expression")


class Expr:
    """Base class of expression nodes (immutable, structurally hashable)."""

    _fields = ()

    def _key(self):
        return (type(self).__name__,) + tuple(getattr(self, f) for f in self._fields)

    def __eq__(self, other):
        return isinstance(other, Expr) and self._key() == other._key()

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._key())

    # arithmetic -----------------------------------------------------------
    def __add__(self, other):
        other = _wrap(other)
        if is_constant(other) and other == 0:
            return self
        a = self.children if isinstance(self, Sum) else (self,)
        b = other.children if isinstance(other, Sum) else (other,)
        return Sum(a + b)

    def __radd__(self, other):
        other = _wrap(other)
        if is_constant(other) and other == 0:
            return self
        b = self.children if isinstance(self, Sum) else (self,)
        a = other.children if isinstance(other, Sum) else (other,)
        return Sum(a + b)

    def __sub__(self, other):
        other = _wrap(other)
        if is_constant(other) and other == 0:
            return self
        return Sum((self, -other))

    def __rsub__(self, other):
        return Sum((_wrap(other), -self))

    def __mul__(self, other):
        other = _wrap(other)
        if is_constant(other):
            if other == 1:
                return self
            if other == 0:
                return 0
        a = self.children if isinstance(self, Product) else (self,)
        b = other.children if isinstance(other, Product) else (other,)
        return Product(a + b)

    def __rmul__(self, other):
        other = _wrap(other)
        if is_constant(other):
            if other == 1:
                return self
            if other == 0:
                return 0
        b = self.children if isinstance(self, Product) else (self,)
        a = other.children if isinstance(other, Product) else (other,)
        return Product(a + b)

    def __truediv__(self, other):
        other = _wrap(other)
        if is_constant(other) and other == 1:
            return self
        return Quotient(self, other)

    def __rtruediv__(self, other):
        return Quotient(_wrap(other), self)

    def __pow__(self, other):
        other = _wrap(other)
        if is_constant(other):
            if other == 1:
                return self
            if other == 0:
                return 1
        return Power(self, other)

    def __rpow__(self, other):
        return Power(_wrap(other), self)

    def __neg__(self):
        return Product((-1, self))

    def __pos__(self):
        return self

    def __getitem__(self, index):
        if not isinstance(index, tuple):
            index = (index,)
        if index == ():
            return self
        return Subscript(self, tuple(_wrap(i) for i in index))

    def __call__(self, *args):
        return Call(self, tuple(_wrap(a) for a in args))

    def __repr__(self):
        return f"{type(self).__name__}({self})"

    def __str__(self):
        return _to_str(self, 0)

    def __bool__(self):
        return True


class Variable(Expr):
    _fields = ("name",)

    def __init__(self, name):
        self.name = name


def var(name):
    """Analogue of ``pymbolic.var``."""
    return Variable(name)


class Sum(Expr):
    _fields = ("children",)

    def __init__(self, children):
        self.children = tuple(children)


class Product(Expr):
    _fields = ("children",)

    def __init__(self, children):
        self.children = tuple(children)


class Quotient(Expr):
    _fields = ("numerator", "denominator")

    def __init__(self, numerator, denominator):
        self.numerator = numerator
        self.denominator = denominator


class Power(Expr):
    _fields = ("base", "exponent")

    def __init__(self, base, exponent):
        self.base = base
        self.exponent = exponent


class Subscript(Expr):
    _fields = ("aggregate", "index")

    def __init__(self, aggregate, index):
        self.aggregate = aggregate
        self.index = tuple(index) if isinstance(index, (tuple, list)) else (index,)

    @property
    def index_tuple(self):
        return self.index

    def __getitem__(self, index):
        # f[0][1] == f[0, 1]
        if not isinstance(index, tuple):
            index = (index,)
        return Subscript(self.aggregate, self.index + tuple(_wrap(i) for i in index))


class Call(Expr):
    _fields = ("function", "parameters")

    def __init__(self, function, parameters):
        self.function = function
        self.parameters = tuple(parameters)


# ---------------------------------------------------------------------------
# a tiny parser for names / subscripts / index arithmetic ("i", "f[0]", "i + h")
# ---------------------------------------------------------------------------

def parse(s):
    import ast

    def conv(node):
        if isinstance(node, ast.Expression):
            return conv(node.body)
        if isinstance(node, ast.Name):
            return Variable(node.id)
        if isinstance(node, ast.Constant):
            return node.value
        if isinstance(node, ast.BinOp):
            a, b = conv(node.left), conv(node.right)
            ops = {ast.Add: lambda: a + b, ast.Sub: lambda: a - b,
                   ast.Mult: lambda: a * b, ast.Div: lambda: a / b,
                   ast.Pow: lambda: a ** b}
            for k, fn in ops.items():
                if isinstance(node.op, k):
                    return fn()
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub):
            v = conv(node.operand)
            return -v
        if isinstance(node, ast.Subscript):
            idx = node.slice
            idx = tuple(conv(e) for e in idx.elts) if isinstance(idx, ast.Tuple) \
                else (conv(idx),)
            return Subscript(conv(node.value), idx)
        if isinstance(node, ast.Call):
            return Call(conv(node.func), tuple(conv(a) for a in node.args))
        raise ValueError(f"cannot parse {s!r}")

    return conv(ast.parse(s, mode="eval"))


def parse_if_str(expr):
    return parse(expr) if isinstance(expr, str) else expr


# ---------------------------------------------------------------------------
# Field / DynamicField  (pystella/field/__init__.py:52-300)
# ---------------------------------------------------------------------------

class Field(Expr):
    """
    An expression leaf that mimics an array by carrying indexing information
    (``pystella/field/__init__.py:52-194``): ``offset`` is the halo padding of the
    looped-over axes, ``shape`` the shape of leading (outer) axes, ``indices`` the
    loop indices (``()`` for scalars passed by value such as ``a`` / ``hubble``).
    """

    init_arg_names = ("child", "offset", "shape", "indices",
                      "ignore_prepends", "base_offset", "dtype", "dim_tags")
    _fields = ("child", "offset", "shape", "indices", "ignore_prepends", "base_offset")

    def __init__(self, child, offset=0, shape=tuple(), indices=("i", "j", "k"),
                 ignore_prepends=False, base_offset=None, dtype=None, dim_tags=None):
        self.child = parse_if_str(child)
        if isinstance(self.child, Subscript):
            self.name = self.child.aggregate.name
        else:
            self.name = self.child.name

        indices = tuple(indices)
        if not isinstance(offset, (list, tuple)):
            offset = (offset,) * len(indices)
        if len(offset) != len(indices):
            raise ValueError(
                "offset (if not length-1) must have same length as indices")

        self.offset = tuple(parse_if_str(o) for o in offset)
        self.base_offset = tuple(base_offset) if base_offset else self.offset
        self.indices = tuple(parse_if_str(i) for i in indices)
        self.shape = tuple(shape)
        self.ignore_prepends = ignore_prepends
        self.dtype = dtype
        self.dim_tags = dim_tags

    def __getinitargs__(self):
        return (self.child, self.offset, self.shape, self.indices,
                self.ignore_prepends, self.base_offset, self.dtype, self.dim_tags)

    @property
    def index_tuple(self):
        """The fully expanded subscript: ``indices`` offset by ``offset``."""
        return tuple(i + o for i, o in zip(self.indices, self.offset))

    def copy(self, **kwargs):
        init_kwargs = dict(zip(self.init_arg_names, self.__getinitargs__()))
        init_kwargs.update(kwargs)
        return type(self)(**init_kwargs)


class DynamicField(Field):
    """
    A :class:`Field` that also carries the Fields of its time derivative
    (``dot``, named ``d{child}dt``, same padding), Laplacian (``lap``, named
    ``lap_{child}``, unpadded) and gradient (``pd``, named ``d{child}dx``,
    shape ``shape + (3,)``, unpadded) -- ``pystella/field/__init__.py:204-300``.
    """

    init_arg_names = ("child", "offset", "shape", "indices", "base_offset",
                      "dot", "lap", "pd", "dtype")

    def __init__(self, child, offset="0", shape=tuple(), indices=("i", "j", "k"),
                 base_offset=None, dot=None, lap=None, pd=None, dtype=None):
        super().__init__(child, offset=offset, indices=indices,
                         base_offset=base_offset, shape=shape, dtype=dtype)
        shape = tuple(shape)
        self.dot = dot or Field(f"d{child}dt", shape=shape, offset=offset,
                                indices=indices, dtype=dtype)
        self.lap = lap or Field(f"lap_{child}", shape=shape, offset=0,
                                indices=indices, ignore_prepends=True, dtype=dtype)
        self.pd = pd or Field(f"d{child}dx", shape=shape + (3,), offset=0,
                              indices=indices, ignore_prepends=True, dtype=dtype)

    def __getinitargs__(self):
        return (self.child, self.offset, self.shape, self.indices, self.base_offset,
                self.dot, self.lap, self.pd, self.dtype)

    def d(self, *args):
        """
        ``f.d(0)`` is ``dfdt``; ``f.d(mu)`` is ``dfdx[mu-1]`` for mu = 1..3; leading
        arguments subscript the outer axes (``pystella/field/__init__.py:260-298``).
        """
        mu = args[-1]
        indices = args[:-1] + (mu - 1,)
        return self.dot[args[:-1]] if mu == 0 else self.pd[indices]


# ---------------------------------------------------------------------------
# printing
# ---------------------------------------------------------------------------

_PREC_SUM, _PREC_PROD, _PREC_POW, _PREC_ATOM = 1, 2, 3, 4


def _to_str(e, prec):
    if is_constant(e):
        s = repr(e) if not isinstance(e, np.generic) else repr(e.item())
        return f"({s})" if (e < 0 and prec > _PREC_SUM) else s
    if isinstance(e, Field):
        return _to_str(e.child, prec)
    if isinstance(e, Variable):
        return e.name
    if isinstance(e, Subscript):
        return f"{_to_str(e.aggregate, _PREC_ATOM)}[" + ", ".join(
            _to_str(i, 0) for i in e.index) + "]"
    if isinstance(e, Call):
        return f"{_to_str(e.function, _PREC_ATOM)}(" + ", ".join(
            _to_str(p, 0) for p in e.parameters) + ")"
    if isinstance(e, Sum):
        s = " + ".join(_to_str(c, _PREC_SUM) for c in e.children)
        return f"({s})" if prec > _PREC_SUM else s
    if isinstance(e, Product):
        s = "*".join(_to_str(c, _PREC_PROD) for c in e.children)
        return f"({s})" if prec > _PREC_PROD else s
    if isinstance(e, Quotient):
        s = f"{_to_str(e.numerator, _PREC_PROD)}/{_to_str(e.denominator, _PREC_POW)}"
        return f"({s})" if prec > _PREC_PROD else s
    if isinstance(e, Power):
        s = f"{_to_str(e.base, _PREC_ATOM)}**{_to_str(e.exponent, _PREC_POW)}"
        return f"({s})" if prec > _PREC_POW else s
    return str(e)


# ---------------------------------------------------------------------------
# generic traversal
# ---------------------------------------------------------------------------

def _map(e, fn):
    """Rebuild ``e`` bottom-up; ``fn(node)`` may return a replacement or None."""
    r = fn(e)
    if r is not None:
        return r
    if is_constant(e) or isinstance(e, (Variable, Field)):
        return e
    if isinstance(e, Sum):
        out = 0
        for c in e.children:
            out = out + _map(c, fn)
        return out
    if isinstance(e, Product):
        out = 1
        for c in e.children:
            out = out * _map(c, fn)
        return out
    if isinstance(e, Quotient):
        return _map(e.numerator, fn) / _map(e.denominator, fn)
    if isinstance(e, Power):
        return _map(e.base, fn) ** _map(e.exponent, fn)
    if isinstance(e, Subscript):
        return Subscript(_map(e.aggregate, fn), tuple(_map(i, fn) for i in e.index))
    if isinstance(e, Call):
        return Call(e.function, tuple(_map(p, fn) for p in e.parameters))
    if isinstance(e, dict):
        return {_map(k, fn): _map(v, fn) for k, v in e.items()}
    if isinstance(e, (list, tuple)):
        return type(e)(_map(x, fn) for x in e)
    raise TypeError(f"unknown expression node {type(e)}")


def _walk(e):
    yield e
    if isinstance(e, (Sum, Product)):
        for c in e.children:
            yield from _walk(c)
    elif isinstance(e, Quotient):
        yield from _walk(e.numerator)
        yield from _walk(e.denominator)
    elif isinstance(e, Power):
        yield from _walk(e.base)
        yield from _walk(e.exponent)
    elif isinstance(e, Subscript):
        yield from _walk(e.aggregate)
        for i in e.index:
            yield from _walk(i)
    elif isinstance(e, Call):
        for p in e.parameters:
            yield from _walk(p)
    elif isinstance(e, dict):
        for k, v in e.items():
            yield from _walk(k)
            yield from _walk(v)
    elif isinstance(e, (list, tuple)):
        for x in e:
            yield from _walk(x)


def index_fields(expr, prepend_with=None):
    """
    Append the subscripts implied by ``indices``/``offset`` to every Field
    (``pystella/field/__init__.py:405-468``)::

        >>> print(index_fields(Field("f", offset="h")))
        f[i + h, j + h, k + h]
    """
    prepend_with = tuple(prepend_with or ())

    def fn(e):
        if isinstance(e, Subscript) and isinstance(e.aggregate, Field):
            fld = e.aggregate
            pre = () if fld.ignore_prepends else prepend_with
            return _index_child(fld, pre + tuple(e.index), fld.index_tuple)
        if isinstance(e, Field):
            pre = () if e.ignore_prepends else prepend_with
            return _index_child(e, pre, e.index_tuple)
        return None

    return _map(expr, fn)


def prepend_subscript(expr, prepend_with):
    """
    The ``prepend_with`` half of :func:`index_fields` alone: every Field that does not
    set ``ignore_prepends`` gains the leading subscript, and stays a Field (so that
    stencil shifts and offsets survive for the code generator).
    """
    pre = tuple(prepend_with)

    def fn(e):
        if isinstance(e, Subscript) and isinstance(e.aggregate, Field):
            fld = e.aggregate
            idx = tuple(_map(i, fn) for i in e.index)
            return Subscript(fld, (() if fld.ignore_prepends else pre) + idx)
        if isinstance(e, Field):
            return e if e.ignore_prepends else Subscript(e, pre)
        return None

    return _map(expr, fn)


def _index_child(fld, outer, spatial):
    # prepended + user subscripts, then the child's own subscript, then i, j, k
    child = fld.child
    own = ()
    if isinstance(child, Subscript):
        own = tuple(child.index)
        child = child.aggregate
    full = tuple(outer) + own + tuple(spatial)
    return Subscript(child, full) if full else child


def shift_fields(expr, shift):
    """
    Offset every Field's ``offset`` by ``shift`` (a 3-tuple)
    (``pystella/field/__init__.py:471-491``).
    """
    def fn(e):
        if isinstance(e, Field) and len(e.indices) == len(shift):
            new = tuple(o + s for o, s in zip(e.offset, shift))
            return e.copy(offset=new, base_offset=e.base_offset)
        return None

    return _map(expr, fn)


def substitute(expr, variable_assignments=None, **kwargs):
    """Replace Variables / Fields (by name or by node) with expressions."""
    assignments = dict(variable_assignments or {})
    assignments.update(kwargs)

    def fn(e):
        if isinstance(e, (Variable, Field)):
            if e in assignments:
                return _wrap(assignments[e])
            name = e.name if isinstance(e, Variable) else None
            if name in assignments:
                return _wrap(assignments[name])
        return None

    return _map(expr, fn)


# ---------------------------------------------------------------------------
# differentiation  (pystella/field/diff.py:29-94)
# ---------------------------------------------------------------------------

_XMU = {"t": 0, "x": 1, "y": 2, "z": 3}


# d f(u) / du for the elementary functions the code generator knows
_DERIVATIVES = {
    "sin": lambda u: var("cos")(u),
    "cos": lambda u: -1 * var("sin")(u),
    "tan": lambda u: 1 + var("tan")(u)**2,
    "exp": lambda u: var("exp")(u),
    "log": lambda u: 1 / u,
    "sqrt": lambda u: 1 / (2 * var("sqrt")(u)),
    "sinh": lambda u: var("cosh")(u),
    "cosh": lambda u: var("sinh")(u),
    "tanh": lambda u: 1 - var("tanh")(u)**2,
}


def _diff1(e, x):
    coord = x.name if isinstance(x, Variable) and x.name in _XMU else None
    if is_constant(e):
        return 0
    if isinstance(e, Subscript) and isinstance(e.aggregate, DynamicField) \
            and coord is not None:
        return e.aggregate.d(*e.index, _XMU[coord])
    if isinstance(e, DynamicField) and coord is not None:
        return e.d(_XMU[coord])
    if isinstance(e, (Variable, Field, Subscript)):
        return 1 if e == x else 0
    if isinstance(e, Sum):
        out = 0
        for c in e.children:
            out = out + _diff1(c, x)
        return out
    if isinstance(e, Product):
        out = 0
        for n, c in enumerate(e.children):
            dc = _diff1(c, x)
            if is_constant(dc) and dc == 0:
                continue
            term = dc
            for m, o in enumerate(e.children):
                if m != n:
                    term = term * o
            out = out + term
        return out
    if isinstance(e, Quotient):
        du, dv = _diff1(e.numerator, x), _diff1(e.denominator, x)
        if is_constant(dv) and dv == 0:
            return du / e.denominator
        return (du * e.denominator - e.numerator * dv) / e.denominator**2
    if isinstance(e, Power):
        db, de = _diff1(e.base, x), _diff1(e.exponent, x)
        if not (is_constant(de) and de == 0):
            raise NotImplementedError("derivative of a non-constant exponent")
        if is_constant(db) and db == 0:
            return 0
        return e.exponent * e.base**(e.exponent - 1) * db
    if isinstance(e, Call):
        name = getattr(e.function, "name", None)
        if len(e.parameters) != 1 or name not in _DERIVATIVES:
            raise NotImplementedError(f"derivative of a call to {e.function}")
        u = e.parameters[0]
        du = _diff1(u, x)
        if is_constant(du) and du == 0:
            return 0
        return _DERIVATIVES[name](u) * du
    raise TypeError(f"cannot differentiate {type(e)}")


def diff(f, *x):
    """
    ``diff(f, x)`` = df/dx understanding Fields; for ``x`` in ``t, x, y, z`` and a
    DynamicField the corresponding derivative Field is returned
    (``pystella/field/diff.py:62-94``).  A list/np.ndarray ``f`` is handled
    element-wise.
    """
    if len(x) > 1:
        return diff(diff(f, x[0]), *x[1:])
    xx = parse_if_str(x[0])
    return _diff1(_wrap(f), xx)


# ---------------------------------------------------------------------------
# evaluation on the host (scalar ODEs: Expansion, stepper order tests)
# ---------------------------------------------------------------------------

_FUNCS = {"fabs": np.abs, "abs": np.abs, "sqrt": np.sqrt, "exp": np.exp, "log": np.log,
          "sin": np.sin, "cos": np.cos, "tan": np.tan, "tanh": np.tanh,
          "sinh": np.sinh, "cosh": np.cosh}


def evaluate(e, context):
    """
    Evaluate ``e`` with NumPy; ``context`` maps names to numbers / arrays.
    Fields evaluate to ``context[name]`` (subscripted by any constant index).
    """
    if is_constant(e):
        return e
    if isinstance(e, Field):
        val = context[e.name]
        if isinstance(e.child, Subscript):
            val = val[tuple(int(evaluate(i, context)) for i in e.child.index)]
        return val
    if isinstance(e, Variable):
        if e.name in context:
            return context[e.name]
        if e.name == "pi":
            return math.pi
        raise KeyError(f"no value supplied for '{e.name}'")
    if isinstance(e, Subscript):
        idx = tuple(int(evaluate(i, context)) for i in e.index)
        agg = e.aggregate
        if isinstance(agg, Field) and isinstance(agg.child, Subscript):
            # user subscripts come before the Field's own (index_fields order)
            own = tuple(int(evaluate(i, context)) for i in agg.child.index)
            return context[agg.name][idx + own]
        val = evaluate(agg, context)
        if isinstance(val, (list, tuple)):       # evaluate_kw(x, f0=[...]) in the examples
            for i in idx:
                val = val[i]
            return val
        return val[idx]
    if isinstance(e, Sum):
        it = iter(e.children)
        out = evaluate(next(it), context)
        for c in it:
            out = out + evaluate(c, context)
        return out
    if isinstance(e, Product):
        it = iter(e.children)
        out = evaluate(next(it), context)
        for c in it:
            out = out * evaluate(c, context)
        return out
    if isinstance(e, Quotient):
        return evaluate(e.numerator, context) / evaluate(e.denominator, context)
    if isinstance(e, Power):
        return evaluate(e.base, context) ** evaluate(e.exponent, context)
    if isinstance(e, Call):
        fn = _FUNCS[e.function.name]
        return fn(*[evaluate(p, context) for p in e.parameters])
    raise TypeError(f"cannot evaluate {type(e)}")


# ---------------------------------------------------------------------------
# polynomial normal form over *array* atoms with scalar-expression coefficients
# ---------------------------------------------------------------------------

class ArrayRef:
    """
    One array operand of a kernel: Field name, constant outer subscript, halo
    padding spec (``offset`` as given by the user, e.g. ``h``/``0``) and stencil
    shift.  ``absolute`` marks the use through ``fabs``.
    """

    __slots__ = ("name", "subscript", "offset", "shift", "field")

    def __init__(self, name, subscript, offset, shift, field):
        self.name = name
        self.subscript = tuple(subscript)
        self.offset = offset
        self.shift = tuple(shift)
        self.field = field

    def _key(self):
        return (self.name, self.subscript, self.shift)

    def __eq__(self, other):
        return isinstance(other, ArrayRef) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __lt__(self, other):
        return self._key() < other._key()

    def __repr__(self):
        sub = "".join(f"[{s}]" for s in self.subscript)
        sh = "" if not any(self.shift) else f"@{self.shift}"
        return f"{self.name}{sub}{sh}"


def _is_array_field(fld):
    return isinstance(fld, Field) and len(fld.indices) > 0


def _field_shift(fld):
    shift = []
    for o, b in zip(fld.offset, fld.base_offset):
        d = o - b if not (is_constant(o) and is_constant(b)) else o - b
        d = _fold(d)
        if not is_constant(d):
            raise NotImplementedError(f"non-constant stencil shift {d}")
        shift.append(int(d))
    return tuple(shift)


def _fold(e):
    """Constant-fold an index expression like ``h + 1 + -1*h``."""
    if is_constant(e):
        return e
    try:
        poly = to_polynomial(e, arrays_ok=False)
    except NotImplementedError:
        return e
    # collapse: all coefficient exprs; try numeric
    total = 0
    for mono, c in poly.items():
        if mono != ():
            return e
        total = total + c
    if is_constant(total):
        return total
    # symbolic cancellation (h - h): evaluate with random substitutions
    names = sorted({n.name for n in _walk(total) if isinstance(n, Variable)})
    vals = []
    for seed in (0.37, 1.91):
        ctx = {n: seed * (k + 1.123) for k, n in enumerate(names)}
        vals.append(evaluate(total, ctx))
    if abs(vals[0] - vals[1]) < 1e-12 and abs(vals[0] - round(vals[0])) < 1e-12:
        return int(round(vals[0]))
    return e


def _poly_mul(p, q):
    out = {}
    for m1, c1 in p.items():
        for m2, c2 in q.items():
            d = dict(m1)
            for a, pw in m2:
                d[a] = d.get(a, 0) + pw
            mono = tuple(sorted(d.items(), key=lambda kv: kv[0]._key()))
            c = c1 * c2
            if mono in out:
                out[mono] = out[mono] + c
            else:
                out[mono] = c
    return out


def _poly_add(p, q, sign=1):
    out = dict(p)
    for m, c in q.items():
        c = c if sign == 1 else -c
        out[m] = out[m] + c if m in out else c
    return out


def to_polynomial(e, arrays_ok=True):
    """
    Normal form ``{monomial: coefficient}`` where a monomial is a sorted tuple of
    ``(ArrayRef, power)`` and coefficients are numbers or scalar expressions
    (Variables and index-free Fields such as ``a`` / ``hubble``).
    Raises ``NotImplementedError`` for expressions that are not polynomial in
    the arrays (those need a generated kernel, which this build does not have).
    """
    if is_constant(e):
        return {(): e}
    if isinstance(e, Field):
        if _is_array_field(e):
            if not arrays_ok:
                raise NotImplementedError("array in scalar context")
            sub = e.child.index if isinstance(e.child, Subscript) else ()
            ref = ArrayRef(e.name, _const_index(sub), e.base_offset, _field_shift(e), e)
            return {((ref, 1),): 1}
        return {(): e}
    if isinstance(e, Variable):
        return {(): e}
    if isinstance(e, Subscript):
        agg = e.aggregate
        if _is_array_field(agg):
            if not arrays_ok:
                raise NotImplementedError("array in scalar context")
            sub = agg.child.index if isinstance(agg.child, Subscript) else ()
            ref = ArrayRef(agg.name, _const_index(tuple(e.index) + tuple(sub)),
                           agg.base_offset, _field_shift(agg), agg)
            return {((ref, 1),): 1}
        return {(): e}
    if isinstance(e, Sum):
        out = {}
        for c in e.children:
            out = _poly_add(out, to_polynomial(c, arrays_ok))
        return out
    if isinstance(e, Product):
        out = {(): 1}
        for c in e.children:
            out = _poly_mul(out, to_polynomial(c, arrays_ok))
        return out
    if isinstance(e, Quotient):
        num = to_polynomial(e.numerator, arrays_ok)
        den = to_polynomial(e.denominator, arrays_ok)
        if set(den.keys()) != {()}:
            raise NotImplementedError(
                f"division by an array expression ({e.denominator}) is not polynomial")
        inv = 1 / den[()]
        return {m: c * inv for m, c in num.items()}
    if isinstance(e, Power):
        base = to_polynomial(e.base, arrays_ok)
        ex = e.exponent
        if set(base.keys()) == {()}:
            return {(): base[()] ** ex}
        if is_constant(ex) and float(ex).is_integer() and ex >= 0:
            out = {(): 1}
            for _ in range(int(ex)):
                out = _poly_mul(out, base)
            return out
        raise NotImplementedError(
            f"power {ex} of an array expression is not polynomial")
    if isinstance(e, Call):
        params = [to_polynomial(p, arrays_ok) for p in e.parameters]
        if all(set(p.keys()) == {()} for p in params):
            return {(): e}
        raise NotImplementedError(f"function {e.function} of an array expression")
    raise TypeError(f"cannot convert {type(e)} to a polynomial")


def _const_index(idx):
    out = []
    for i in idx:
        i = _fold(i)
        if not is_constant(i):
            raise NotImplementedError(f"non-constant array subscript {i}")
        out.append(int(i))
    return tuple(out)


# ---------------------------------------------------------------------------
# argument inference  (pystella/field/__init__.py:529-606)
# ---------------------------------------------------------------------------

class FieldArg:
    """Description of one kernel argument inferred from the Fields of an expression."""

    def __init__(self, name, shape, offset, dtype, is_value):
        self.name = name
        self.shape = shape
        self.offset = offset
        self.dtype = dtype
        self.is_value = is_value

    def __repr__(self):
        kind = "ValueArg" if self.is_value else "GlobalArg"
        return f"{kind}({self.name!r}, shape={self.shape})"


def get_field_args(expressions, unpadded_shape=None, prepend_with=None):
    """
    Collect all Fields of ``expressions`` and describe the arrays they index:
    padded shape ``shape + (N_i + 2*base_offset_i)``, sorted by name; Fields with
    no indices and no shape are value arguments
    (``pystella/field/__init__.py:536-606``).
    """
    unpadded_shape = unpadded_shape or ("Nx", "Ny", "Nz")
    prepend = tuple(prepend_with or ())
    found = {}
    for node in _walk(expressions):
        if isinstance(node, Field):
            found.setdefault(node.name, node)
            if isinstance(node, DynamicField):
                pass
    args = []
    for name in sorted(found):
        f = found[name]
        spatial = tuple(_wrap(N) + 2 * o if not (is_constant(o) and o == 0) else N
                        for N, o in zip(unpadded_shape, f.base_offset))
        pre = () if f.ignore_prepends else prepend
        full = pre + tuple(f.shape) + (spatial if len(f.indices) else ())
        args.append(FieldArg(name, full, f.base_offset, f.dtype,
                             is_value=(len(full) == 0)))
    return args


def collect_field_indices(expressions):
    """Names of the index variables of every Field in ``expressions`` -- the loop
    variables a kernel over them needs (``pystella/field/__init__.py:609-624``)."""
    names = set()
    for node in _walk(expressions):
        if isinstance(node, Field):
            for idx in node.indices:
                names.add(getattr(idx, "name", str(idx)))
    return names


def indices_to_domain(indices):
    """ISL-style domain string over ``indices`` (``field/__init__.py:627-630``)."""
    constraints = " and ".join(f"0 <= {idx} < N{idx}" for idx in indices)
    return "{{[{}]: {}}}".format(",".join(indices), constraints)


def infer_field_domains(expressions):
    """``field/__init__.py:633-635``."""
    return indices_to_domain(sorted(collect_field_indices(expressions)))
