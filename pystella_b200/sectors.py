"""
Physics sectors: which unknowns exist, what their right-hand sides are, and
which energy components get reduced.  API (class names, constructor arguments,
``rhs_dict`` / ``reducers`` / ``stress_tensor`` contents and their insertion
order) follows ``pystella/sectors.py:42-229``; the dictionaries produced here are
consumed by :mod:`pystella_b200.step` (which recognises the scalar / tensor
structure and dispatches the hand-written CUDA stage kernels) and by
:mod:`pystella_b200.reduction`.
"""

import math

from pystella_b200.field import DynamicField, Field, diff, var

__all__ = ["Sector", "ScalarSector", "TensorPerturbationSector", "tensor_index",
           "get_rho_and_p"]

#: index-free Fields are passed to kernels by value (sectors.py:120-121)
_HUBBLE = Field("hubble", indices=[])
_SCALE_FACTOR = Field("a", indices=[])


class Sector:
    """Interface of a sector (``pystella/sectors.py:42-89``)."""

    def __init__(self):
        raise NotImplementedError("Sector is an interface; use a subclass")

    @property
    def rhs_dict(self):
        """``{unknown: d(unknown)/dt}`` in evolution order."""
        raise NotImplementedError

    @property
    def reducers(self):
        """``{name: [expr | (expr, op), ...]}`` for :class:`Reduction`."""
        raise NotImplementedError

    def stress_tensor(self, mu, nu, drop_trace=True):
        """Component ``T_mu_nu`` (used as the source of tensor perturbations)."""
        raise NotImplementedError


class ScalarSector(Sector):
    """
    ``nscalars`` minimally coupled scalars in conformal FLRW
    (``pystella/sectors.py:92-161``).

    :arg nscalars: number of fields.
    :arg f: the :class:`DynamicField` (default
        ``DynamicField("f", offset="h", shape=(nscalars,))``).
    :arg potential: callable mapping the (subscriptable) field to ``V(f)``.
    """

    def __init__(self, nscalars, **kwargs):
        self.nscalars = nscalars
        default = DynamicField("f", offset="h", shape=(nscalars,))
        self.f = kwargs.pop("f", default)
        self.potential = kwargs.pop("potential", lambda x: 0)

    def _field_equation(self, n, V):
        """``dfdt_n' = lap_f_n - 2 H dfdt_n - a^2 dV/df_n`` (sectors.py:127-130)."""
        f = self.f
        friction = 2 * _HUBBLE * f.dot[n]
        force = _SCALE_FACTOR**2 * diff(V, f[n])
        return f.lap[n] - friction - force

    @property
    def rhs_dict(self):
        V = self.potential(self.f)
        out = {}
        for n in range(self.nscalars):
            out[self.f[n]] = self.f.dot[n]
            out[self.f.dot[n]] = self._field_equation(n, V)
        return out

    @property
    def reducers(self):
        # sectors.py:133-144 -- note `a` is a plain variable here, not a Field
        f, a = self.f, var("a")
        fields = range(self.nscalars)
        return {
            "kinetic": [f.dot[n]**2 / 2 / a**2 for n in fields],
            "potential": [self.potential(f)],
            "gradient": [- f[n] * f.lap[n] / 2 / a**2 for n in fields],
        }

    def stress_tensor(self, mu, nu, drop_trace=False):
        # sectors.py:146-161
        f, a = self.f, _SCALE_FACTOR
        fields = range(self.nscalars)
        T = sum(f.d(n, mu) * f.d(n, nu) for n in fields)
        if drop_trace:
            return T
        g_up = (-1 / a**2, 1 / a**2, 1 / a**2, 1 / a**2)
        g_dn = (-a**2, a**2, a**2, a**2)
        kinetic = sum(sum(g_up[al] * f.d(n, al) * f.d(n, al) for al in range(4))
                      for n in fields)
        lagrangian = - kinetic / 2 - self.potential(f)
        return T + (g_dn[mu] * lagrangian if mu == nu else 0)


def tensor_index(i, j):
    """Packed index of the symmetric pair (i, j), 1-based: (1,1)->0 ... (3,3)->5
    (``pystella/sectors.py:164-167``)."""
    lo, hi = (i, j) if i <= j else (j, i)
    return (7 - lo) * lo // 2 - 4 + hi


class TensorPerturbationSector:
    """
    Tensor perturbations ``hij`` (six packed components) sourced by the
    trace-dropped stress tensors of ``sectors`` (``pystella/sectors.py:170-208``).
    """

    def __init__(self, sectors, **kwargs):
        self.hij = kwargs.pop("hij", DynamicField("hij", offset="h", shape=(6,)))
        self.sectors = sectors

    @property
    def rhs_dict(self):
        hij = self.hij
        out = {}
        pairs = [(i, j) for i in range(1, 4) for j in range(i, 4)]
        for i, j in pairs:
            c = tensor_index(i, j)
            source = sum(s.stress_tensor(i, j, drop_trace=True) for s in self.sectors)
            out[hij[c]] = hij.dot[c]
            out[hij.dot[c]] = (hij.lap[c] - 2 * _HUBBLE * hij.dot[c]
                               + 16 * math.pi * source)
        return out

    @property
    def reducers(self):
        return {}


def get_rho_and_p(energy):
    """
    Reduction callback adding ``total`` = sum of all components and
    ``pressure`` = kinetic - gradient/3 - potential (``pystella/sectors.py:211-229``).
    """
    total = sum(sum(component) for component in energy.values())
    pressure = 0
    signs = {"kinetic": 1., "gradient": -1 / 3, "potential": -1.}
    for name in ("kinetic", "gradient", "potential"):
        if name in energy:
            if name == "gradient":
                pressure += - sum(energy[name]) / 3
            else:
                pressure += signs[name] * sum(energy[name])
    energy["total"] = total
    energy["pressure"] = pressure
    return energy
