"""
Finite differences: :class:`FiniteDifferencer` with the call contract of
``pystella/derivs.py:234-470`` on top of ``psb_finite_difference``.

The stencils are the centred differences whose tables live in the CUDA kernels
(``csrc/common.cuh``; reference tables ``derivs.py:127-131, 160-165``); the Python
side keeps the tables and their symbolic eigenvalues for users and tests.
"""

from itertools import product

import numpy as np

from pystella_b200 import _kernels
from pystella_b200.field import shift_fields

__all__ = ["FiniteDifferencer", "FirstCenteredDifference", "SecondCenteredDifference",
           "expand_stencil", "centered_diff"]


def expand_stencil(f, coefs):
    """
    ``sum_offset coefs[offset] * f(shifted by offset)`` as a symbolic expression;
    keys are 3-tuples (``pystella/derivs.py:37-58``).
    """
    out = 0
    for offset, c in coefs.items():
        out = out + c * shift_fields(f, shift=offset)
    return out


def centered_diff(f, coefs, direction, order):
    """
    Centred difference along ``direction`` (1, 2 or 3) from the non-redundant
    ``{offset: coef}``; the mirrored offsets get ``(-1)**order * coef``
    (``pystella/derivs.py:61-108``).
    """
    full = {}
    axis = direction - 1
    for s, c in coefs.items():
        if s != 0 or order % 2 == 0:
            key = [0, 0, 0]
            key[axis] = s
            full[tuple(key)] = c
        if s != 0:
            key = [0, 0, 0]
            key[axis] = -s
            full[tuple(key)] = (-1)**order * c
    return expand_stencil(f, full)


class FiniteDifferenceStencil:
    coefs = NotImplemented
    truncation_order = NotImplemented
    order = NotImplemented
    is_centered = NotImplemented

    def __call__(self, f, direction):
        if self.is_centered:
            return centered_diff(f, self.coefs, direction, self.order)
        return expand_stencil(f, self.coefs)

    def get_eigenvalues(self, k, dx):
        raise NotImplementedError


_GRAD = {
    1: {1: 1/2},
    2: {1: 8/12, 2: -1/12},
    3: {1: 45/60, 2: -9/60, 3: 1/60},
    4: {1: 672/840, 2: -168/840, 3: 32/840, 4: -3/840},
}

_LAP = {
    1: {0: -2, 1: 1},
    2: {0: -30/12, 1: 16/12, 2: -1/12},
    3: {0: -490/180, 1: 270/180, 2: -27/180, 3: 2/180},
    4: {0: -14350/5040, 1: 8064/5040, 2: -1008/5040, 3: 128/5040, 4: -9/5040},
}


class FirstCenteredDifference(FiniteDifferenceStencil):
    """Centred first derivative of radius ``h`` (accuracy order ``2h``)."""

    def __init__(self, h):
        self.h = h
        self.coefs = _GRAD[h]
        self.truncation_order = 2 * h
        self.order = 1
        self.is_centered = True

    def get_eigenvalues(self, k, dx):
        """Effective wavenumber of the stencil acting on ``exp(i k x)``: the
        Fourier symbol ``2 sum_s c_s sin(s k dx) / dx`` (cf. derivs.py:141-157)."""
        th = k * dx
        return sum(2 * c * np.sin(s * th) for s, c in self.coefs.items()) / dx


class SecondCenteredDifference(FiniteDifferenceStencil):
    """Centred second derivative of radius ``h`` (accuracy order ``2h``)."""

    def __init__(self, h):
        self.h = h
        self.coefs = _LAP[h]
        self.truncation_order = 2 * h
        self.order = 2
        self.is_centered = True

    def get_eigenvalues(self, k, dx):
        """Fourier symbol ``(c_0 + 2 sum_s c_s cos(s k dx)) / dx**2``
        (cf. derivs.py:175-191)."""
        th = k * dx
        val = self.coefs[0]
        for s, c in self.coefs.items():
            if s != 0:
                val = val + 2 * c * np.cos(s * th)
        return val / dx**2


class FiniteDifferencer:
    """
    Computes gradients, Laplacians and combinations (``pystella/derivs.py:234-429``).

    :arg decomp: a :class:`~pystella_b200.DomainDecomposition`; its
        ``share_halos`` is called on ``fx`` before every kernel, like the reference.
    :arg halo_shape: int, the padding of ``fx`` (and default stencil radius).
    :arg dx: 3-tuple of grid spacings.

    Keyword-only: ``rank_shape``, ``first_stencil`` / ``second_stencil``
    (:class:`FirstCenteredDifference` / :class:`SecondCenteredDifference` of a
    radius <= ``halo_shape``); ``stream``, ``device`` and ``*_lsize`` are accepted
    for signature compatibility and ignored (the kernels are hand-tuned).
    """

    def __init__(self, decomp, halo_shape, dx, stream=False, rank_shape=None,
                 device=None, first_stencil=None, second_stencil=None,
                 gradlap_lsize=None, grad_lsize=None, lap_lsize=None,
                 strict_fp=False, jit=None, dtype="float64", tune=None):
        if not isinstance(halo_shape, (int, np.integer)):
            raise NotImplementedError("halo_shape must currently be an int")
        self.decomp = decomp
        self.halo_shape = int(halo_shape)
        self.dx = tuple(float(d) for d in dx)
        self.rank_shape = rank_shape
        self.strict_fp = strict_fp
        #: kernel choice of psb_finite_difference_v: {"variant": 0 auto | 1 generic tile kernel}
        self.tune = dict(tune or {})
        first = first_stencil or FirstCenteredDifference(self.halo_shape)
        second = second_stencil or SecondCenteredDifference(self.halo_shape)
        self.first_stencil, self.second_stencil = first, second
        self._generated = None

        def standard(st, table):
            h = getattr(st, "h", None)
            return h is not None and getattr(st, "coefs", None) == table.get(h)

        if not (standard(first, _GRAD) and standard(second, _LAP) and first.h == second.h):
            # custom stencils (derivs.py:254-263): callables (f, direction) -> expression
            from pystella_b200.step import jit_enabled
            if not jit_enabled(jit):
                raise NotImplementedError(
                    "custom stencils need the run-time generated kernels, which were "
                    "switched off (jit=False / PSB_JIT=0); the hand-written kernels provide "
                    "the centred differences of radius 1..4")
            self._generated = self._generate_custom(first, second, dtype)
            self.h = self.halo_shape
            return
        self.h = first.h
        if self.h > self.halo_shape:
            raise ValueError("stencil radius exceeds halo_shape")

    def _custom_statements(self, first, second):
        """The symbolic kernels of the reference (derivs.py:291-312) for arbitrary stencils."""
        from pystella_b200.field import Field
        fx = Field("fx", offset="h")
        pd = {n: first(fx, i + 1) * (1 / dxi)
              for i, (n, dxi) in enumerate(zip(("pdx", "pdy", "pdz"), self.dx))}
        lap = 0
        for i, dxi in enumerate(self.dx):
            lap = lap + second(fx, i + 1) * dxi**-2
        return fx, pd, lap

    def _generate_custom(self, first, second, dtype):
        from pystella_b200.codegen import GenericMap
        from pystella_b200.field import Field
        _, pd, lap = self._custom_statements(first, second)
        kw = dict(dtype=dtype, halo_shape=self.halo_shape)
        maps = {
            ("lap",): GenericMap([(Field("lap"), lap)], **kw),
            ("pdx", "pdy", "pdz"): GenericMap([(Field(n), e) for n, e in pd.items()], **kw),
            ("lap", "pdx", "pdy", "pdz"): GenericMap(
                [(Field(n), e) for n, e in pd.items()] + [(Field("lap"), lap)], **kw),
        }
        for n, e in pd.items():
            maps[(n,)] = GenericMap([(Field(n), e)], **kw)
            maps[(n, "incr")] = GenericMap([(Field("div"), Field("div") + e)], **kw)
        return maps

    def _run(self, queue, fx, **outs):
        if self._generated is not None:
            key = tuple(n for n in ("lap", "pdx", "pdy", "pdz") if outs.get(n) is not None)
            if key not in self._generated:
                raise ValueError(f"unsupported output combination {key}")
            self._generated[key](queue, fx=fx, **{n: outs[n] for n in key})
            return
        _kernels.finite_difference(fx, h=self.h, halo=self.halo_shape, dx=self.dx,
                                   queue=queue, strict=self.strict_fp,
                                   variant=self.tune.get("variant", 0), **outs)

    def __call__(self, queue, fx, *, lap=None, pdx=None, pdy=None, pdz=None,
                 grd=None, allocator=None):
        """
        Valid output combinations: all of ``lap, pdx, pdy, pdz`` (or ``lap, grd``);
        any single one; ``pdx, pdy, pdz`` (or ``grd``).  Outer axes of ``fx`` are
        looped over; halos of each slice are shared first (derivs.py:339-429).
        """
        slices = list(product(*[range(n) for n in fx.shape[:-3]]))

        if grd is not None:
            if isinstance(grd, (tuple, list)):
                pdx, pdy, pdz = grd
            else:
                pdx = grd[..., 0, :, :, :]
                pdy = grd[..., 1, :, :, :]
                pdz = grd[..., 2, :, :, :]

        have_grad = pdx is not None and pdy is not None and pdz is not None
        for s in slices:
            self.decomp.share_halos(queue, fx[s])
            if lap is not None and have_grad:
                self._run(queue, fx[s], lap=lap[s], pdx=pdx[s], pdy=pdy[s], pdz=pdz[s])
            elif have_grad:
                self._run(queue, fx[s], pdx=pdx[s], pdy=pdy[s], pdz=pdz[s])
            elif lap is not None:
                self._run(queue, fx[s], lap=lap[s])
            elif pdx is not None:
                self._run(queue, fx[s], pdx=pdx[s])
            elif pdy is not None:
                self._run(queue, fx[s], pdy=pdy[s])
            elif pdz is not None:
                self._run(queue, fx[s], pdz=pdz[s])
        return None

    def divergence(self, queue, vec, div, allocator=None):
        """``div = d_x vec[0] + d_y vec[1] + d_z vec[2]`` (derivs.py:431-470)."""
        slices = list(product(*[range(n) for n in vec.shape[:-4]]))
        for s in slices:
            self.decomp.share_halos(queue, vec[s][0])
            self._run(queue, vec[s][0], pdx=div[s])
            for comp, name in ((1, "pdy"), (2, "pdz")):
                self.decomp.share_halos(queue, vec[s][comp])
                if self._generated is not None:
                    self._generated[(name, "incr")](queue, fx=vec[s][comp], div=div[s])
                else:
                    _kernels.finite_difference(vec[s][comp], h=self.h, halo=self.halo_shape,
                                               dx=self.dx, accumulate=True, queue=queue,
                                               strict=self.strict_fp, **{name: div[s]})
        return None
