"""
Time steppers.  Same classes, constructor arguments, ``__call__(stage, ...)``
contract, ``num_stages`` / ``expected_order`` / ``num_unknowns`` / ``tmp_arrays``
attributes as ``pystella/step.py:67-853``.

What differs is *what a stage is*: the reference generates one OpenCL kernel per
stage from the symbolic ``rhs_dict``; here ``rhs_dict`` is *recognised*
(:func:`recognise_scalar_system`) and each stage dispatches a hand-written sm_100a
kernel through the C ABI:

* ``MapKernel=ElementWiseMap`` (default): the element-wise stage that reads a
  precomputed Laplacian, as in the reference (``psb_scalar_stage_unfused``);
* ``MapKernel=FusedStencilStage`` (+ ``dx=``, optionally ``decomp=``): the fused
  stencil + RHS + update kernel (``psb_scalar_stage_fused``) -- the reference's
  documented ``MapKernel`` hook (``pystella/step.py:94-103``) is where it plugs in.

Scalar ODEs on host NumPy arrays (``rank_shape=(0, 0, 0)``; the reference runs
these through loopy's C target, e.g. for :class:`~pystella_b200.Expansion`) are
evaluated on the host.
"""

import numpy as np
import torch

from pystella_b200 import _kernels, _lib
from pystella_b200._lib import PsbError
from pystella_b200.field import (DynamicField, Field, Subscript, Variable, evaluate,
                                 is_constant, to_polynomial)

__all__ = [
    "ElementWiseMap", "FusedStencilStage", "Stepper", "RungeKuttaStepper",
    "LowStorageRKStepper", "RungeKutta4", "RungeKutta3Heun", "RungeKutta3Nystrom",
    "RungeKutta3Ralston", "RungeKutta3SSP", "RungeKutta2Midpoint", "RungeKutta2Heun",
    "RungeKutta2Ralston", "LowStorageRK54", "LowStorageRK144", "LowStorageRK134",
    "LowStorageRK124", "LowStorageRK3Williamson", "LowStorageRK3Inhomogeneous",
    "LowStorageRK3Symmetric", "LowStorageRK3PredictorCorrector", "LowStorageRK3SSP",
    "all_steppers", "recognise_scalar_system",
]


def jit_enabled(flag=None):
    """Run-time generated kernels (NVRTC) are the DEFAULT fallback for everything outside the
    hand-written menu -- as loopy is for the reference (``pystella/elementwise.py:164-235``).
    ``jit=False`` (per object) or ``PSB_JIT=0`` (process-wide) turns them off: systems outside
    the menu then raise :class:`NotImplementedError` naming the offending term.  Measured on
    B200 (512^3): the generated fused stage runs within 0.3 % of the hand-written instance."""
    import os
    if flag is not None:
        return bool(flag)
    return os.environ.get("PSB_JIT", "1") != "0"


#: programmatic dependent launch of the fused stage kernels; per stepper: ``tune={"pdl": bool}``
DEFAULT_PDL = False


class ElementWiseMap:
    """
    "One element-wise kernel per stage" (``pystella/elementwise.py:81-361``).

    As a ``MapKernel`` *class* it selects the element-wise stage of a stepper: the
    hand-written scalar-sector / tensor-sector kernels for the systems on the menu.
    Instantiated with statements (``ElementWiseMap(map_instructions,
    tmp_instructions=..., halo_shape=..., rank_shape=..., dtype=...)``) it is the
    reference's generic map, served by the run-time generated kernels of
    :mod:`pystella_b200.codegen` (NVRTC, compiled once and cached on disk; ``jit=False`` /
    ``PSB_JIT=0`` disables them and the call raises :class:`NotImplementedError`).
    """

    def __init__(self, map_instructions=None, **kwargs):
        self._generic = None
        if map_instructions is None:
            return
        if not jit_enabled(kwargs.pop("jit", None)):
            raise NotImplementedError(
                "generic ElementWiseMap needs the run-time generated kernels, which were "
                "switched off (jit=False / PSB_JIT=0); the hand-written kernels cover "
                "ScalarSector, TensorPerturbationSector and the wave equation")
        from pystella_b200.codegen import GenericMap
        self._generic = GenericMap(
            map_instructions, kwargs.pop("tmp_instructions", None),
            dtype=kwargs.pop("dtype", None) or "float64",
            halo_shape=kwargs.pop("halo_shape", None),
            rank_shape=kwargs.pop("rank_shape", None))

    def __call__(self, queue=None, filter_args=False, **kwargs):
        """-> ``(None, {})`` like ``(evt, outputs)`` of the reference."""
        if self._generic is None:
            raise TypeError("this ElementWiseMap holds no statements")
        return self._generic(queue, filter_args=filter_args, **kwargs)


class Stencil(ElementWiseMap):
    """
    ``Stencil(map_instructions, halo_shape, prefetch_args=[...], **kwargs)``
    (``pystella/stencil.py:36-98``): an element-wise map whose statements read neighbouring
    sites.  The reference stages ``prefetch_args`` in local memory; so does the generated
    kernel here (``codegen.generate_stencil_kernel``: a ring of x planes of the stencil's
    bounding box in shared memory per named array, filled with asynchronous copies while the
    CTA streams along x), arrays not named are read through L1/L2.  ``tile=False`` (or an
    empty ``prefetch_args``) selects the direct kernel for every array.

    Both kernels evaluate the same statements; when the map is idempotent (no array is both
    read and written) the first call times each of them on the caller's arrays (3 launches
    each, CUDA events on the launch stream) and keeps the faster one (``autotune=False``:
    always the shared-memory kernel).  The hand-written stencil kernels are
    :class:`FiniteDifferencer` and :class:`FusedStencilStage`.
    As a ``MapKernel`` class it behaves like :class:`ElementWiseMap`.
    """

    def __init__(self, map_instructions=None, halo_shape=None, **kwargs):
        self.prefetch_args = kwargs.pop("prefetch_args", [])
        kwargs.pop("lsize", None)
        tile = kwargs.pop("tile", True)
        autotune = kwargs.pop("autotune", True)
        if map_instructions is not None and halo_shape is None:
            raise TypeError("Stencil requires halo_shape")
        tiled_kwargs = dict(kwargs)
        super().__init__(map_instructions, halo_shape=halo_shape, **kwargs)
        self._candidates = None
        self.tuning = None               # {"tiled": ms, "direct": ms, "chosen": name}
        if self._generic is not None and tile and self.prefetch_args:
            from pystella_b200.codegen import TiledStencilMap
            try:
                tiled = TiledStencilMap(
                    map_instructions, tiled_kwargs.get("tmp_instructions"),
                    dtype=tiled_kwargs.get("dtype") or "float64", halo_shape=halo_shape,
                    rank_shape=tiled_kwargs.get("rank_shape"),
                    prefetch_args=self.prefetch_args)
            except (NotImplementedError, ValueError):
                return      # tiles larger than shared memory / nothing to stage: keep
                #             the direct kernel
            if autotune and tiled.ksrc.idempotent:
                self._candidates = {"tiled": tiled, "direct": self._generic}
            self._generic = tiled

    def _autotune(self, queue, kwargs):
        """Time both kernels on the caller's arrays; -> the faster one."""
        import torch
        candidates, self._candidates = self._candidates, None
        if not torch.cuda.is_available() or torch.cuda.is_current_stream_capturing():
            return
        stream = queue if queue is not None else torch.cuda.current_stream()
        times = {}
        try:
            for name, kern in candidates.items():
                kern(queue, **kwargs)                       # compile + first launch
                a = torch.cuda.Event(enable_timing=True)
                b = torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(3):
                    kern(queue, **kwargs)
                b.record(stream)
                b.synchronize()
                times[name] = a.elapsed_time(b) / 3
        except (ValueError, TypeError):
            raise                                           # the caller's arguments
        except Exception:  # noqa: BLE001
            return                                          # keep the shared-memory kernel
        best = min(times, key=times.get)
        self._generic = candidates[best]
        self.tuning = dict(times, chosen=best)

    def __call__(self, queue=None, filter_args=False, **kwargs):
        if self._candidates is not None:
            self._autotune(queue, kwargs)
        return super().__call__(queue, filter_args=filter_args, **kwargs)


class StreamingStencil(Stencil):
    """``pystella/stencil.py:101-143``: the streaming-prefetch variant; same semantics."""

    def __init__(self, map_instructions=None, halo_shape=None, **kwargs):
        if len(kwargs.get("prefetch_args", [])) > 1:
            raise NotImplementedError(
                "Streaming codegen can only handle one prefetch array for now")
        super().__init__(map_instructions, halo_shape, **kwargs)


class FusedStencilStage(ElementWiseMap):
    """
    ``MapKernel`` that evaluates ``lap_f`` in-kernel from ``f`` (TMA-staged tile,
    finite-difference stencil of radius ``halo_shape``) instead of reading a
    precomputed ``lap_f`` array, and maintains the face halos of ``f`` itself.
    Requires ``dx=`` (and ``decomp=`` for more than one rank).
    """


# ---------------------------------------------------------------------------
# system recognition
# ---------------------------------------------------------------------------

class ScalarSystem:
    """A set of (f_n, dfdt_n) pairs obeying the scalar-sector equations."""

    def __init__(self, field, nscalars, hubble, a, dV, has_outer_axis, V=None):
        self.V = V                      # list of (coef_expr, powers) or None
        self.field = field              # the DynamicField
        self.f_name = field.name
        self.dot_name = field.dot.name
        self.lap_name = field.lap.name
        self.nscalars = nscalars
        self.hubble = hubble            # scalar expression
        self.a = a                      # scalar expression
        self.dV = dV                    # [n] -> list of (coef_expr, powers)
        self.has_outer_axis = has_outer_axis

    def scalar_names(self):
        names = set()
        from pystella_b200.field import _walk
        exprs = [self.hubble, self.a] + [c for terms in self.dV for c, _ in terms]
        exprs += [c for c, _ in (self.V or [])]
        for e in exprs:
            for node in _walk(e):
                if isinstance(node, (Variable, Field)):
                    names.add(node.name)
        return names


def _unknown_parts(key):
    """-> (Field, outer subscript tuple) of an rhs_dict key."""
    if isinstance(key, Field):
        return key, ()
    if isinstance(key, Subscript) and isinstance(key.aggregate, Field):
        return key.aggregate, tuple(int(i) for i in key.index)
    raise ValueError("rhs_dict keys must be Field instances")


def recognise_scalar_system(rhs_dict):
    """
    Match ``rhs_dict`` against::

        f[n]'    = dfdt[n]
        dfdt[n]' = lap_f[n] + c_d * dfdt[n] + sum_t c_t * prod_m f[m]**p_tm

    with scalar-expression coefficients (ScalarSector.rhs_dict,
    ``pystella/sectors.py:117-131``; the wave equation of
    ``examples/wave_equation.py:49-53`` is the case c_d = 0, no polynomial).
    Returns a :class:`ScalarSystem` or raises ``NotImplementedError`` saying what
    did not match.
    """
    keys = list(rhs_dict.keys())
    dyn = None
    for k in keys:
        fld, _ = _unknown_parts(k)
        if isinstance(fld, DynamicField):
            if dyn is not None and fld.name != dyn.name:
                raise NotImplementedError(
                    "more than one DynamicField in one system: pass sectors "
                    "separately (each gets its own stage kernel)")
            dyn = fld
    if dyn is None:
        raise NotImplementedError(
            "no DynamicField among the unknowns: device steppers implement the "
            "scalar-sector / wave-equation system (f' = dfdt, dfdt' = lap_f + ...)")
    if len(dyn.shape) > 1:
        raise NotImplementedError("DynamicField with more than one outer axis")
    nscalars = dyn.shape[0] if dyn.shape else 1
    has_outer = bool(dyn.shape)

    def sub(n):
        return (n,) if has_outer else ()

    if len(keys) != 2 * nscalars:
        raise NotImplementedError(
            f"expected {2 * nscalars} unknowns (f and dfdt per field), got {len(keys)}")

    hubble = None
    dV = []
    for n in range(nscalars):
        kf = dyn[sub(n)] if has_outer else dyn
        kd = dyn.dot[sub(n)] if has_outer else dyn.dot
        if kf not in rhs_dict or kd not in rhs_dict:
            raise NotImplementedError(f"missing equation for field {n}")
        # f' = dfdt
        pf = to_polynomial(rhs_dict[kf])
        ok = (len(pf) == 1)
        if ok:
            (mono, c), = pf.items()
            ok = (len(mono) == 1 and mono[0][1] == 1 and mono[0][0].name == dyn.dot.name
                  and mono[0][0].subscript == sub(n) and is_constant(c) and c == 1)
        if not ok:
            raise NotImplementedError(
                f"equation for {kf} is not `{kd}` (got {rhs_dict[kf]})")
        # dfdt' = lap + c_d dfdt + poly(f)
        pd = to_polynomial(rhs_dict[kd])
        c_d = 0
        terms = []
        saw_lap = False
        for mono, c in pd.items():
            names = {ref.name for ref, _ in mono}
            if any(any(ref.shift) for ref, _ in mono):
                raise NotImplementedError("shifted (stencil) operand in a stepper RHS")
            if names == {dyn.lap.name}:
                ref, pw = mono[0]
                if not (len(mono) == 1 and pw == 1 and ref.subscript == sub(n)
                        and is_constant(c) and c == 1):
                    raise NotImplementedError(
                        f"Laplacian term of field {n} must be exactly lap_f[{n}]")
                saw_lap = True
            elif names == {dyn.dot.name}:
                ref, pw = mono[0]
                if not (len(mono) == 1 and pw == 1 and ref.subscript == sub(n)):
                    raise NotImplementedError(
                        "only a linear friction term c*dfdt[n] is supported")
                c_d = c
            elif names <= {dyn.name}:
                powers = [0] * nscalars
                for ref, pw in mono:
                    m = ref.subscript[0] if has_outer else 0
                    powers[m] = pw
                terms.append((-1 * c, tuple(powers)))
            else:
                raise NotImplementedError(
                    f"term {mono} of d({kd})/dt is outside the scalar-sector menu")
        if not saw_lap:
            raise NotImplementedError(f"d({kd})/dt has no Laplacian term")
        h_n = -1 * c_d / 2 if not (is_constant(c_d) and c_d == 0) else 0
        if hubble is None:
            hubble = h_n
        elif not _same_scalar(hubble, h_n):
            raise NotImplementedError("fields with different friction coefficients")
        dV.append(terms)
    return ScalarSystem(dyn, nscalars, hubble, 1, dV, has_outer)


def _same_scalar(a, b):
    if is_constant(a) and is_constant(b):
        return a == b
    from pystella_b200.field import _walk
    names = sorted({n.name for e in (a, b) for n in _walk(e)
                    if isinstance(n, (Variable, Field))})
    for seed in (0.731, 1.377):
        ctx = {n: seed * (k + 1.41) for k, n in enumerate(names)}
        if abs(evaluate(a, ctx) - evaluate(b, ctx)) > 1e-13:
            return False
    return True


def scalar_system_from_sector(sector):
    """
    Exact structure of a :class:`~pystella_b200.ScalarSector`: keeps ``a`` and
    ``hubble`` as separate kernel scalars and ``dV/df_n`` as its own polynomial,
    so the kernel evaluates ``(lap - 2*hubble*dfdt) - a**2 * dV`` in the
    reference's association order.
    """
    from pystella_b200.field import diff
    f = sector.f
    n = sector.nscalars
    V = sector.potential(f)
    dV = []
    for fld in range(n):
        poly = to_polynomial(diff(V, f[fld]))
        terms = []
        for mono, c in poly.items():
            powers = [0] * n
            for ref, pw in mono:
                if ref.name != f.name or any(ref.shift):
                    raise NotImplementedError(
                        "potential must be a polynomial in the scalar fields")
                powers[ref.subscript[0] if f.shape else 0] = pw
            terms.append((c, tuple(powers)))
        dV.append(terms)
    Vterms = []
    for mono, c in to_polynomial(V).items():
        powers = [0] * n
        for ref, pw in mono:
            if ref.name != f.name or any(ref.shift):
                raise NotImplementedError(
                    "potential must be a polynomial in the scalar fields")
            powers[ref.subscript[0] if f.shape else 0] = pw
        Vterms.append((c, tuple(powers)))
    return ScalarSystem(f, n, Field("hubble", indices=[]), Field("a", indices=[]),
                        dV, bool(f.shape), V=Vterms)


def _is_quadratic(system):
    """One field with dV/df = c f: the loop-free path of the hand-written stage."""
    return (system.nscalars == 1 and len(system.dV[0]) == 1
            and tuple(system.dV[0][0][1]) == (1,))


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------

def get_name(expr):
    if isinstance(expr, Field):
        return expr.name
    if isinstance(expr, Subscript):
        return get_name(expr.aggregate)
    if isinstance(expr, Variable):
        return expr.name
    if isinstance(expr, str):
        return expr
    raise ValueError(f"cannot name {expr}")


def gen_tmp_name(expr, prefix="_", suffix="_tmp"):
    return prefix + get_name(expr) + suffix


def _scalar_value(x):
    """Host scalar from float / 0-d or 1-element ndarray; None for device tensors."""
    if isinstance(x, torch.Tensor):
        return None
    arr = np.asarray(x, dtype=np.float64)
    if arr.size != 1:
        raise ValueError("scalar arguments must have exactly one element")
    return float(arr.reshape(-1)[0])


def _is_device(x):
    return isinstance(x, torch.Tensor) and _lib.on_device(x)


class Stepper:
    """
    Base class of time steppers (``pystella/step.py:67-170``).

    :arg input: a ``dict`` ``{y: dy/dt}``, a :class:`~pystella_b200.Sector` or a
        list of Sectors.

    Keyword arguments: ``MapKernel`` (:class:`ElementWiseMap` or
    :class:`FusedStencilStage`), ``dt`` (fixes the time step), ``halo_shape``,
    ``rank_shape``, ``dtype``, and for the fused kernel ``dx`` / ``decomp``.
    Unknown keyword arguments of the reference's code generator (``target``,
    ``lsize``, ``args`` ...) are accepted and ignored.
    """

    num_stages = None
    expected_order = None
    num_copies = None

    def __init__(self, input, MapKernel=ElementWiseMap, **kwargs):
        from pystella_b200.sectors import ScalarSector, Sector, TensorPerturbationSector
        self._sectors = None
        if isinstance(input, (Sector, TensorPerturbationSector)):
            self._sectors = [input]
            self.rhs_dict = input.rhs_dict
        elif isinstance(input, list):
            self._sectors = list(input)
            self.rhs_dict = dict(i for s in input for i in s.rhs_dict.items())
        elif isinstance(input, dict):
            self.rhs_dict = input
        else:
            raise TypeError("input must be a dict, a Sector or a list of Sectors")

        for key in self.rhs_dict:
            _unknown_parts(key)  # raises ValueError for non-Field keys

        if not (isinstance(MapKernel, type) and issubclass(MapKernel, ElementWiseMap)):
            raise TypeError("MapKernel must be ElementWiseMap or a subclass")
        self.MapKernel = MapKernel
        self.fused = issubclass(MapKernel, FusedStencilStage)

        self.dt = kwargs.pop("dt", None)
        fixed = kwargs.pop("fixed_parameters", None) or {}
        if self.dt is None and "dt" in fixed:
            self.dt = fixed["dt"]
        self.halo_shape = kwargs.pop("halo_shape", None)
        self.rank_shape = kwargs.pop("rank_shape", None)
        self.dtype = kwargs.pop("dtype", None)
        self.dx = kwargs.pop("dx", None)
        self.decomp = kwargs.pop("decomp", None)
        self.strict_fp = kwargs.pop("strict_fp", False)
        self.jit = kwargs.pop("jit", None)
        self._jit_maps = {}
        #: fused stage also accumulates the energy sums of its input state
        self.fused_energy = kwargs.pop("energy", False)
        self.energy_partials = None
        self.energy_nslots = 0
        self.tune = kwargs.pop("tune", None)
        #: PSB_LAUNCH_* flags of the fused stage kernels and the peer exchange kernel:
        #: ``tune={"pdl": bool}`` -- programmatic dependent launch, the next kernel's CTAs
        #: set up while the previous kernel drains (include/pystella_b200.h)
        self.launch_flags = _lib.LAUNCH_PDL if (self.tune or {}).get("pdl", DEFAULT_PDL) else 0
        self._ignored_kwargs = kwargs  # target=, lsize=, args=, options= ...

        if self.fused and self.dx is None:
            raise ValueError("FusedStencilStage needs the grid spacing: pass dx=(...)")
        if isinstance(self.halo_shape, (tuple, list)):
            if len(set(self.halo_shape)) != 1:
                raise NotImplementedError("steppers need an isotropic halo_shape")
            self.halo_shape = self.halo_shape[0]

        self.num_unknowns = len(self.rhs_dict)
        self.dof_names = {get_name(key) for key in self.rhs_dict}
        self._systems = None
        self._ScalarSector = ScalarSector
        self._TensorSector = TensorPerturbationSector

    # -- device systems -------------------------------------------------------
    def _build_systems(self):
        def general(rhs_dict, potential, reason):
            # EXPERIMENTAL: outside the polynomial menu the fused pipeline is compiled at
            # run time with a generated right-hand side (pystella_b200/fused_jit.py)
            if not (self.fused and jit_enabled(self.jit)):
                raise reason
            from pystella_b200.fused_jit import recognise_general_system
            return recognise_general_system(rhs_dict, potential)

        systems = []
        if self._sectors is not None:
            for s in self._sectors:
                if isinstance(s, self._ScalarSector):
                    try:
                        system = scalar_system_from_sector(s)
                    except NotImplementedError as exc:
                        systems.append(general(s.rhs_dict, s.potential(s.f), exc))
                        continue
                    # The hand-written stage evaluates polynomial potentials from run-time
                    # tables (loops over terms and powers per site).  Only the single-field
                    # quadratic potential has a loop-free path there; everything else is
                    # faster through the generated functor (measured, 2 fields 512^3 fp32:
                    # 1.815 -> 1.474 ms per stage, 0.73 -> 0.90 of the roofline), so fused
                    # steppers prefer it whenever run-time compilation is available.
                    # tune={"tables": True} keeps the table-driven kernel.
                    tune = self.tune or {}
                    if (self.fused and jit_enabled(self.jit) and not _is_quadratic(system)
                            and not tune.get("tables", False)
                            and tune.get("variant", 0) == 0):   # explicit variants: tables
                        try:
                            tables = system
                            system = general(s.rhs_dict, s.potential(s.f), None)
                            # layouts the TMA pipeline cannot describe (odd radii, odd
                            # rows, unaligned views) are served by the table-driven kernel
                            system.tables = tables
                        except NotImplementedError:
                            pass
                    systems.append(system)
                elif isinstance(s, self._TensorSector):
                    from pystella_b200.gw import tensor_system_from_sector
                    systems.append(tensor_system_from_sector(s))
                else:
                    try:
                        systems.append(recognise_scalar_system(s.rhs_dict))
                    except NotImplementedError as exc:
                        systems.append(general(s.rhs_dict, None, exc))
        else:
            try:
                systems.append(recognise_scalar_system(self.rhs_dict))
            except NotImplementedError as exc:
                systems.append(general(self.rhs_dict, None, exc))
        return systems

    @property
    def systems(self):
        if self._systems is None:
            self._systems = self._build_systems()
        return self._systems

    def _resolve_dt(self, kwargs):
        dt = self.dt if self.dt is not None else kwargs.get("dt")
        if dt is None:
            raise ValueError("dt was not fixed at construction; pass dt=...")
        return float(np.asarray(dt).reshape(-1)[0])

    def _on_device(self, kwargs):
        arrays = [kwargs[n] for n in self.dof_names if n in kwargs]
        if not arrays:
            missing = sorted(self.dof_names)
            raise TypeError(f"missing array arguments {missing}")
        if all(isinstance(x, np.ndarray) for x in arrays):
            return False
        if all(isinstance(x, torch.Tensor) for x in arrays):
            for n in self.dof_names:
                if not _lib.on_device(kwargs[n]):
                    raise PsbError(
                        f"'{n}' lives on {kwargs[n].device}: the stepper kernels run on "
                        "the GPU only (no CPU fallback)")
            return True
        raise TypeError("unknown arrays must be all torch.Tensor (device) or all "
                        "numpy.ndarray (host scalar ODEs)")

    def __call__(self, stage, queue=None, **kwargs):
        raise NotImplementedError


# ---------------------------------------------------------------------------
# low-storage RK (pystella/step.py:441-528)
# ---------------------------------------------------------------------------

class LowStorageRKStepper(Stepper):
    """
    2N-storage Runge-Kutta: per unknown ``u`` with twin ``k``
    ``k <- A[s]*k + dt*rhs ; u <- u + B[s]*k`` with all right-hand sides evaluated
    from the old state (``pystella/step.py:473-491``).  Twins (``_<name>_tmp``) are
    allocated zero-filled on the first call and kept in :attr:`tmp_arrays`
    (``step.py:493-528``).
    """

    _A = []
    _B = []
    _C = []

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.tmp_arrays = {}
        self._twins = {}        # fused: second padded buffer per stencilled unknown
        self._fresh = {}        # fused: per system, whose halos the stage kernel last wrote
        self._streams = None

    def get_tmp_arrays_like(self, **kwargs):
        """Zero-filled twins matching the unknown arrays passed by keyword."""
        tmp_arrays = {}
        for name in self.dof_names:
            f = kwargs[name]
            tmp_name = gen_tmp_name(name)
            if isinstance(f, torch.Tensor):
                tmp_arrays[tmp_name] = torch.zeros_like(f)
            elif isinstance(f, np.ndarray):
                tmp_arrays[tmp_name] = np.zeros_like(f)
            else:
                raise ValueError(f"Could not generate tmp array for {f}"
                                 f"of type {type(f)}")
        return tmp_arrays

    def reset(self):
        """Forget halo freshness.  Not needed after in-place edits or when passing a new
        tensor (both are detected, see :func:`pystella_b200.fused.halos_are_fresh`); kept for
        writes through raw pointers the stepper cannot see.  Also re-arms the rows of the
        in-kernel stage tail (needed only after a kernel was aborted mid-run)."""
        self._fresh = {}
        from pystella_b200 import _lib
        for rows in (self.__dict__.get("energy_partials"), self.__dict__.get("_tail_rows")):
            if rows is not None:
                rows.view(torch.int64).fill_(_lib.TAIL_SENTINEL)

    def __call__(self, stage, *, queue=None, **kwargs):
        if not 0 <= stage < self.num_stages:
            raise IndexError(f"stage {stage} out of range for {type(self).__name__}")
        kwargs.pop("filter_args", None)
        kwargs.pop("allocator", None)
        if len(self.tmp_arrays) == 0:
            self.tmp_arrays = self.get_tmp_arrays_like(**kwargs)
        if self._on_device(kwargs):
            return self._device_stage(stage, queue, kwargs)
        return self._host_stage(stage, kwargs)

    # -- host scalar ODEs -------------------------------------------------------
    #: host evaluation exists for the scalar ODEs the reference itself integrates on the
    #: host (Expansion through loopy's C target, pystella/expansion.py:91-99; the steppers'
    #: own order tests, test/test_step.py): NOT a CPU fallback for lattices
    HOST_MAX_ELEMENTS = 4096

    def _host_stage(self, stage, kwargs):
        for name in self.dof_names:
            if np.asarray(kwargs[name]).size > self.HOST_MAX_ELEMENTS:
                raise PsbError(
                    f"'{name}' is a NumPy array of {np.asarray(kwargs[name]).size} elements: "
                    "lattice fields must be torch tensors on the GPU (the host path only "
                    "integrates small ODE systems such as the FLRW background; there is no "
                    "CPU fallback)")
        ctx = dict(kwargs)
        ctx.update(self.tmp_arrays)
        if self.dt is not None:
            ctx["dt"] = self.dt
        dt = ctx["dt"]
        A, B = self._A[stage], self._B[stage]
        rhs = [np.array(evaluate(v, ctx), copy=True) for v in self.rhs_dict.values()]
        for key, r in zip(self.rhs_dict.keys(), rhs):
            fld, sub = _unknown_parts(key)
            if isinstance(fld.child, Subscript):
                sub = sub + tuple(int(i) for i in fld.child.index)
            u = kwargs[fld.name]
            k = self.tmp_arrays[gen_tmp_name(fld.name)]
            idx = sub if sub else Ellipsis
            k[idx] = A * k[idx] + dt * r
            u[idx] = u[idx] + B * k[idx]
        return None

    # -- device ---------------------------------------------------------------
    def _jit_stage(self, stage, queue, kwargs):
        """
        EXPERIMENTAL: the reference's own construction (``step.py:473-491``) -- right-hand
        sides into temporaries, then ``k = A k + dt rhs ; u = u + B k`` per unknown --
        handed to the run-time code generator instead of loopy.
        """
        from pystella_b200.codegen import GenericMap
        if stage not in self._jit_maps:
            maps, tmps = self._jit_statements(stage)
            some = kwargs[next(iter(self.dof_names))]
            dtype = str(some.dtype).replace("torch.", "")
            self._jit_maps[stage] = GenericMap(maps, tmps, dtype=dtype,
                                               halo_shape=self.halo_shape,
                                               rank_shape=self.rank_shape)
        call = dict(kwargs)
        call.update(self.tmp_arrays)
        call["dt"] = self._resolve_dt(kwargs)
        self._jit_maps[stage](queue, **call)
        return None

    def _jit_statements(self, stage):
        """-> (map_instructions, tmp_instructions) of one stage; step.py:473-491."""
        from pystella_b200.field import var
        tmps, maps = [], []
        for i, (key, val) in enumerate(self.rhs_dict.items()):
            tmps.append((var(gen_tmp_name(key, suffix=f"_rhs_{i}")), val))
        for i, key in enumerate(self.rhs_dict):
            fld, sub = _unknown_parts(key)
            twin = Field(gen_tmp_name(fld), offset=fld.offset, shape=fld.shape,
                         indices=fld.indices)
            k = twin[sub] if sub else twin
            rhs = var(gen_tmp_name(key, suffix=f"_rhs_{i}"))
            maps.append((k, self._A[stage] * k + var("dt") * rhs))
            maps.append((key, key + self._B[stage] * k))
        return maps, tmps

    def _device_stage(self, stage, queue, kwargs):
        try:
            systems = self.systems
        except NotImplementedError:
            if not jit_enabled(self.jit) or self.fused:
                raise
            return self._jit_stage(stage, queue, kwargs)
        dt = self._resolve_dt(kwargs)
        A, B = self._A[stage], self._B[stage]
        self._stage_index = stage
        self._call_serial = self.__dict__.get("_call_serial", 0) + 1
        for n, system in enumerate(systems):
            # the last kernel of the stage carries the in-kernel tail (inter-GPU barrier,
            # background stage): pystella_b200/csrc/stage_tail.cuh
            self._last_of_stage = (n == len(systems) - 1)
            hook = self.__dict__.get("_timing_hook")
            if hook is not None:               # per-kernel CUDA events (bench.py roofline)
                cur = queue if queue is not None else torch.cuda.current_stream()
                ev0 = torch.cuda.Event(enable_timing=True)
                ev1 = torch.cuda.Event(enable_timing=True)
                ev0.record(cur)
            system_stage = getattr(system, "run_stage", None)
            if system_stage is not None:       # tensor sector etc.
                system_stage(self, stage, queue, kwargs, dt, A, B)
            else:
                self._scalar_stage(system, stage, queue, kwargs, dt, A, B)
            if hook is not None:
                ev1.record(cur)
                hook.append((stage, n, ev0, ev1))
        return None

    # -- in-kernel stage tail ---------------------------------------------------
    def attach_expansion(self, expand):
        """Let the last kernel of every stage advance ``expand`` (a
        :class:`~pystella_b200.DeviceExpansion`) in its last CTA: no launch between stages.
        ``expand.step(stage, stepper, dt)`` then returns immediately."""
        self._expansion = expand

    def _make_tail(self, *, fold, dt, nrows=0, device=None):
        """psb_stage_tail of the launch being issued (None when there is nothing to do).
        ``nrows``: CTAs of a launch WITHOUT energy partials (it publishes marker rows)."""
        from pystella_b200 import _lib
        mode = 0
        if fold:
            mode |= _lib.TAIL_FOLD
        peer_mode = getattr(self, "_halo_mode", None) == "peer"
        expand = getattr(self, "_expansion", None)
        last = getattr(self, "_last_of_stage", True)
        # Multi-rank: the inter-GPU exchange stays a separate tiny kernel after the stage
        # (psb_expansion_stage_peer).  Doing it in the watcher CTA needs a system-scope fence
        # in every CTA that stored halo planes into a neighbour's memory, and that fence keeps
        # the CTA's SM slot for ~5 us: measured at 8 GPUs +35 us per stage at 512^3 (a third of
        # the CTAs are boundary CTAs) and +68 us at 1024^3.  tune={"peer_tail": True} selects
        # the in-kernel variant.
        peer_tail = bool((self.tune or {}).get("peer_tail", False))
        if last and peer_mode and peer_tail:
            mode |= _lib.TAIL_PEER
        if last and expand is not None and self.fused_energy and (not peer_mode or peer_tail):
            mode |= _lib.TAIL_EXPAND
        if mode == 0:
            return None
        if getattr(self, "_tail_sums", None) is None:
            #: folded energy sums {sum dfdt_n^2.., sum V, sum -f_n lap f_n..} of the state
            #: that entered the last stage (device, float64[16])
            self._tail_sums = torch.zeros(16, dtype=torch.float64, device=device)
            self.energy_sums = self._tail_sums
        tail = _lib.StageTail()
        tail.mode = mode
        tail.sums = self._tail_sums.data_ptr()
        if not fold and nrows > 0:
            rows = getattr(self, "_tail_rows", None)
            if rows is None or rows.shape[0] < nrows:
                if rows is not None:
                    torch.cuda.synchronize()
                rows = _kernels.sentinel_rows(nrows, device)
                self._tail_rows = rows
            tail.rows, tail.nrows = rows.data_ptr(), rows.shape[0]
        if self.fused_energy:
            tail.nscalars = self.energy_nvals // 2 if getattr(self, "energy_nvals", 0) else 0
        if mode & _lib.TAIL_EXPAND:
            stage = self._stage_index
            tail.nscalars = expand.nscalars
            tail.grid_size, tail.mpl, tail.dt = expand.grid_size, expand.mpl, float(dt)
            tail.A, tail.B = expand._A[stage], expand._B[stage]
            tail.state = expand.state.data_ptr()
            tail.energies = expand.energies.data_ptr()
        if mode & _lib.TAIL_PEER:
            pa = self._peer.args
            tail.peers.nranks, tail.peers.rank = pa.nranks, pa.rank
            tail.peers.peer_slots, tail.peers.peer_flags = pa.peer_slots, pa.peer_flags
            tail.peers.epoch = pa.epoch
        return tail

    def _tail_issued(self, tail):
        """Book-keeping after a launch that carried ``tail``."""
        from pystella_b200 import _lib
        mode = tail.mode if tail is not None else 0
        if mode & _lib.TAIL_FOLD:
            self._sums_folded = True
        if mode & _lib.TAIL_PEER:
            self._peer_barrier_pending = False
        if mode & _lib.TAIL_EXPAND:
            self._expanded_stage = self._stage_index

    def _scalars(self, system, kwargs):
        out = {}
        for which, expr in (("a", system.a), ("hubble", system.hubble)):
            dev = None
            if isinstance(expr, Field) and expr.name in kwargs \
                    and isinstance(kwargs[expr.name], torch.Tensor):
                dev = kwargs[expr.name]
                if not (_lib.on_device(dev) and dev.dtype == torch.float64
                        and dev.numel() == 1):
                    raise ValueError(f"device scalar '{expr.name}' must be one float64")
                out[which] = (1.0, dev)
                continue
            ctx = {n: _scalar_value(kwargs[n]) for n in system.scalar_names()
                   if n in kwargs and not isinstance(kwargs[n], torch.Tensor)}
            try:
                val = expr if is_constant(expr) else evaluate(expr, ctx)
            except KeyError as exc:
                # the reference's kernel fails on a missing argument as well
                raise TypeError(f"missing scalar argument of the stage kernel: {exc}") \
                    from None
            out[which] = (float(val), None)
        ctx = {n: _scalar_value(kwargs[n]) for n in system.scalar_names()
               if n in kwargs and not isinstance(kwargs[n], torch.Tensor)}
        dV = [[(c if is_constant(c) else float(evaluate(c, ctx)), p) for c, p in terms]
              for terms in system.dV]
        self._V_values = None
        if system.V is not None:
            self._V_values = [(c if is_constant(c) else float(evaluate(c, ctx)), p)
                              for c, p in system.V]
        return out, dV

    def _generated_instance_applies(self, system, f, dfdt, kf, kd, h):
        """The generated instance of a polynomial system is the TMA pipeline compiled with
        the potential inlined: it exists for the layouts that pipeline serves.  Asked of the
        library itself (dry-run launch plan), once per layout."""
        key = (system.f_name, tuple(f.shape), f.dtype, h,
               tuple(t.data_ptr() % 16 for t in (f, dfdt, kf, kd)))
        cache = self.__dict__.setdefault("_generated_ok", {})
        if key not in cache:
            tune = self.tune or {}
            cache[key] = bool(_kernels.scalar_stage(
                fused=True, probe_plan=True, f_in=f, f_out=kf, dfdt=dfdt, kf=kf, kdfdt=kd,
                h=h, dx=self.dx, dt=0.0, A=0.0, B=0.0, dV_terms=None,
                strict=self.strict_fp, xchunk=tune.get("xchunk", 0),
                variant=tune.get("variant", 0)))
        return cache[key]

    def _scalar_stage(self, system, stage, queue, kwargs, dt, A, B):
        try:
            f = kwargs[system.f_name]
            dfdt = kwargs[system.dot_name]
        except KeyError as e:
            raise TypeError(f"missing array argument {e}") from None
        kf = self.tmp_arrays[gen_tmp_name(system.f_name)]
        kd = self.tmp_arrays[gen_tmp_name(system.dot_name)]
        h = self.halo_shape
        if h is None:
            raise ValueError("pass halo_shape= when constructing a device stepper")
        if (self.fused and getattr(system, "generated", None) is not None
                and not self._generated_instance_applies(system, f, dfdt, kf, kd, h)):
            if getattr(system, "tables", None) is None:
                raise NotImplementedError(
                    "the fused stage of a system with a run-time generated right-hand side "
                    "is an instance of the TMA pipeline: it needs an even halo_shape, an even "
                    "number of z sites and 16-byte aligned rows and arrays.  For this layout "
                    "use the element-wise stepper (default MapKernel) with FiniteDifferencer")
            system = system.tables
        scal, dV = self._scalars(system, kwargs)
        common = dict(dfdt=dfdt, kf=kf, kdfdt=kd, h=h, dt=dt, A=A, B=B,
                      a=scal["a"][0], hubble=scal["hubble"][0],
                      a_dev=scal["a"][1], hubble_dev=scal["hubble"][1],
                      dV_terms=dV, queue=queue, strict=self.strict_fp)
        generated = getattr(system, "generated", None)
        if generated is not None:
            values = []
            for name in system.user_scalars:
                if name not in kwargs:
                    raise TypeError(f"missing scalar argument '{name}'")
                value = _scalar_value(kwargs[name])
                if value is None:
                    raise TypeError(f"scalar '{name}' of a generated right-hand side must "
                                    "be a host value (only `a` and `hubble` may live on "
                                    "the device)")
                values.append(value)
            common["generated"] = (generated, values)
        if not self.fused:
            lap = kwargs.get(system.lap_name)
            if lap is None:
                raise TypeError(
                    f"missing array argument '{system.lap_name}' (the element-wise "
                    "stage reads a precomputed Laplacian; use "
                    "MapKernel=FusedStencilStage to compute it in-kernel)")
            _kernels.scalar_stage(fused=False, f_in=f, f_out=f, lap_f=lap, **common)
            return
        from pystella_b200.fused import fused_scalar_stage
        fused_scalar_stage(self, system, f, common, queue)


class LowStorageRK54(LowStorageRKStepper):
    """
    Five-stage, fourth-order 2N-storage scheme of Carpenter & Kennedy (NASA TM
    109112, 1994); coefficients as in ``pystella/step.py:543-557``.
    """

    num_stages = 5
    expected_order = 4

    _A = [
        0,
        -567301805773 / 1357537059087,
        -2404267990393 / 2016746695238,
        -3550918686646 / 2091501179385,
        -1275806237668 / 842570457699,
    ]

    _B = [
        1432997174477 / 9575080441755,
        5161836677717 / 13612068292357,
        1720146321549 / 2090206949498,
        3134564353537 / 4481467310338,
        2277821191437 / 14882151754819,
    ]

    _C = [
        0,
        1432997174477 / 9575080441755,
        2526269341429 / 6820363962896,
        2006345519317 / 3224310063776,
        2802321613138 / 2924317926251,
    ]


def _load_long_tables():
    """
    The 12/13/14-stage schemes (``pystella/step.py:568-754``) are pure coefficient
    tables; they are kept as data in ``lsrk_tables.json`` next to this module.
    """
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lsrk_tables.json")
    with open(path) as fh:
        return json.load(fh)


_LONG = _load_long_tables()


def _make_lsrk(name, doc):
    tab = _LONG[name]
    return type(name, (LowStorageRKStepper,), {
        "__doc__": doc, "num_stages": tab["num_stages"],
        "expected_order": tab["expected_order"],
        "_A": tab["A"], "_B": tab["B"], "_C": tab["C"]})


LowStorageRK144 = _make_lsrk(
    "LowStorageRK144", "14-stage, fourth-order low-storage scheme optimised for "
    "elliptic stability regions (Niegemann, Diehl & Busch 2012); step.py:568-628.")
LowStorageRK134 = _make_lsrk(
    "LowStorageRK134", "13-stage, fourth-order low-storage scheme optimised for "
    "circular stability regions (Niegemann et al. 2012); step.py:631-690.")
LowStorageRK124 = _make_lsrk(
    "LowStorageRK124", "12-stage, fourth-order low-storage scheme optimised for "
    "inviscid problems (Niegemann et al. 2012); step.py:693-754.")


class LowStorageRK3Williamson(LowStorageRKStepper):
    """Williamson (1980) three-stage third-order scheme; step.py:757-773."""
    num_stages = 3
    expected_order = 3
    _A = [0, -5/9, -153/128]
    _B = [1/3, 15/16, 8/15]
    _C = [0, 4/9, 15/32]


class LowStorageRK3Inhomogeneous(LowStorageRKStepper):
    """Three-stage third-order scheme; step.py:776-788."""
    num_stages = 3
    expected_order = 3
    _A = [0, -17/32, -32/27]
    _B = [1/4, 8/9, 3/4]
    _C = [0, 15/32, 4/9]


class LowStorageRK3Symmetric(LowStorageRKStepper):
    """step.py:791-800 (possible order reduction)."""
    num_stages = 3
    expected_order = 3
    _A = [0, -2/3, -1]
    _B = [1/3, 1, 1/2]
    _C = [0, 1/3, 2/3]


class LowStorageRK3PredictorCorrector(LowStorageRKStepper):
    """step.py:803-812 (possible order reduction)."""
    num_stages = 3
    expected_order = 3
    _A = [0, -1/4, -4/3]
    _B = [1/2, 2/3, 1/2]
    _C = [0, 1/2, 1]


def _ssp3_coefficients():
    # closed forms of step.py:815-833 in terms of the free parameter c2
    c2 = .924574
    z1 = np.sqrt(36 * c2**4 + 36 * c2**3 - 135 * c2**2 + 84 * c2 - 12)
    z2 = 2 * c2**2 + c2 - 2
    z3 = 12 * c2**4 - 18 * c2**3 + 18 * c2**2 - 11 * c2 + 2
    z4 = 36 * c2**4 - 36 * c2**3 + 13 * c2**2 - 8 * c2 + 4
    z5 = 69 * c2**3 - 62 * c2**2 + 28 * c2 - 8
    z6 = 34 * c2**4 - 46 * c2**3 + 34 * c2**2 - 13 * c2 + 2
    w = 3 * z2 - z1
    B1 = c2
    B2 = (12 * c2 * (c2 - 1) * w - w**2) / (144 * c2 * (3 * c2 - 2) * (c2 - 1)**2)
    B3 = - 24 * (3 * c2 - 2) * (c2 - 1)**2 / (w**2 - 12 * c2 * (c2 - 1) * w)
    A2 = ((- z1 * (6 * c2**2 - 4 * c2 + 1) + 3 * z3)
          / ((2 * c2 + 1) * z1 - 3 * (c2 + 2) * (2 * c2 - 1)**2))
    A3 = ((- z4 * z1 + 108 * (2 * c2 - 1) * c2**5 - 3 * (2 * c2 - 1) * z5)
          / (24 * z1 * c2 * (c2 - 1)**4 + 72 * c2 * z6 + 72 * c2**6 * (2 * c2 - 13)))
    return [0, A2, A3], [B1, B2, B3], [0, B1, B1 + B2 * (A2 + 1)]


class LowStorageRK3SSP(LowStorageRKStepper):
    """Strong-stability-preserving three-stage third-order scheme; step.py:836-846."""
    num_stages = 3
    expected_order = 3
    _A, _B, _C = _ssp3_coefficients()


# ---------------------------------------------------------------------------
# classical RK with storage copies (pystella/step.py:173-411)
# ---------------------------------------------------------------------------

class RungeKuttaStepper(Stepper):
    """
    Classical explicit RK operating on ``num_copies`` copies of each unknown,
    stored along a leading axis (``pystella/step.py:173-239``).  The right-hand
    side reads copy 0 at stage 0 and copy 1 afterwards (``step.py:231``).
    Subclasses provide :meth:`step_statements`: ``{copy: [(coef, src_copy), ...],
    rhs_coef}`` per stage.

    Host NumPy arrays (scalar ODEs, as the reference's own test exercises them on its C
    target, ``test/test_step.py:42-99``) or, for the scalar-sector / wave-equation menu,
    device arrays with the copies axis first through ``psb_scalar_rk_stage``.
    """

    def step_statements(self, stage):
        """-> list of ``(dst_copy, [(coef, src_copy), ...], dt_coef)`` in order."""
        raise NotImplementedError

    def __call__(self, stage, queue=None, **kwargs):
        if not 0 <= stage < self.num_stages:
            raise IndexError(f"stage {stage} out of range for {type(self).__name__}")
        kwargs.pop("filter_args", None)
        if self._on_device(kwargs):
            try:
                return self._device_rk_stage(stage, queue, kwargs)
            except NotImplementedError:
                if not jit_enabled(self.jit) or self.fused:
                    raise
                return self._jit_rk_stage(stage, queue, kwargs)
        ctx = dict(kwargs)
        if self.dt is not None:
            ctx["dt"] = self.dt
        dt = ctx["dt"]
        q = 0 if stage == 0 else 1
        names = self.dof_names
        rctx = dict(ctx)
        for n in names:
            rctx[n] = kwargs[n][q]
        rhs = [np.array(evaluate(v, rctx), copy=True) for v in self.rhs_dict.values()]
        for key, r in zip(self.rhs_dict.keys(), rhs):
            fld, sub = _unknown_parts(key)
            arr = kwargs[fld.name]
            new = {}
            for dst, combo, c_dt in self.step_statements(stage):
                val = sum(c * arr[(src,) + sub] for c, src in combo) + c_dt * dt * r
                new[dst] = val
            for dst, val in new.items():
                arr[(dst,) + sub] = val
        return None


def _jit_rk_stage(self, stage, queue, kwargs):
    """
    EXPERIMENTAL: the reference's construction (``step.py:181-239``) for any ``rhs_dict``:
    right-hand sides (reading copy ``q``) into temporaries, then the stage's statements on
    the copies axis, handed to the run-time code generator.
    """
    from pystella_b200.codegen import GenericMap
    if stage not in self._jit_maps:
        maps, tmps = self._jit_statements(stage)
        some = kwargs[next(iter(self.dof_names))]
        dtype = str(some.dtype).replace("torch.", "")
        self._jit_maps[stage] = GenericMap(maps, tmps, dtype=dtype,
                                           halo_shape=self.halo_shape,
                                           rank_shape=self.rank_shape)
    call = dict(kwargs)
    call["dt"] = self._resolve_dt(kwargs)
    self._jit_maps[stage](queue, **call)
    return None


def _jit_rk_statements(self, stage):
    """-> (map_instructions, tmp_instructions) of one classical stage; step.py:181-239."""
    from pystella_b200.field import prepend_subscript, var
    q = 0 if stage == 0 else 1
    tmps, maps = [], []
    for i, (key, val) in enumerate(self.rhs_dict.items()):
        tmps.append((var(gen_tmp_name(key, suffix=f"_rhs_{i}")),
                     prepend_subscript(val, (q,))))
    for i, key in enumerate(self.rhs_dict):
        rhs = var(gen_tmp_name(key, suffix=f"_rhs_{i}"))
        for dst, combo, c_dt in self.step_statements(stage):
            expr = 0
            for c, src in combo:
                expr = expr + c * prepend_subscript(key, (src,))
            expr = expr + c_dt * var("dt") * rhs
            maps.append((prepend_subscript(key, (dst,)), expr))
    return maps, tmps


def _device_rk_stage(self, stage, queue, kwargs):
    """Element-wise device stage for the scalar-sector menu (``psb_scalar_rk_stage``):
    arrays carry the copies axis first, ``lap_f`` is computed by the caller from
    copy ``q = 0 if stage == 0 else 1`` (pystella/step.py:181-187, 231)."""
    import ctypes as C
    from pystella_b200 import _lib
    if self.fused:
        raise NotImplementedError("FusedStencilStage with classical RK: use LowStorageRK*")
    if len(self.systems) != 1 or hasattr(self.systems[0], "run_stage"):
        raise NotImplementedError("device classical RK: a single scalar system")
    system = self.systems[0]
    dt = self._resolve_dt(kwargs)
    f, dfdt = kwargs[system.f_name], kwargs[system.dot_name]
    lap = kwargs.get(system.lap_name)
    if lap is None:
        raise TypeError(f"missing array argument '{system.lap_name}'")
    h = self.halo_shape
    if h is None:
        raise ValueError("pass halo_shape= when constructing a device stepper")
    n = system.nscalars
    want = 5 if system.has_outer_axis else 4
    if f.dim() != want or f.shape[0] != self.num_copies or dfdt.shape != f.shape:
        raise ValueError(f"f / dfdt must have shape ({self.num_copies}, "
                         f"{'n, ' if system.has_outer_axis else ''}Nx+2h, Ny+2h, Nz+2h)")
    for name, t in ((system.f_name, f), (system.dot_name, dfdt), (system.lap_name, lap)):
        _lib.require_cuda_tensor(t, name)
    scal, dV = self._scalars(system, kwargs)
    if scal["a"][1] is not None or scal["hubble"][1] is not None:
        raise NotImplementedError("device scalars with classical RK")
    Nx, Ny, Nz = (x - 2 * h for x in f.shape[-3:])
    args = _lib.RKStageArgs()
    args.dtype = _lib.torch_dtype_code(f)
    args.nscalars, args.h, args.ncopies = n, h, self.num_copies
    args.Nx, args.Ny, args.Nz = Nx, Ny, Nz
    args.f, args.dfdt, args.lap_f = f.data_ptr(), dfdt.data_ptr(), lap.data_ptr()
    args.q_read = 0 if stage == 0 else 1
    stmts = self.step_statements(stage)
    args.nstatements = len(stmts)
    for s, (dst, combo, c_dt) in enumerate(stmts):
        args.dst[s] = dst
        args.ncomb[s] = len(combo)
        args.dt_coef[s] = c_dt
        for m, (c, src) in enumerate(combo):
            args.coef[s][m] = c
            args.src[s][m] = src
    args.dt, args.a, args.hubble = dt, scal["a"][0], scal["hubble"][0]
    for i in range(n):
        args.nterms[i] = len(dV[i])
        for t, (c, powers) in enumerate(dV[i]):
            args.pcoef[i][t] = float(c)
            for m in range(n):
                args.power[i][t][m] = int(powers[m])
    lib = _lib.load(self.strict_fp)
    rc = lib.psb_scalar_rk_stage(C.byref(args), _lib.stream_ptr(queue))
    _lib.check(lib, rc, "psb_scalar_rk_stage")
    return None


RungeKuttaStepper._device_rk_stage = _device_rk_stage
RungeKuttaStepper._jit_rk_stage = _jit_rk_stage
RungeKuttaStepper._jit_statements = _jit_rk_statements
RungeKuttaStepper._scalars = LowStorageRKStepper._scalars


class RungeKutta4(RungeKuttaStepper):
    """Classical four-stage fourth-order method, three copies; step.py:242-265."""
    num_stages, expected_order, num_copies = 4, 4, 3

    def step_statements(self, stage):
        return [
            [(1, [(1, 0)], 1/2), (2, [(1, 0)], 1/6)],
            [(1, [(1, 0)], 1/2), (2, [(1, 2)], 1/3)],
            [(1, [(1, 0)], 1), (2, [(1, 2)], 1/3)],
            [(0, [(1, 2)], 1/6)],
        ][stage]


class RungeKutta3Heun(RungeKuttaStepper):
    """Heun's third-order method; step.py:268-285."""
    num_stages, expected_order, num_copies = 3, 3, 3

    def step_statements(self, stage):
        return [
            [(1, [(1, 0)], 1/3), (2, [(1, 0)], 1/4)],
            [(1, [(1, 0)], 2/3)],
            [(0, [(1, 2)], 3/4)],
        ][stage]


class RungeKutta3Nystrom(RungeKuttaStepper):
    """Nystrom's third-order method; step.py:288-307."""
    num_stages, expected_order, num_copies = 3, 3, 3

    def step_statements(self, stage):
        return [
            [(1, [(1, 0)], 2/3), (2, [(1, 0)], 2/8)],
            [(1, [(1, 0)], 2/3), (2, [(1, 2)], 3/8)],
            [(0, [(1, 2)], 3/8)],
        ][stage]


class RungeKutta3Ralston(RungeKuttaStepper):
    """Ralston's third-order method; step.py:310-329."""
    num_stages, expected_order, num_copies = 3, 3, 3

    def step_statements(self, stage):
        return [
            [(1, [(1, 0)], 1/2), (2, [(1, 0)], 2/9)],
            [(1, [(1, 0)], 3/4), (2, [(1, 2)], 1/3)],
            [(0, [(1, 2)], 4/9)],
        ][stage]


class RungeKutta3SSP(RungeKuttaStepper):
    """Three-stage third-order SSP method, two copies; step.py:332-350."""
    num_stages, expected_order, num_copies = 3, 3, 2

    def step_statements(self, stage):
        return [
            [(1, [(1, 0)], 1)],
            [(1, [(3/4, 0), (1/4, 1)], 1/4)],
            [(0, [(1/3, 0), (2/3, 1)], 2/3)],
        ][stage]


class RungeKutta2Midpoint(RungeKuttaStepper):
    """Midpoint method; step.py:353-370."""
    num_stages, expected_order, num_copies = 2, 2, 2

    def step_statements(self, stage):
        return [
            [(1, [(1, 0)], 1/2)],
            [(0, [(1, 0)], 1)],
        ][stage]


class RungeKutta2Heun(RungeKuttaStepper):
    """Heun's second-order method (possible order reduction); step.py:374-387."""
    num_stages, expected_order, num_copies = 2, 2, 2

    def step_statements(self, stage):
        return [
            [(1, [(1, 0)], 1), (0, [(1, 0)], 1/2)],
            [(0, [(1, 0)], 1/2)],
        ][stage]


class RungeKutta2Ralston(RungeKuttaStepper):
    """Ralston's second-order method; step.py:390-408."""
    num_stages, expected_order, num_copies = 2, 2, 2

    def step_statements(self, stage):
        return [
            [(1, [(1, 0)], 2/3), (0, [(1, 0)], 1/4)],
            [(0, [(1, 0)], 3/4)],
        ][stage]


all_steppers = [RungeKutta4, RungeKutta3SSP, RungeKutta3Heun, RungeKutta3Nystrom,
                RungeKutta3Ralston, RungeKutta2Midpoint,
                RungeKutta2Ralston, LowStorageRK54, LowStorageRK144,
                LowStorageRK3Williamson, LowStorageRK3Inhomogeneous,
                LowStorageRK3SSP]
