"""
EXPERIMENTAL (row f-2 of the scope table): arbitrary right-hand sides inside the fused TMA
pipeline.

pystella accepts any ``rhs_dict`` (``pystella/step.py:473-491``); the hand-written fused
stage (``csrc/fused_stage_ws.cuh``) evaluates the scalar-sector form with a polynomial
potential.  For second-order systems outside that menu -- non-polynomial potentials
(``cos``, ``tanh``, ``exp`` ...), couplings to ``dfdt``, extra scalars --

    f[n]'    = F_n(f, dfdt, lap_f; scalars)
    dfdt[n]' = G_n(f, dfdt, lap_f; scalars)

this module turns ``F_n, G_n`` (and the potential, for the in-kernel energy sums) into a
small device functor, compiles *the same pipeline source* with it through NVRTC
(``-DPSB_GENERATED_RHS``), and launches that instance with exactly the parameters the
built-in launch would use (``psb_scalar_stage_fused_plan``): rings, stencil, stores,
periodic and peer halos and energy folding are the hand-written code, only the per-site
right-hand side is generated.  Opt-in like the other generated kernels (``jit=True`` /
``PSB_JIT=1``); compiles without a GPU (tests/test_fused_jit.py), not yet run on hardware.
"""

import ctypes as C
import os

import numpy as np

from pystella_b200 import _lib
from pystella_b200.codegen import JitLibrary, expr_to_c, _scalar_name
from pystella_b200.field import (DynamicField, Field, Subscript, Variable, _const_index,
                                 _field_shift, _walk, is_constant)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MAX_USER_SCALARS = 32

__all__ = ["GeneralScalarSystem", "recognise_general_system", "GeneratedFusedStage"]


class GeneralScalarSystem:
    """Pairs ``(f_n, dfdt_n)`` with arbitrary site-local right-hand sides.  Duck-compatible
    with :class:`pystella_b200.step.ScalarSystem` where the stepper needs it."""

    def __init__(self, field, nscalars, has_outer_axis, rhs_f, rhs_d, potential=None):
        self.field = field
        self.f_name = field.name
        self.dot_name = field.dot.name
        self.lap_name = field.lap.name
        self.nscalars = nscalars
        self.has_outer_axis = has_outer_axis
        self.rhs_f, self.rhs_d = rhs_f, rhs_d
        self.potential = potential
        self.dV = [[] for _ in range(nscalars)]       # nothing for the polynomial tables
        self.V = []
        names = set()
        for e in list(rhs_f) + list(rhs_d) + ([potential] if potential is not None else []):
            for node in _walk(e):
                if isinstance(node, Variable) or (isinstance(node, Field)
                                                  and len(node.indices) == 0):
                    names.add(node.name)
        names -= {"sin", "cos", "tan", "exp", "log", "sqrt", "fabs", "abs", "tanh", "sinh",
                  "cosh", "pi"}
        # `a` and `hubble` travel in the kernel's own parameters (host values or device
        # pointers, like the built-in stage); everything else by value
        self.a = Field("a", indices=[]) if "a" in names else 1
        self.hubble = Field("hubble", indices=[]) if "hubble" in names else 0
        self.user_scalars = sorted(names - {"a", "hubble"})
        if len(self.user_scalars) > MAX_USER_SCALARS:
            raise NotImplementedError("too many scalars in a generated right-hand side")
        self.generated = GeneratedFusedStage(self)

    def scalar_names(self):
        return {"a", "hubble"} | set(self.user_scalars)


def recognise_general_system(rhs_dict, potential=None):
    """``rhs_dict`` over one DynamicField ``f`` (keys ``f[n]``, ``f.dot[n]``) whose values
    read ``f``, ``f.dot``, ``f.lap`` at the site itself and any scalars."""
    from pystella_b200.step import _unknown_parts
    dyn = None
    for key in rhs_dict:
        fld, _ = _unknown_parts(key)
        if isinstance(fld, DynamicField):
            if dyn is not None and fld.name != dyn.name:
                raise NotImplementedError("more than one DynamicField in one system")
            dyn = fld
    if dyn is None:
        raise NotImplementedError("no DynamicField among the unknowns")
    if len(dyn.shape) > 1:
        raise NotImplementedError("DynamicField with more than one outer axis")
    n = dyn.shape[0] if dyn.shape else 1
    has_outer = bool(dyn.shape)
    if n > _lib.MAX_SCALARS:
        raise NotImplementedError(f"{n} fields per fused launch (max {_lib.MAX_SCALARS})")
    if len(rhs_dict) != 2 * n:
        raise NotImplementedError(f"expected {2 * n} unknowns, got {len(rhs_dict)}")
    allowed = {dyn.name, dyn.dot.name, dyn.lap.name}
    rhs_f, rhs_d = [], []
    for i in range(n):
        kf = dyn[i] if has_outer else dyn
        kd = dyn.dot[i] if has_outer else dyn.dot
        if kf not in rhs_dict or kd not in rhs_dict:
            raise NotImplementedError(f"missing equation for field {i}")
        for e in (rhs_dict[kf], rhs_dict[kd]):
            for node in _walk(e):
                if isinstance(node, Field) and len(node.indices) > 0:
                    if node.name not in allowed:
                        raise NotImplementedError(
                            f"array '{node.name}' in a fused right-hand side: only "
                            f"{sorted(allowed)} are available in the stage kernel")
                    if any(_field_shift(node)):
                        raise NotImplementedError("shifted operand in a fused right-hand "
                                                  "side (only the Laplacian is stencilled)")
        rhs_f.append(rhs_dict[kf])
        rhs_d.append(rhs_dict[kd])
    return GeneralScalarSystem(dyn, n, has_outer, rhs_f, rhs_d, potential)


def _functor_source(system):
    """CUDA C for ``psb_generated_rhs`` / ``psb_generated_potential``."""
    names = {system.f_name: "f", system.dot_name: "d", system.lap_name: "lap"}

    def leaf(node):
        fld, idx = None, ()
        if isinstance(node, Subscript) and isinstance(node.aggregate, Field) \
                and len(node.aggregate.indices) > 0:
            fld, idx = node.aggregate, _const_index(node.index)
        elif isinstance(node, Field) and len(node.indices) > 0:
            fld = node
        if fld is not None:
            own = _const_index(fld.child.index) if isinstance(fld.child, Subscript) else ()
            sub = tuple(idx) + tuple(own)
            if system.has_outer_axis:
                if len(sub) != 1 or not 0 <= sub[0] < system.nscalars:
                    raise NotImplementedError(f"subscript {sub} of '{fld.name}'")
                m = sub[0]
            else:
                if sub:
                    raise NotImplementedError(f"subscript {sub} of '{fld.name}'")
                m = 0
            return f"{names[fld.name]}[{m}]"
        sname = _scalar_name(node)
        if sname is None:
            return None
        if sname in ("a", "hubble"):
            return sname
        if sname == "pi":
            return "(T)(3.14159265358979323846)"
        if sname in system.user_scalars:
            return f"((T)s[{system.user_scalars.index(sname)}])"
        return None

    lines = ["#define PSB_GENERATED_RHS 1",
             "template <typename T> __device__ __forceinline__ void psb_generated_rhs(",
             "    int n, const T* f, const T* d, const T* lap, T a, T hubble, const double* s,",
             "    T& rhs_f, T& rhs_d) {",
             "  switch (n) {"]
    for i in range(system.nscalars):
        lines.append(f"    case {i}:")
        lines.append(f"      rhs_f = {expr_to_c(system.rhs_f[i], leaf)};")
        lines.append(f"      rhs_d = {expr_to_c(system.rhs_d[i], leaf)};")
        lines.append("      break;")
    lines += ["    default: rhs_f = (T)0; rhs_d = (T)0;", "  }", "}",
              "template <typename T> __device__ __forceinline__ T psb_generated_potential(",
              "    const T* f, T a, T hubble, const double* s) {"]
    pot = system.potential
    lines.append(f"  return {expr_to_c(pot, leaf) if pot is not None else '(T)0'};")
    lines += ["}", '#include "fused_stage_ws.cuh"', ""]
    return "\n".join(lines)


class _JitScalars(C.Structure):
    _fields_ = [("v", C.c_double * MAX_USER_SCALARS)]


def _aligned(ctype, alignment=64):
    """A ``ctype`` instance at an ``alignment``-byte aligned address (+ its keep-alive)."""
    raw = C.create_string_buffer(C.sizeof(ctype) + alignment)
    addr = (C.addressof(raw) + alignment - 1) // alignment * alignment
    return ctype.from_address(addr), raw


class GeneratedFusedStage:
    """Compiles (once per pipeline configuration) and launches the fused stage with the
    generated right-hand side of ``system``."""

    def __init__(self, system):
        self.system = system
        self.source = _functor_source(system)
        self._handles = {}
        self._plan, self._plan_keepalive = _aligned(_lib.FusedPlan)
        self._scalars, self._scalars_keepalive = _aligned(_JitScalars)

    @staticmethod
    def name_expression(plan):
        T = "double" if plan.dtype == _lib.PSB_F64 else "float"
        return (f"psb::fused_stage_ws_kernel<{T}, {plan.h}, {plan.nscalars}, {plan.by}, "
                f"{plan.tz}, {plan.p}, {'true' if plan.energy else 'false'}, {plan.minb}, "
                f"{plan.vec}>")

    def compile(self, plan, strict=False):
        key = (plan.dtype, plan.h, plan.nscalars, plan.by, plan.tz, plan.p, plan.energy,
               plan.minb, plan.vec, bool(strict))
        if key not in self._handles:
            lib = JitLibrary.get()
            opts = [b"--std=c++17", b"-default-device", b"-lineinfo",
                    b"--fmad=false" if strict else b"--fmad=true",
                    b"-I" + os.path.join(HERE, "csrc").encode(),
                    b"-I" + os.path.join(ROOT, "include").encode()]
            hnames = [b"stdint.h", b"stddef.h"]
            hsrc = [b"", b""]
            name_expr = self.name_expression(plan)

            def build():
                handle = lib.psb_jit_compile_ex(
                    self.source.encode(), None, name_expr.encode(), b"sm_100a",
                    (C.c_char_p * len(opts))(*opts), len(opts),
                    (C.c_char_p * 2)(*hnames), (C.c_char_p * 2)(*hsrc), 2)
                if handle == 0:
                    raise _lib.PsbError(
                        "NVRTC compilation of the fused stage with a generated right-hand "
                        "side failed:\n" + lib.psb_jit_last_log().decode(errors="replace")
                        + "\n--- functor ---\n" + self.source)
                return handle
            csrc = os.path.join(HERE, "csrc")
            headers = [os.path.join(csrc, n) for n in ("fused_stage_ws.cuh",
                                                       "fused_stage_impl.cuh",
                                                       "stage_tail.cuh", "common.cuh")]
            headers.append(os.path.join(ROOT, "include", "pystella_b200.h"))
            handle = JitLibrary.cached(self.source, name_expr, [o.decode() for o in opts],
                                       headers, build)
            self._handles[key] = handle
        return self._handles[key]

    def launch(self, lib, args, queue, scalar_values, strict=False):
        """``args``: a filled ScalarStageArgs; replaces ``psb_scalar_stage_fused(args)``."""
        plan = self._plan
        plan.reserved = 0
        rc = lib.psb_scalar_stage_fused_plan(C.byref(args), C.byref(plan))
        _lib.check(lib, rc, "psb_scalar_stage_fused_plan")
        handle = self.compile(plan, strict)
        js = self._scalars
        if len(scalar_values) != len(self.system.user_scalars):
            raise ValueError("wrong number of scalar values for the generated stage")
        for n, v in enumerate(scalar_values):
            js.v[n] = float(v)
        params = (C.c_void_p * 6)(
            C.addressof(plan) + _lib.FusedPlan.params.offset,
            *[C.addressof(plan) + _lib.FusedPlan.tensor_maps.offset + 128 * m
              for m in range(4)],
            C.addressof(js))
        jit = JitLibrary.get()
        rc = jit.psb_jit_launch_params_v(handle, params, plan.grid[0], plan.grid[1],
                                         plan.grid[2], plan.threads, 1, 1, plan.smem_bytes,
                                         int(args.launch_flags), _lib.stream_ptr(queue))
        if rc != 0:
            raise _lib.PsbError("launch of the generated fused stage failed: "
                                + jit.psb_jit_last_log().decode(errors="replace"))
