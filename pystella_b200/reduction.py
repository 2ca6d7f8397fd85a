"""
Reductions: :class:`Reduction` and :class:`FieldStatistics` with the call
contract of ``pystella/reduction.py:80-343``.

The reference launches a kernel that sums over x, then one pyopencl reduction
and one blocking device-to-host read *per expression*, then one pickled MPI
allreduce per expression (``reduction.py:192-204, 237-255``).  Here every
expression of the call is folded by one ``psb_reduce`` launch into a small float64
device vector, which is all-reduced once per distinct operation (NCCL) and read
back once.

Expressions must be polynomial in the array Fields with scalar coefficients
(everything the sectors produce, plus ``fabs`` of such a polynomial for
``abs_max``/``abs_min``); they are brought to that form by
:func:`pystella_b200.field.to_polynomial`.
"""

from itertools import product

import numpy as np
import torch

from pystella_b200 import _kernels, _lib
from pystella_b200._lib import PsbError
from pystella_b200.field import (Call, Field, Variable, _walk, evaluate, is_constant,
                                 substitute, to_polynomial)

__all__ = ["Reduction", "FieldStatistics"]

_OPS = {"avg": 0, "sum": 0, "max": 1, "min": 2}


class Reduction:
    """
    :arg decomp: a :class:`~pystella_b200.DomainDecomposition`.
    :arg input: ``{key: [expr | (expr, op), ...]}``, a Sector (its ``reducers``)
        or a list of Sectors.  ``op`` in ``"avg"`` (default), ``"sum"``, ``"max"``,
        ``"min"`` (``"prod"`` is not provided by the kernel).

    Keyword-only: ``grid_size`` (global number of sites, for averages; inferred by
    an allreduce when omitted), ``callback``, ``rank_shape``, ``halo_shape``,
    ``tmp_instructions`` (list of ``(Variable, expr)`` substituted into the
    reducers, ``elementwise.py:173-203``).
    """

    def __init__(self, decomp, input, **kwargs):
        self.decomp = decomp
        from pystella_b200.sectors import Sector
        if isinstance(input, Sector):
            self.reducers = input.reducers
        elif isinstance(input, list):
            self.reducers = dict(i for s in input for i in s.reducers.items())
        elif isinstance(input, dict):
            self.reducers = input
        else:
            raise NotImplementedError

        self.grid_size = kwargs.pop("grid_size", None)
        #: kernel choice of psb_reduce: {"variant": 0 auto | 1 generic scalar-load kernel}
        self.tune = dict(kwargs.pop("tune", None) or {})
        self.callback = kwargs.pop("callback", lambda x: x)
        self.halo_shape = kwargs.pop("halo_shape", None)
        self.rank_shape = kwargs.pop("rank_shape", None)
        self.strict_fp = kwargs.pop("strict_fp", False)
        self.jit = kwargs.pop("jit", None)
        self.dtype = kwargs.pop("dtype", "float64")
        tmp = kwargs.pop("tmp_instructions", None) or []
        if isinstance(tmp, dict):
            tmp = list(tmp.items())
        self._tmp = tmp

        self.num_reductions = sum(len(i) for i in self.reducers.values())
        self.tmp_dict = {}
        flat, ops = [], []
        i = 0
        for key, val in self.reducers.items():
            self.tmp_dict[key] = range(i, i + len(val))
            i += len(val)
            for v in val:
                expr, op = v if isinstance(v, tuple) else (v, "avg")
                if op not in _OPS:
                    raise NotImplementedError(f"reduction operation {op}")
                flat.append(expr)
                ops.append(op)
        self.reduction_ops = ops

        # substitute temporaries (later ones may use earlier ones)
        subs = {}
        for lhs, rhs in tmp:
            subs[lhs] = substitute(rhs, subs)
        self._polys = []
        self._abs = []
        generated = []      # (scratch Field, expression): reducers outside the menu
        for r, expr in enumerate(flat):
            expr = substitute(expr, subs) if subs else expr
            absolute = False
            inner = expr
            if isinstance(expr, Call) and getattr(expr.function, "name", "") in (
                    "fabs", "abs"):
                absolute = True
                inner = expr.parameters[0]
            try:
                poly = to_polynomial(inner)
            except NotImplementedError:
                from pystella_b200.step import jit_enabled
                if not jit_enabled(self.jit):
                    raise
                # EXPERIMENTAL: evaluate the reducer site by site with a generated
                # element-wise kernel into a scratch array, then fold that array
                scratch = Field(f"_psb_reducer_{r}")
                generated.append((scratch, expr))
                poly, absolute = to_polynomial(scratch), False
            self._polys.append(poly)
            self._abs.append(absolute)
        self._generated_map = None
        self._scratch_names = [fld.name for fld, _ in generated]
        if generated:
            from pystella_b200.codegen import GenericMap
            self._generated_map = GenericMap(generated, dtype=self.dtype,
                                             halo_shape=self.halo_shape,
                                             rank_shape=self.rank_shape)

        # the distinct arrays and scalars
        self._arrays = []
        for poly in self._polys:
            for mono in poly:
                for ref, _ in mono:
                    if ref not in self._arrays:
                        self._arrays.append(ref)
        self._scalar_names = set()
        for poly in self._polys:
            for c in poly.values():
                for node in _walk(c):
                    if isinstance(node, (Variable, Field)):
                        self._scalar_names.add(node.name)

    # ------------------------------------------------------------------------
    def _halo_of(self, ref, h_value):
        off = ref.offset[0] if ref.offset else 0
        if is_constant(off):
            return int(off)
        if isinstance(off, Variable) and off.name == "h":
            if h_value is None:
                raise ValueError("pass halo_shape= to Reduction for padded arrays")
            return h_value
        return int(evaluate(off, {"h": h_value}))

    def reduce_array(self, arr, op):
        """Reduce one (device) array over all ranks (reduction.py:192-204)."""
        if op not in _OPS:
            raise NotImplementedError(f"reduction operation {op}")
        dev = _kernels.reduce_poly([arr], [0], [(0, 1.0, [1])], [_OPS[op]], [0],
                                   tuple(arr.shape[-3:]), strict=self.strict_fp)
        val = self.decomp.allreduce(dev, op=op) if self.decomp.comm is not None else dev
        return float(val.cpu()[0])

    def __call__(self, queue=None, filter_args=False, **kwargs):
        """
        :returns: ``dict`` with the keys of the input whose values are
            ``np.array([...])`` of the reduced expressions; averages are divided
            by ``grid_size`` (reduction.py:206-255).
        """
        kwargs.pop("allocator", None)
        h = self.halo_shape
        if isinstance(h, (tuple, list)):
            h = h[0]
        if self._generated_map is not None:
            kwargs = self._run_generated(queue, kwargs)
        arrays, halos = [], []
        for ref in self._arrays:
            if ref.name not in kwargs:
                raise TypeError(f"missing array argument '{ref.name}'")
            t = kwargs[ref.name]
            if not isinstance(t, torch.Tensor):
                raise TypeError(f"'{ref.name}' must be a torch.Tensor on the GPU")
            if not _lib.on_device(t):
                raise PsbError(f"'{ref.name}' lives on {t.device}: reductions run on "
                               "the GPU only (no CPU fallback)")
            if any(ref.shift):
                raise NotImplementedError("shifted operands in a reduction")
            sub = t[ref.subscript] if ref.subscript else t
            if sub.dim() != 3:
                raise ValueError(f"'{ref.name}' has wrong number of outer axes")
            arrays.append(sub)
            halos.append(self._halo_of(ref, h))
        if not arrays:
            raise NotImplementedError("a Reduction needs at least one array operand")
        shape = tuple(n - 2 * g for n, g in zip(arrays[0].shape, (halos[0],) * 3))

        ctx = {}
        for n in self._scalar_names:
            if n in kwargs:
                v = kwargs[n]
                v = v.cpu().numpy() if isinstance(v, torch.Tensor) else v
                ctx[n] = float(np.asarray(v).reshape(-1)[0])
        terms, ops, flags = [], [], []
        for r, (poly, op, ab) in enumerate(zip(self._polys, self.reduction_ops,
                                               self._abs)):
            ops.append(_OPS[op])
            flags.append(1 if ab else 0)
            for mono, c in poly.items():
                powers = [0] * len(arrays)
                for ref, pw in mono:
                    powers[self._arrays.index(ref)] = pw
                cval = c if is_constant(c) else evaluate(c, ctx)
                terms.append((r, float(cval), powers))

        # everything after the launch (slicing, the all-reduce, the read-back) is ordered on
        # the stream the kernels run on: `queue` may be a non-current stream (the reference
        # passes the queue explicitly, pystella/reduction.py:206)
        import contextlib
        ctx_stream = torch.cuda.stream(queue) if queue is not None else contextlib.nullcontext()
        with ctx_stream:
            out = self._launch(arrays, halos, terms, ops, flags, shape, queue)

            # one allreduce per distinct operation
            if self.decomp.comm is not None:
                for opname, code in (("sum", 0), ("max", 1), ("min", 2)):
                    idx = [r for r, o in enumerate(ops) if o == code]
                    if idx:
                        sel = out[idx].contiguous()
                        sel = self.decomp.allreduce(sel, op=opname)
                        out[idx] = sel
            host = out.cpu().numpy()

        grid_size = self.grid_size
        if grid_size is None and "avg" in self.reduction_ops:
            grid_size = self.decomp.allreduce(int(np.prod(shape)))
        vals = {}
        for key, sub_indices in self.tmp_dict.items():
            red = []
            for j in sub_indices:
                v = host[j]
                if self.reduction_ops[j] == "avg":
                    v = v / grid_size
                red.append(v)
            vals[key] = np.array(red)
        return self.callback(vals)

    def _run_generated(self, queue, kwargs):
        gm = self._generated_map
        shape = gm.rank_shape
        if shape is None:
            for name, sub, field in gm.ksrc.slots:
                if name in self._scratch_names:
                    continue
                if name not in kwargs:
                    raise TypeError(f"missing array argument '{name}'")
                t = kwargs[name][sub] if sub else kwargs[name]
                g = gm._halo_of(field)
                shape = tuple(n - 2 * g for n in t.shape[-3:])
                device = t.device
                break
            else:
                raise NotImplementedError("a Reduction needs at least one array operand")
        else:
            device = next(v.device for v in kwargs.values() if isinstance(v, torch.Tensor))
        kwargs = dict(kwargs)
        for name in self._scratch_names:
            kwargs[name] = torch.empty(shape, dtype=getattr(torch, np.dtype(self.dtype).name),
                                       device=device)
        gm(queue, **kwargs)
        return kwargs

    def _launch(self, arrays, halos, terms, ops, flags, shape, queue):
        # split into launches that respect the kernel's static limits
        from pystella_b200._lib import MAX_RED_OUTPUTS, MAX_RED_TERMS
        out = torch.empty(len(ops), dtype=torch.float64, device=arrays[0].device)
        start = 0
        while start < len(ops):
            stop = start
            nterm = 0
            while stop < len(ops) and stop - start < MAX_RED_OUTPUTS:
                n_r = sum(1 for t in terms if t[0] == stop)
                if nterm + n_r > MAX_RED_TERMS:
                    break
                nterm += n_r
                stop += 1
            if stop == start:
                raise NotImplementedError("a single reducer with too many monomials")
            sel_terms = [(r - start, c, p) for r, c, p in terms if start <= r < stop]
            out[start:stop] = _kernels.reduce_poly(
                arrays, halos, sel_terms, ops[start:stop], flags[start:stop], shape,
                queue=queue, strict=self.strict_fp, variant=self.tune.get("variant", 0))
            start = stop
        return out


class FieldStatistics(Reduction):
    """
    Mean and variance (optionally max/min/abs_max/abs_min) of fields
    (``pystella/reduction.py:258-343``).
    """

    def __init__(self, decomp, halo_shape, **kwargs):
        self.min_max = kwargs.pop("max_min", False)
        f = Field("f", offset="h")
        reducers = {"mean": [f], "variance": [f**2]}
        if self.min_max:
            from pystella_b200.field import var
            fabs = var("fabs")
            reducers["max"] = [(f, "max")]
            reducers["min"] = [(f, "min")]
            reducers["abs_max"] = [(fabs(f), "max")]
            reducers["abs_min"] = [(fabs(f), "min")]
        super().__init__(decomp, reducers, halo_shape=halo_shape, **kwargs)

    def __call__(self, f, queue=None, allocator=None):
        """Loops over the outer axes of ``f``; returns arrays of that outer shape."""
        outer_shape = f.shape[:-3]
        slices = list(product(*[range(n) for n in outer_shape]))
        out = {k: np.zeros(outer_shape) for k in self.reducers}
        for s in slices:
            stats = super().__call__(queue, f=f[s])
            for k, v in stats.items():
                if k == "variance":
                    out[k][s] = stats["variance"][0] - stats["mean"][0]**2
                else:
                    out[k][s] = v[0]
        return {k: np.array(v) for k, v in out.items()}
