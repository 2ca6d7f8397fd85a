"""
Histograms next to the time loop: :class:`Histogrammer`
and :class:`FieldHistogrammer` with the call contracts of ``pystella/histogram.py:29-350``.

A histogram is ``hist[floor(bin_expr)] += weight_expr`` over all sites, where both
expressions are symbolic (Fields, scalars by keyword).  As in the reference (which has loopy
generate the kernel, histogram.py:60-170) ALL histograms of one :class:`Histogrammer` are
evaluated by ONE run-time generated kernel in a single pass over the arrays
(``codegen.generate_histogram_kernel``: warp-aggregated runs -> block-private shared-memory
copies -> one global atomic per bin and block, accumulation in float64); out-of-range bins
are dropped (the reference's kernel would write out of bounds; its
:class:`FieldHistogrammer` clips, as does this one).  The bounds of
:class:`FieldHistogrammer` come from one ``psb_reduce`` launch (max/min of ``f`` and
``|f|``).

Arrays must be CUDA tensors; host tensors are accepted only under the decomposition's
``host_ops`` test hook (like ``share_halos``) and are then evaluated with tensor operations.
"""

from contextlib import nullcontext as _nullcontext
from itertools import product

import numpy as np
import torch

from pystella_b200 import _lib
from pystella_b200._lib import PsbError
from pystella_b200.field import (Call, Field, Power, Product, Quotient, Subscript, Sum,
                                 Variable, _const_index, _field_shift, is_constant)

__all__ = ["Histogrammer", "FieldHistogrammer"]


def _t_min(a, b):
    return torch.minimum(*_as_pair(a, b))


def _t_max(a, b):
    return torch.maximum(*_as_pair(a, b))


def _as_pair(a, b):
    ref = a if isinstance(a, torch.Tensor) else b
    if not isinstance(ref, torch.Tensor):
        return torch.tensor(float(a)), torch.tensor(float(b))
    to_t = lambda x: x if isinstance(x, torch.Tensor) else torch.full_like(ref, float(x))  # noqa
    return to_t(a), to_t(b)


def _unary(fn, np_fn):
    return lambda x: fn(x) if isinstance(x, torch.Tensor) else float(np_fn(x))


_FUNCS = {
    "fabs": _unary(torch.abs, np.abs), "abs": _unary(torch.abs, np.abs),
    "log": _unary(torch.log, np.log), "exp": _unary(torch.exp, np.exp),
    "sqrt": _unary(torch.sqrt, np.sqrt), "floor": _unary(torch.floor, np.floor),
    # C rounding (half away from zero), what the reference's generated kernel does
    "round": _unary(lambda x: torch.sign(x) * torch.floor(torch.abs(x) + 0.5),
                    lambda x: np.sign(x) * np.floor(np.abs(x) + 0.5)),
    "sin": _unary(torch.sin, np.sin), "cos": _unary(torch.cos, np.cos),
    "tanh": _unary(torch.tanh, np.tanh),
    "min": _t_min, "max": _t_max, "fmin": _t_min, "fmax": _t_max,
}


class _TensorEvaluator:
    """Evaluate an expression tree site-wise: array Fields become interior views of the
    tensors passed by keyword (halo offset and stencil shift applied), everything else a
    Python float."""

    def __init__(self, kwargs, halo_shape, rank_shape):
        self.kwargs = kwargs
        self.halo = halo_shape
        self.rank_shape = rank_shape
        self.device = None

    def _halo_of(self, fld):
        off = fld.base_offset[0] if fld.base_offset else 0
        if is_constant(off):
            return int(off)
        h = self.halo[0] if isinstance(self.halo, (tuple, list)) else self.halo
        if h is None:
            raise ValueError("pass halo_shape= for arrays with a symbolic offset")
        from pystella_b200.field import evaluate
        return int(evaluate(off, {"h": h, "hx": h, "hy": h, "hz": h}))

    def _array(self, fld, outer):
        if fld.name not in self.kwargs:
            raise TypeError(f"missing array argument '{fld.name}'")
        t = self.kwargs[fld.name]
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"'{fld.name}' must be a torch.Tensor")
        own = fld.child.index if isinstance(fld.child, Subscript) else ()
        sub = _const_index(tuple(outer) + tuple(own))
        view = t[sub] if sub else t
        if view.dim() != 3:
            raise ValueError(f"'{fld.name}' has wrong number of outer axes")
        g = self._halo_of(fld)
        shape = tuple(n - 2 * g for n in view.shape)
        if self.rank_shape is None:
            self.rank_shape = shape
        elif tuple(self.rank_shape) != shape:
            raise ValueError(f"'{fld.name}' has rank shape {shape}, expected "
                             f"{tuple(self.rank_shape)}")
        self.device = t.device
        sh = _field_shift(fld)
        idx = tuple(slice(g + s, g + s + n) for s, n in zip(sh, shape))
        return view[idx]

    def _scalar(self, name, idx=()):
        if name not in self.kwargs:
            if name == "pi":
                return float(np.pi)
            raise TypeError(f"missing scalar argument '{name}'")
        v = self.kwargs[name]
        v = v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        return float(v[idx]) if idx else float(v.reshape(-1)[0])

    def __call__(self, e):
        if is_constant(e):
            return float(e)
        if isinstance(e, Field):
            if len(e.indices) > 0:
                return self._array(e, ())
            return self._scalar(e.name)
        if isinstance(e, Variable):
            return self._scalar(e.name)
        if isinstance(e, Subscript):
            agg = e.aggregate
            if isinstance(agg, Field) and len(agg.indices) > 0:
                return self._array(agg, e.index)
            if isinstance(agg, (Variable, Field)):
                return self._scalar(agg.name, _const_index(e.index))
            raise NotImplementedError(f"subscript of {type(agg).__name__}")
        if isinstance(e, Sum):
            vals = [self(c) for c in e.children]
            out = vals[0]
            for v in vals[1:]:
                out = out + v
            return out
        if isinstance(e, Product):
            vals = [self(c) for c in e.children]
            out = vals[0]
            for v in vals[1:]:
                out = out * v
            return out
        if isinstance(e, Quotient):
            return self(e.numerator) / self(e.denominator)
        if isinstance(e, Power):
            return self(e.base) ** self(e.exponent)
        if isinstance(e, Call):
            name = getattr(e.function, "name", None)
            if name not in _FUNCS:
                raise NotImplementedError(f"function {name} in a histogram expression")
            return _FUNCS[name](*[self(p) for p in e.parameters])
        raise TypeError(f"cannot evaluate {type(e).__name__}")


class Histogrammer:
    """
    ``Histogrammer(decomp, histograms, num_bins, dtype, halo_shape=0, rank_shape=None)``
    with ``histograms = {name: (bin_expr, weight_expr)}`` (histogram.py:29-208).
    ``__call__(queue=None, filter_args=False, **arrays_and_scalars)`` returns
    ``{name: numpy array of num_bins}``, summed over ranks.
    """

    def __init__(self, decomp, histograms, num_bins, dtype, **kwargs):
        self.decomp = decomp
        self.histograms = histograms
        self.num_bins = int(num_bins)
        self.dtype = np.dtype(dtype)
        self.halo_shape = kwargs.pop("halo_shape", 0)
        self.rank_shape = kwargs.pop("rank_shape", None)
        self._ignored_kwargs = kwargs
        self._generated = {}

    def _check_device(self, kwargs):
        for name, v in kwargs.items():
            if isinstance(v, torch.Tensor) and v.dim() >= 3 and not _lib.on_device(v) \
                    and getattr(self.decomp, "host_ops", None) is None:
                raise PsbError(f"'{name}' lives on {v.device}: histograms run on the GPU "
                               "only (no CPU fallback)")

    def _device_histograms(self, queue, kwargs):
        """-> float64 CUDA tensor (num_hists, num_bins) from the generated kernel."""
        from pystella_b200.codegen import GeneratedHistogram
        arrays = [v for v in kwargs.values() if isinstance(v, torch.Tensor) and v.dim() >= 3]
        tdtype = arrays[0].dtype if arrays else torch.float64
        if tdtype not in (torch.float64, torch.float32):
            raise TypeError(f"histograms of {tdtype} arrays are not supported")
        key = str(tdtype)
        if key not in self._generated:
            self._generated[key] = GeneratedHistogram(
                list(self.histograms.values()), self.num_bins,
                dtype=str(tdtype).replace("torch.", ""), halo_shape=self.halo_shape,
                rank_shape=self.rank_shape)
        return self._generated[key](queue, **kwargs)

    def __call__(self, queue=None, filter_args=False, **kwargs):
        kwargs.pop("allocator", None)
        self._check_device(kwargs)
        tensors = [v for v in kwargs.values() if isinstance(v, torch.Tensor) and v.dim() >= 3]
        on_host = any(not _lib.on_device(v) for v in tensors)
        if on_host or (not tensors and getattr(self.decomp, "host_ops", None) is not None
                       and not torch.cuda.is_available()):
            hist = self._host_histograms(kwargs)
        else:
            hist = self._device_histograms(queue, kwargs)
        if self.decomp.comm is not None:
            hist = self.decomp.allreduce(hist)
        full = hist.cpu().numpy().astype(self.dtype)
        return {name: full[j] for j, name in enumerate(self.histograms)}

    def _host_histograms(self, kwargs):
        """Host tensors (test hook of the decomposition): tensor operations + index_add_."""
        ev = _TensorEvaluator(kwargs, self.halo_shape, self.rank_shape)
        rows = []
        for bin_expr, weight_expr in self.histograms.values():
            b, w = ev(bin_expr), ev(weight_expr)
            rows.append((b, w))
        shape = ev.rank_shape
        if shape is None:
            raise ValueError("no array argument: pass rank_shape= to the constructor")
        device = ev.device
        if device is None:
            device = torch.device("cpu")
        nsites = int(np.prod(shape))
        hist = torch.zeros((len(rows), self.num_bins), dtype=torch.float64, device=device)
        for j, (b, w) in enumerate(rows):
            if not isinstance(b, torch.Tensor):          # every site in one bin
                k = int(np.floor(b))
                total = (w.double().sum() if isinstance(w, torch.Tensor) else w * nsites)
                if 0 <= k < self.num_bins:
                    hist[j, k] = total
                continue
            idx = torch.floor(b).to(torch.int64).reshape(-1)
            wt = (w.double().reshape(-1) if isinstance(w, torch.Tensor)
                  else torch.full((nsites,), float(w), dtype=torch.float64, device=device))
            ok = (idx >= 0) & (idx < self.num_bins)
            hist[j].index_add_(0, idx[ok], wt[ok])
        return hist


class FieldHistogrammer(Histogrammer):
    """
    Linear and logarithmic histograms of fields (histogram.py:210-350):
    ``FieldHistogrammer(decomp, num_bins, dtype, halo_shape=0, rank_shape=None)``;
    ``__call__(f, queue=None, min_f=..., max_f=..., min_log_f=..., max_log_f=...)`` loops
    over the outer axes of ``f`` and returns ``linear``, ``linear_bins``, ``log``,
    ``log_bins``.  Bounds are computed with one reduction per slice unless all four are
    passed.
    """

    def __init__(self, decomp, num_bins, dtype, **kwargs):
        from pystella_b200.field import var
        halo_shape = kwargs.pop("halo_shape", 0)
        f = Field("f", offset=halo_shape)
        max_f, min_f = var("max_f"), var("min_f")
        max_log_f, min_log_f = var("max_log_f"), var("min_log_f")
        log, fabs, vmin, vmax = var("log"), var("fabs"), var("min"), var("max")

        def clip(expr):
            return vmax(vmin(expr, num_bins - 1), 0)

        linear_bin = (f - min_f) / (max_f - min_f)
        log_bin = (log(fabs(f)) - min_log_f) / (max_log_f - min_log_f)
        histograms = {"linear": (clip(linear_bin * num_bins), 1),
                      "log": (clip(log_bin * num_bins), 1)}
        super().__init__(decomp, histograms, num_bins, dtype, halo_shape=halo_shape,
                         **kwargs)
        self._f = f
        self._bound_names = ("max_f", "min_f", "max_log_f", "min_log_f")

    def _bounds(self, f, queue=None):
        if _lib.on_device(f):
            # one pass: max f, min f, max |f|, min |f| (psb_reduce ops 1 = max, 2 = min,
            # flag 1 = absolute value); log is monotone, so the bounds of log|f| are the
            # logs of the bounds of |f|
            from pystella_b200 import _kernels
            h = _TensorEvaluator({"f": f}, self.halo_shape, None)._halo_of(self._f)
            shape = tuple(n - 2 * h for n in f.shape[-3:])
            terms = [(r, 1.0, (1,)) for r in range(4)]
            with torch.cuda.stream(queue) if queue is not None else _nullcontext():
                vals = _kernels.reduce_poly([f], [h], terms, [1, 2, 1, 2], [0, 0, 1, 1],
                                            shape, queue=queue)
                vals = vals * torch.tensor([1., -1., 1., -1.], dtype=torch.float64,
                                           device=f.device)
        else:
            ev = _TensorEvaluator({"f": f}, self.halo_shape, None)
            x = ev(self._f)
            ax = torch.abs(x)
            vals = torch.stack([x.max(), -x.min(), ax.max(), -ax.min()]).double()
        if self.decomp.comm is not None:
            vals = self.decomp.allreduce(vals, op="max")
        vals = vals.cpu().numpy()
        with np.errstate(divide="ignore"):
            return {"max_f": vals[0], "min_f": -vals[1], "max_log_f": np.log(vals[2]),
                    "min_log_f": np.log(-vals[3])}

    def __call__(self, f, queue=None, **kwargs):
        kwargs.pop("allocator", None)
        self._check_device({"f": f})
        outer_shape = tuple(f.shape[:-3])
        bounds_passed = set(self._bound_names).issubset(kwargs)
        out = {}
        for key in ("linear", "log"):
            out[key] = np.zeros(outer_shape + (self.num_bins,))
            out[key + "_bins"] = np.zeros(outer_shape + (self.num_bins + 1,))
        for s in product(*[range(n) for n in outer_shape]):
            if bounds_passed:
                bounds = {k: float(np.asarray(kwargs[k])[s]) for k in self._bound_names}
            else:
                bounds = self._bounds(f[s], queue)
            hists = super().__call__(queue, f=f[s], **bounds)
            for key, val in hists.items():
                out[key][s] = val
            out["linear_bins"][s] = np.linspace(bounds["min_f"], bounds["max_f"],
                                                self.num_bins + 1)
            out["log_bins"][s] = np.exp(np.linspace(bounds["min_log_f"],
                                                    bounds["max_log_f"],
                                                    self.num_bins + 1))
        return out
