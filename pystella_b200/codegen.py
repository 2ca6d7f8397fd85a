"""
EXPERIMENTAL (opt-in, ``PSB_JIT=1`` / ``jit=True``): run-time specialisation of
element-wise statements, the role loopy plays in the reference
(``pystella/elementwise.py:164-235``).

A list of statements ``assignee = expr`` over :class:`~pystella_b200.Field` s is
turned into ONE small CUDA C kernel (one thread per lattice site, all statements
in program order, every array element touched once for reading and once for
writing), compiled with NVRTC for sm_100a by ``libpystella_b200_jit.so``
(``include/pystella_b200_jit.h``) and launched through the driver API.  The
hand-written kernels stay the path for the sectors on the menu; this covers
arbitrary ``rhs_dict`` / ``ElementWiseMap`` users (non-polynomial potentials,
custom stencils, extra unknowns).

Everything up to and including NVRTC compilation is exercised without a GPU
(``tests/test_codegen.py``).
"""

import ctypes as C
import os

import numpy as np

from pystella_b200.field import (Call, Field, Power, Product, Quotient, Subscript, Sum,
                                 Variable, _field_shift, _const_index, is_constant)

__all__ = ["expr_to_c", "KernelSource", "generate_kernel", "JitLibrary", "GenericMap",
           "generate_histogram_kernel", "GeneratedHistogram", "generate_stencil_kernel",
           "TiledStencilMap"]

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_ARRAYS, MAX_SCALARS = 48, 32

_C_FUNCS = {"fabs": "fabs", "abs": "fabs", "sqrt": "sqrt", "exp": "exp", "log": "log",
            "sin": "sin", "cos": "cos", "tan": "tan", "tanh": "tanh", "sinh": "sinh",
            "cosh": "cosh", "floor": "floor", "round": "round", "min": "fmin",
            "max": "fmax", "fmin": "fmin", "fmax": "fmax"}


class _Ref:
    """One array operand: tensor name, constant outer subscript, stencil shift.  A raw
    ``var("x")[i + a, j + b, k + c]`` operand has no Field: ``shift`` then holds the
    absolute offsets ``(a, b, c)`` into the array as passed (``absolute``)."""

    def __init__(self, field, subscript, shift, name=None):
        self.field = field
        self.name = field.name if field is not None else name
        self.subscript = tuple(subscript)
        self.shift = tuple(shift)
        self.absolute = field is None

    @property
    def key(self):
        return (self.name, self.subscript, self.shift)

    @property
    def slot_key(self):          # one pointer/halo slot per (array, outer subscript)
        return (self.name, self.subscript)


def _is_array(fld):
    return isinstance(fld, Field) and len(fld.indices) > 0


_HALO_VALUES = {}       # h, hx, hy, hz while a kernel is being generated


def _raw_ref(e):
    """``var("x")[..., i + a, j + b, k + c]`` (the reference's tests index plain variables
    this way, test/test_stencil.py:48-56) -> _Ref with absolute offsets, else None."""
    if not (isinstance(e, Subscript) and isinstance(e.aggregate, Variable)
            and len(e.index) >= 3):
        return None
    from pystella_b200.field import evaluate
    offsets = []
    for axis, idx in zip("ijk", e.index[-3:]):
        names = {n.name for n in _walk_vars(idx)}
        if axis not in names or names - {axis, "h", "hx", "hy", "hz"}:
            return None
        base = dict(_HALO_VALUES, i=0, j=0, k=0)
        try:
            at0 = evaluate(idx, base)
            at1 = evaluate(idx, dict(base, **{axis: 1}))
        except KeyError:
            raise ValueError("pass halo_shape= to fix `h` in array subscripts") from None
        if at1 - at0 != 1 or int(at0) != at0:
            return None
        offsets.append(int(at0))
    return _Ref(None, _const_index(e.index[:-3]), offsets, name=e.aggregate.name)


def _walk_vars(e):
    from pystella_b200.field import _walk
    return [n for n in _walk(e) if isinstance(n, Variable)]


def _as_ref(e):
    """-> _Ref if ``e`` denotes an array element, else None."""
    raw = _raw_ref(e)
    if raw is not None:
        return raw
    if isinstance(e, Field) and _is_array(e):
        sub = e.child.index if isinstance(e.child, Subscript) else ()
        return _Ref(e, _const_index(sub), _field_shift(e))
    if isinstance(e, Subscript) and _is_array(e.aggregate):
        fld = e.aggregate
        own = fld.child.index if isinstance(fld.child, Subscript) else ()
        return _Ref(fld, _const_index(tuple(e.index) + tuple(own)), _field_shift(fld))
    return None


def _scalar_name(e):
    """-> C-safe name if ``e`` is a by-value scalar: a Variable, an index-free Field, or
    a constant subscript of either (``a[1]`` -> ``a__1``)."""
    if isinstance(e, Variable):
        return e.name
    if isinstance(e, Field) and not _is_array(e) and not isinstance(e.child, Subscript):
        return e.name
    if isinstance(e, Subscript):
        base = e.aggregate
        if isinstance(base, Variable) or (isinstance(base, Field) and not _is_array(base)
                                          and not isinstance(base.child, Subscript)):
            try:
                idx = _const_index(e.index)
            except NotImplementedError:
                return None
            return base.name + "__" + "_".join(str(i) for i in idx)
    return None


def _lit(x):
    if isinstance(x, (bool, np.bool_)):
        x = int(x)
    if isinstance(x, (int, np.integer)):
        return f"(T)({int(x)})"
    return f"(T)({float(x):.17g})"


def expr_to_c(e, leaf):
    """
    C expression (type ``T``) for ``e``; ``leaf(node)`` must return the C name of
    array elements / scalars / temporaries.  Small non-negative integer powers
    become repeated products (what a code generator emits, and the association
    order of the hand-written kernels); other powers use ``pow``.
    """
    if is_constant(e):
        return _lit(e)
    name = leaf(e)
    if name is not None:
        return name
    if isinstance(e, Sum):
        return "(" + " + ".join(expr_to_c(c, leaf) for c in e.children) + ")"
    if isinstance(e, Product):
        return "(" + " * ".join(expr_to_c(c, leaf) for c in e.children) + ")"
    if isinstance(e, Quotient):
        return f"({expr_to_c(e.numerator, leaf)} / {expr_to_c(e.denominator, leaf)})"
    if isinstance(e, Power):
        base = expr_to_c(e.base, leaf)
        ex = e.exponent
        if is_constant(ex) and float(ex).is_integer() and 0 <= int(ex) <= 8:
            n = int(ex)
            return "(T)(1)" if n == 0 else "(" + " * ".join([base] * n) + ")"
        if is_constant(ex) and float(ex).is_integer() and -8 <= int(ex) < 0:
            n = -int(ex)
            return "((T)(1) / (" + " * ".join([base] * n) + "))"
        return f"pow({base}, {expr_to_c(ex, leaf)})"
    if isinstance(e, Call):
        fn = _C_FUNCS.get(getattr(e.function, "name", None))
        if fn is None:
            raise NotImplementedError(f"function {e.function} in generated code")
        return f"{fn}(" + ", ".join(expr_to_c(p, leaf) for p in e.parameters) + ")"
    raise NotImplementedError(f"cannot generate code for {type(e).__name__}: {e}")


class KernelSource:
    """Generated source + the binding tables the launcher needs."""

    def __init__(self, source, name, slots, scalars, dtype, domain_slots=()):
        self.source = source
        self.name = name
        self.slots = slots          # list of (array name, outer subscript, Field or None)
        self.scalars = scalars      # list of scalar names
        self.dtype = dtype
        #: raw (Field-less) arrays written at offset (0, 0, 0): their shape is the domain
        self.domain_slots = tuple(domain_slots)


def _analyse_map(map_instructions, tmp_instructions, halo_shape):
    """Operands and statement bodies of an element-wise map: -> ``(slots, refs, scalars,
    body)`` (shared by the direct and the shared-memory-tile emitters)."""
    def as_list(x):
        if x is None:
            return []
        return list(x.items()) if isinstance(x, dict) else list(x)

    tmps, maps = as_list(tmp_instructions), as_list(map_instructions)
    if not maps:
        raise ValueError("no map instructions")
    _HALO_VALUES.clear()
    if halo_shape is not None:
        hs = (halo_shape,) * 3 if isinstance(halo_shape, (int, np.integer)) \
            else tuple(halo_shape)
        _HALO_VALUES.update(h=hs[0], hx=hs[0], hy=hs[1], hz=hs[2])

    slots, slot_index = [], {}
    refs, scalars = {}, []
    tmp_names = []

    def slot_of(ref):
        if ref.slot_key not in slot_index:
            if len(slots) >= MAX_ARRAYS:
                raise NotImplementedError("too many distinct arrays in one generated kernel")
            slot_index[ref.slot_key] = len(slots)
            slots.append((ref.name, ref.subscript, ref.field))
        return slot_index[ref.slot_key]

    def local_of(ref):
        if ref.key not in refs:
            refs[ref.key] = dict(ref=ref, slot=slot_of(ref), var=f"a{len(refs)}",
                                 read=False, written=False)
        return refs[ref.key]

    def leaf(node):
        ref = _as_ref(node)
        if ref is not None:
            info = local_of(ref)
            if not info["written"]:
                info["read"] = True
            return info["var"]
        sname = _scalar_name(node)
        if sname is not None:
            if sname in tmp_names:
                return f"t_{sname}"
            if sname not in scalars:
                if len(scalars) >= MAX_SCALARS:
                    raise NotImplementedError("too many scalars in one generated kernel")
                scalars.append(sname)
            return f"s_{sname}"
        return None

    body = []
    for lhs, rhs in tmps:
        tname = _scalar_name(lhs)
        if tname is None:
            raise NotImplementedError("temporaries must be plain variables")
        body.append(f"  const T t_{tname} = {expr_to_c(rhs, leaf)};")
        tmp_names.append(tname)
    for lhs, rhs in maps:
        ref = _as_ref(lhs)
        if ref is None:
            raise ValueError("map assignees must be Fields (array elements)")
        if any(ref.shift) and not ref.absolute:
            raise NotImplementedError("assignment to a shifted array element")
        code = expr_to_c(rhs, leaf)
        info = local_of(ref)
        body.append(f"  {info['var']} = {code};")
        info["written"] = True
    # a shifted read of an array that is also written would see a mix of old and new
    written = {info["ref"].slot_key: info["ref"].shift
               for info in refs.values() if info["written"]}
    for info in refs.values():
        key = info["ref"].slot_key
        if key in written and info["ref"].shift != written[key]:
            raise NotImplementedError(
                f"stencil read of '{info['ref'].name}', which the same kernel writes "
                "(evaluate the stencil into a separate array first)")

    return slots, refs, scalars, body


def generate_kernel(map_instructions, tmp_instructions=None, dtype="double",
                    name="psb_generated_map", halo_shape=None):
    """
    ``map_instructions`` / ``tmp_instructions``: lists of ``(assignee, expr)`` (dicts
    are accepted).  Temporaries (Variables) are evaluated first, in order; then every
    map statement in order, later statements seeing earlier assignments
    (``elementwise.py:173-203``).
    """
    slots, refs, scalars, body = _analyse_map(map_instructions, tmp_instructions, halo_shape)
    lines = [
        "// generated by pystella_b200.codegen -- do not edit",
        f"typedef {dtype} T;",
        "struct psb_jit_args {",
        f"  void* arrays[{MAX_ARRAYS}]; int halos[{MAX_ARRAYS}];",
        f"  double scalars[{MAX_SCALARS}]; long long Nx, Ny, Nz;",
        "};",
        "#define IDX(h, si, sj, sk) ((((long long)(i + (h) + (si))) * (Ny + 2 * (h)) + "
        "(j + (h) + (sj))) * (Nz + 2 * (h)) + (k + (h) + (sk)))",
        "#define IDXA(h, a, b, c) ((((long long)(i + (a))) * (Ny + 2 * (h)) + "
        "(j + (b))) * (Nz + 2 * (h)) + (k + (c)))",
        f'extern "C" __global__ void __launch_bounds__(256) {name}(const psb_jit_args args) {{',
        "  const long long Ny = args.Ny, Nz = args.Nz;",
        "  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;",
        "  const long long j = (long long)blockIdx.y * blockDim.y + threadIdx.y;",
        "  const long long i = blockIdx.z;",
        "  if (k >= Nz || j >= Ny || i >= args.Nx) return;",
    ]
    for n, sname in enumerate(scalars):
        lines.append(f"  const T s_{sname} = (T)args.scalars[{n}];")
    for info in refs.values():
        s, (si, sj, sk) = info["slot"], info["ref"].shift
        macro = "IDXA" if info["ref"].absolute else "IDX"
        addr = f"((T*)args.arrays[{s}])[{macro}(args.halos[{s}], {si}, {sj}, {sk})]"
        info["addr"] = addr
        if info["read"]:
            lines.append(f"  T {info['var']} = {addr};")
        else:
            lines.append(f"  T {info['var']};")
    lines += body
    for info in refs.values():
        if info["written"]:
            lines.append(f"  {info['addr']} = {info['var']};")
    lines.append("}")
    domain = sorted({info["slot"] for info in refs.values()
                     if info["written"] and info["ref"].absolute
                     and not any(info["ref"].shift)})
    return KernelSource("\n".join(lines) + "\n", name, slots, scalars, dtype, domain)


# ---------------------------------------------------------------------------
# stencil maps with a shared-memory tile (pystella/stencil.py:36-143)
# ---------------------------------------------------------------------------

TILE_TZ, TILE_TY, TILE_RY = 64, 4, 2     # block (64, 4); every thread owns 2 rows: tile 8 x 64
TILE_SMEM_LIMIT = 200 * 1024


def generate_stencil_kernel(map_instructions, tmp_instructions=None, dtype="double",
                            name="psb_generated_stencil", halo_shape=None, prefetch_args=()):
    """
    The reference's ``Stencil`` / ``StreamingStencil`` stage ``prefetch_args`` in local
    memory (``lp.add_prefetch(..., fetch_bounding_box=True)``, streaming along ``i`` for the
    latter: stencil.py:66-84, 113-133).  Here both get the streaming form, which is the
    one that reads every plane once: a CTA owns a (8 x 64) column of sites in (y, z) and
    marches along x over a chunk of planes; for every array named in ``prefetch_args`` the
    bounding box of its stencil offsets is kept in shared memory as a ring of
    ``hi_x - lo_x + 1`` planes of ``(8 + hi_y - lo_y) x (64 + hi_z - lo_z)`` elements,
    filled cooperatively (coalesced rows) one plane per step; the statements then read that
    array from shared memory only.  Arrays not named are read directly.  Same statement
    semantics as :func:`generate_kernel`.
    """
    slots, refs, scalars, body = _analyse_map(map_instructions, tmp_instructions, halo_shape)
    prefetch = []
    for pname in prefetch_args:
        pname = str(pname)
        group = [info for info in refs.values() if info["ref"].name == pname]
        if not group:
            raise ValueError(f"prefetch_args names '{pname}', which no statement reads")
        if any(info["written"] for info in group):
            continue                      # read at the centre only (see _analyse_map)
        for slot in sorted({info["slot"] for info in group}):
            g = [info for info in group if info["slot"] == slot]
            lo = [min(info["ref"].shift[ax] for info in g) for ax in range(3)]
            hi = [max(info["ref"].shift[ax] for info in g) for ax in range(3)]
            prefetch.append(dict(name=pname, slot=slot, absolute=g[0]["ref"].absolute,
                                 lo=lo, hi=hi, refs=g))
    if not prefetch:
        raise ValueError("no array to prefetch")
    S = 8 if dtype == "double" else 4
    TYT = TILE_TY * TILE_RY
    off = 0
    for p in prefetch:
        p["ex"] = p["hi"][0] - p["lo"][0] + 1
        p["ey"] = TYT + p["hi"][1] - p["lo"][1]
        p["ez"] = TILE_TZ + p["hi"][2] - p["lo"][2]
        p["ns"] = p["ex"] + 1             # ring slots: the planes in use + the one in flight
        p["off"] = off
        off += p["ns"] * p["ey"] * p["ez"]
        off = (off + 3) // 4 * 4                      # keep 16-byte alignment of the rings
    smem_bytes = off * S
    if smem_bytes > TILE_SMEM_LIMIT:
        raise NotImplementedError(f"the prefetch tiles need {smem_bytes} bytes of shared "
                                  "memory")
    in_tile = {}
    for n, p in enumerate(prefetch):
        for info in p["refs"]:
            in_tile[info["var"]] = (n, p)

    lines = [
        "// generated by pystella_b200.codegen -- do not edit",
        f"typedef {dtype} T;",
        "struct psb_jit_args {",
        f"  void* arrays[{MAX_ARRAYS}]; int halos[{MAX_ARRAYS}];",
        f"  double scalars[{MAX_SCALARS}]; long long Nx, Ny, Nz;",
        "};",
        "#define IDX(h, si, sj, sk) ((((long long)(i + (h) + (si))) * (Ny + 2 * (h)) + "
        "(j + (h) + (sj))) * (Nz + 2 * (h)) + (k + (h) + (sk)))",
        "#define IDXA(h, a, b, c) ((((long long)(i + (a))) * (Ny + 2 * (h)) + "
        "(j + (b))) * (Nz + 2 * (h)) + (k + (c)))",
        f"#define TZ {TILE_TZ}",
        f"#define TY {TILE_TY}",
        f"#define RY {TILE_RY}",
        "#define TYT (TY * RY)",
        "// asynchronous global -> shared copies (LDGSTS): the next plane is in flight while",
        "// the current one is evaluated; plain copies under the CPU execution model",
        "#ifdef PSB_CUDA_EMU",
        "#define PSB_CP_ASYNC(dst, src) (*(dst) = *(src))",
        "#define PSB_CP_WAIT() ((void)0)",
        "#else",
        "__device__ __forceinline__ void psb_cp_async(T* dst, const T* src) {",
        "  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);",
        f'  asm volatile("cp.async.ca.shared.global [%0], [%1], {S};" ::"r"(d), "l"(src) '
        ': "memory");',
        "}",
        "#define PSB_CP_ASYNC(dst, src) psb_cp_async((dst), (src))",
        '#define PSB_CP_WAIT() asm volatile("cp.async.wait_all;" ::: "memory")',
        "#endif",
        f'extern "C" __global__ void __launch_bounds__({TILE_TZ * TILE_TY}) {name}'
        "(const psb_jit_args args) {",
        "  const long long Nx = args.Nx, Ny = args.Ny, Nz = args.Nz;",
        "#ifdef PSB_CUDA_EMU",
        "  T* const smem = (T*)emu::dynamic_smem();",
        "#else",
        "  extern __shared__ __align__(16) unsigned char psb_smem_raw[];",
        "  T* const smem = (T*)psb_smem_raw;",
        "#endif",
        "  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TZ + tx;",
        "  const long long z0 = (long long)blockIdx.x * TZ, j0 = (long long)blockIdx.y * TYT;",
        "  const long long cx = (Nx + gridDim.z - 1) / gridDim.z;   // planes per x chunk",
        "  const long long i0 = (long long)blockIdx.z * cx;",
        "  const long long i1 = (i0 + cx < Nx) ? i0 + cx : Nx;",
        "  if (i0 >= i1) return;                                   // uniform in the CTA",
        "  const long long k = z0 + tx;",
    ]
    for n, sname in enumerate(scalars):
        lines.append(f"  const T s_{sname} = (T)args.scalars[{n}];")
    for n, p in enumerate(prefetch):
        s_ = p["slot"]
        if p["absolute"]:
            ox, oy, oz = (str(v) for v in p["lo"])
        else:
            ox, oy, oz = (f"(args.halos[{s_}] + ({v}))" for v in p["lo"])
        lines += [
            f"  // ring of '{p['name']}': {p['ns']} slots ({p['ex']} planes in use + 1 in "
            f"flight) of {p['ey']} x {p['ez']}",
            f"  T* const ring{n} = smem + {p['off']};",
            f"  const T* const src{n} = (const T*)args.arrays[{s_}];",
            f"  const long long py{n} = Ny + 2 * args.halos[{s_}], "
            f"pz{n} = Nz + 2 * args.halos[{s_}];",
            f"  const long long ox{n} = {ox}, oy{n} = j0 + {oy}, oz{n} = z0 + {oz};",
            f"#define LOAD_PLANE{n}(c) do {{ \\",
            f"    T* const dst_ = ring{n} + ((c) % {p['ns']}) * {p['ey'] * p['ez']}; \\",
            f"    const T* const row_ = src{n} + (i0 + (c) + ox{n}) * py{n} * pz{n}; \\",
            f"    for (int e_ = tid; e_ < {p['ey'] * p['ez']}; e_ += TZ * TY) {{ \\",
            f"      const int jj_ = e_ / {p['ez']}, kk_ = e_ - jj_ * {p['ez']}; \\",
            f"      const long long pj_ = oy{n} + jj_, pk_ = oz{n} + kk_; \\",
            f"      if (pj_ < py{n} && pk_ < pz{n}) "
            f"PSB_CP_ASYNC(dst_ + e_, row_ + pj_ * pz{n} + pk_); \\",
            "    } } while (0)",
            f"  for (int c = 0; c < {p['ex']}; ++c) LOAD_PLANE{n}(c);",
        ]
    lines += ["  PSB_CP_WAIT();",
              "  __syncthreads();",
              "  // q: plane counter within the chunk (int: cheap ring arithmetic)",
              "  const int nq = (int)(i1 - i0);",
              "  for (int q = 0; q < nq; ++q) {",
              "    const long long i = i0 + q;",
              "    if (q + 1 < nq) {        // the plane step q + 1 adds, into the slot step "
              "q - 1 released"]
    for n, p in enumerate(prefetch):
        lines.append(f"      LOAD_PLANE{n}(q + {p['ex']});")
    lines += [
        "    }",
        "    #pragma unroll",
        "    for (int r = 0; r < RY; ++r) {",
        "      const int jl = ty + TY * r;",
        "      const long long j = j0 + jl;",
        "      if (j < Ny && k < Nz) {",
    ]
    for info in refs.values():
        s_, (si, sj, sk) = info["slot"], info["ref"].shift
        macro = "IDXA" if info["ref"].absolute else "IDX"
        addr = f"((T*)args.arrays[{s_}])[{macro}(args.halos[{s_}], {si}, {sj}, {sk})]"
        info["addr"] = addr
        if info["var"] in in_tile:
            n, p = in_tile[info["var"]]
            dx, dy, dz = si - p["lo"][0], sj - p["lo"][1], sk - p["lo"][2]
            lines.append(f"        T {info['var']} = ring{n}[((q + {dx}) % {p['ns']}) * "
                         f"{p['ey'] * p['ez']} + (jl + {dy}) * {p['ez']} + (tx + {dz})];")
        elif info["read"]:
            lines.append(f"        T {info['var']} = {addr};")
        else:
            lines.append(f"        T {info['var']};")
    lines += ["      " + ln for ln in body]
    for info in refs.values():
        if info["written"]:
            lines.append(f"        {info['addr']} = {info['var']};")
    lines += ["      }", "    }",
              "    PSB_CP_WAIT();           // the plane in flight has landed ...",
              "    __syncthreads();         // ... for everybody, and step q's reads are done",
              "  }"]
    for n in range(len(prefetch)):
        lines.append(f"#undef LOAD_PLANE{n}")
    lines.append("}")
    domain = sorted({info["slot"] for info in refs.values()
                     if info["written"] and info["ref"].absolute
                     and not any(info["ref"].shift)})
    ksrc = KernelSource("\n".join(lines) + "\n", name, slots, scalars, dtype, domain)
    ksrc.smem_bytes = smem_bytes
    written_slots = {info["slot"] for info in refs.values() if info["written"]}
    #: running the map twice gives the same result (no array is both read and written)
    ksrc.idempotent = not any(info["read"] and info["slot"] in written_slots
                              for info in refs.values())
    ksrc.prefetch = [dict(name=p["name"], slot=p["slot"], absolute=p["absolute"],
                          lo=tuple(p["lo"]), hi=tuple(p["hi"])) for p in prefetch]
    return ksrc


# ---------------------------------------------------------------------------
# histograms (pystella/histogram.py:60-170)
# ---------------------------------------------------------------------------

HIST_WARPS = 8                 # block = (32, HIST_WARPS): a warp owns one y row at a time
HIST_SMEM_BYTES = 48 * 1024    # static shared memory limit: one copy of the bins must fit
HIST_COPY_BYTES = 24 * 1024    # ... further per-warp copies only up to here (occupancy)
HIST_UNROLL = 4                # z chunks of 32 sites whose loads are issued together


class _Operands:
    """Array / scalar operands of a set of read-only expressions (the binding tables of
    :class:`KernelSource`)."""

    def __init__(self):
        self.slots, self.slot_index = [], {}
        self.refs, self.scalars = {}, []

    def leaf(self, node):
        ref = _as_ref(node)
        if ref is not None:
            if ref.key not in self.refs:
                if ref.slot_key not in self.slot_index:
                    if len(self.slots) >= MAX_ARRAYS - 1:      # the last slot: the histogram
                        raise NotImplementedError("too many distinct arrays in one kernel")
                    self.slot_index[ref.slot_key] = len(self.slots)
                    self.slots.append((ref.name, ref.subscript, ref.field))
                n = len(self.refs)
                self.refs[ref.key] = dict(ref=ref, slot=self.slot_index[ref.slot_key],
                                          base=f"a{n}", var=f"a{n}[u]")
            return self.refs[ref.key]["var"]
        sname = _scalar_name(node)
        if sname is not None:
            if sname not in self.scalars:
                if len(self.scalars) >= MAX_SCALARS:
                    raise NotImplementedError("too many scalars in one generated kernel")
                self.scalars.append(sname)
            return f"s_{sname}"
        return None


def histogram_layout(num_hists, num_bins):
    """-> ``(use_shared, copies)``: block-private copies of all histograms in shared memory
    (one per warp while they fit in 48 KB, at least one), else global atomics."""
    one = num_hists * num_bins * 8
    if one > HIST_SMEM_BYTES:
        return False, 0
    return True, max(1, min(HIST_WARPS, HIST_COPY_BYTES // one))


def generate_histogram_kernel(histograms, num_bins, dtype="double",
                              name="psb_generated_histogram", halo_shape=None):
    """
    ``histograms``: list of ``(bin_expr, weight_expr)``.  One kernel evaluates all of them
    in a single pass over the lattice: ``hist[j, floor(bin_j)] += weight_j`` in float64,
    bins outside ``[0, num_bins)`` (and NaN) dropped.  The histogram array (``num_hists *
    num_bins`` doubles, zeroed by the caller) is passed in array slot ``len(slots)``.
    Accumulation: warp-aggregated runs -> block-private shared-memory copies -> one global
    atomic per bin and block (``csrc/hist_accum.cuh``, inlined into the source).
    """
    hists = list(histograms)
    if not hists:
        raise ValueError("no histograms")
    num_bins, nh = int(num_bins), len(hists)
    if num_bins < 1:
        raise ValueError("num_bins must be positive")
    _HALO_VALUES.clear()
    if halo_shape is not None:
        hs = (halo_shape,) * 3 if isinstance(halo_shape, (int, np.integer)) \
            else tuple(halo_shape)
        _HALO_VALUES.update(h=hs[0], hx=hs[0], hy=hs[1], hz=hs[2])
    ops = _Operands()
    code = [(expr_to_c(b, ops.leaf), expr_to_c(w, ops.leaf)) for b, w in hists]
    use_shared, copies = histogram_layout(nh, num_bins)
    hist_slot = len(ops.slots)
    with open(os.path.join(HERE, "csrc", "hist_accum.cuh")) as fh:
        accum = fh.read().replace("#pragma once\n", "")

    lines = [
        "// generated by pystella_b200.codegen -- do not edit",
        f"typedef {dtype} T;",
        "struct psb_jit_args {",
        f"  void* arrays[{MAX_ARRAYS}]; int halos[{MAX_ARRAYS}];",
        f"  double scalars[{MAX_SCALARS}]; long long Nx, Ny, Nz;",
        "};",
        accum,
        "#define IDX(h, si, sj, sk) ((((long long)(i + (h) + (si))) * (Ny + 2 * (h)) + "
        "(j + (h) + (sj))) * (Nz + 2 * (h)) + (k + (h) + (sk)))",
        "#define IDXA(h, a, b, c) ((((long long)(i + (a))) * (Ny + 2 * (h)) + "
        "(j + (b))) * (Nz + 2 * (h)) + (k + (c)))",
        f"#define NUM_HISTS {nh}",
        f"#define NUM_BINS {num_bins}",
        f"#define HIST_WARPS {HIST_WARPS}",
        f"#define HIST_COPIES {max(copies, 1)}",
        f"#define HIST_UNROLL {HIST_UNROLL}",
        f'extern "C" __global__ void __launch_bounds__({32 * HIST_WARPS}) {name}'
        "(const psb_jit_args args) {",
        "  const long long Nx = args.Nx, Ny = args.Ny, Nz = args.Nz;",
        "  const int lane = threadIdx.x, warp = threadIdx.y;",
        "  const int tid = warp * 32 + lane, nthreads = 32 * HIST_WARPS;",
        f"  double* const hist = (double*)args.arrays[{hist_slot}];",
    ]
    if use_shared:
        lines += [
            "  __shared__ double temp[HIST_COPIES * NUM_HISTS * NUM_BINS];",
            "  for (int b = tid; b < HIST_COPIES * NUM_HISTS * NUM_BINS; b += nthreads) "
            "temp[b] = 0.0;",
            "  __syncthreads();",
            "  double* const mine = temp + (warp % HIST_COPIES) * (NUM_HISTS * NUM_BINS);",
        ]
    else:
        lines.append("  double* const mine = hist;")
    for n, sname in enumerate(ops.scalars):
        lines.append(f"  const T s_{sname} = (T)args.scalars[{n}];")
    lines += [
        "  const long long ngy = (Ny + HIST_WARPS - 1) / HIST_WARPS;",
        "  const long long ngroups = Nx * ngy;",
        "  for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {",
        "    const long long i = g / ngy;",
        "    const long long j = (g - i * ngy) * HIST_WARPS + warp;",
        "    if (j >= Ny) continue;                 // uniform in the warp",
        "    for (long long k0 = 0; k0 < Nz; k0 += 32 * HIST_UNROLL) {",
        "      bool valid[HIST_UNROLL];",
    ]
    for info in ops.refs.values():
        lines.append(f"      T {info['base']}[HIST_UNROLL];")
    lines += [
        "      #pragma unroll",
        "      for (int u = 0; u < HIST_UNROLL; ++u) {      // all loads first",
        "        const long long k = k0 + 32 * u + lane;",
        "        valid[u] = k < Nz;",
        "        if (valid[u]) {",
    ]
    for info in ops.refs.values():
        s_, (si, sj, sk) = info["slot"], info["ref"].shift
        macro = "IDXA" if info["ref"].absolute else "IDX"
        lines.append(f"          {info['var']} = ((const T*)args.arrays[{s_}])"
                     f"[{macro}(args.halos[{s_}], {si}, {sj}, {sk})];")
    lines += [
        "        }",
        "      }",
        "      #pragma unroll",
        "      for (int u = 0; u < HIST_UNROLL; ++u) {",
        "        if (k0 + 32 * u >= Nz) break;              // uniform in the warp",
    ]
    for j in range(nh):
        lines += [f"        int bin{j} = -1;", f"        double w{j} = 0.0;"]
    lines.append("        if (valid[u]) {")
    for j, (bcode, wcode) in enumerate(code):
        lines += [
            f"          const T q{j} = floor({bcode});",
            f"          if (q{j} >= (T)0 && q{j} < (T)NUM_BINS) bin{j} = (int)q{j};",
            f"          w{j} = (double)({wcode});",
        ]
    lines.append("        }")
    for j in range(nh):
        lines.append(f"        psb::hist_accumulate(mine + {j} * NUM_BINS, NUM_BINS, bin{j}, "
                     f"w{j}, lane);")
    lines.append("      }")
    lines += ["    }", "  }"]
    if use_shared:
        lines += ["  __syncthreads();",
                  "  psb::hist_flush(temp, HIST_COPIES, NUM_HISTS * NUM_BINS, hist, tid, "
                  "nthreads);"]
    lines.append("}")
    ksrc = KernelSource("\n".join(lines) + "\n", name, ops.slots, ops.scalars, dtype)
    ksrc.hist_slot = hist_slot
    ksrc.num_hists, ksrc.num_bins = nh, num_bins
    return ksrc


# ---------------------------------------------------------------------------
# runtime
# ---------------------------------------------------------------------------

class JitArgs(C.Structure):
    _fields_ = [("arrays", C.c_void_p * MAX_ARRAYS), ("halos", C.c_int32 * MAX_ARRAYS),
                ("scalars", C.c_double * MAX_SCALARS),
                ("Nx", C.c_int64), ("Ny", C.c_int64), ("Nz", C.c_int64)]


class JitLibrary:
    """ctypes binding of ``libpystella_b200_jit.so``."""

    _lib = None

    @classmethod
    def get(cls):
        if cls._lib is None:
            from pystella_b200._lib import PsbError
            path = os.path.join(HERE, "_C", "libpystella_b200_jit.so")
            if not os.path.exists(path):
                raise PsbError(f"{path} not found: python -m pystella_b200.build")
            lib = C.CDLL(path)
            lib.psb_jit_compile.restype = C.c_int64
            lib.psb_jit_compile.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
            lib.psb_jit_last_log.restype = C.c_char_p
            lib.psb_jit_cubin_size.restype = C.c_size_t
            lib.psb_jit_cubin_size.argtypes = [C.c_int64]
            lib.psb_jit_launch.restype = C.c_int
            lib.psb_jit_launch.argtypes = [C.c_int64, C.POINTER(JitArgs)] + [C.c_uint] * 6 \
                + [C.c_void_p]
            lib.psb_jit_release.restype = C.c_int
            lib.psb_jit_release.argtypes = [C.c_int64]
            lib.psb_jit_compile_ex.restype = C.c_int64
            lib.psb_jit_compile_ex.argtypes = [
                C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p),
                C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int]
            lib.psb_jit_kernel_symbol.restype = C.c_char_p
            lib.psb_jit_kernel_symbol.argtypes = [C.c_int64]
            lib.psb_jit_write_cubin.restype = C.c_int
            lib.psb_jit_write_cubin.argtypes = [C.c_int64, C.c_char_p]
            lib.psb_jit_load_cubin.restype = C.c_int64
            lib.psb_jit_load_cubin.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
            lib.psb_jit_launch_params.restype = C.c_int
            lib.psb_jit_launch_params.argtypes = [C.c_int64, C.POINTER(C.c_void_p)] \
                + [C.c_uint] * 7 + [C.c_void_p]
            lib.psb_jit_launch_params_v.restype = C.c_int
            lib.psb_jit_launch_params_v.argtypes = [C.c_int64, C.POINTER(C.c_void_p)] \
                + [C.c_uint] * 7 + [C.c_int, C.c_void_p]
            cls._lib = lib
        return cls._lib

    # -- on-disk cache of compiled kernels ----------------------------------------------
    # key = sha256(source, kernel name, options, NVRTC version, header digests): a hit skips
    # NVRTC entirely (a generated fused stage takes ~1.5 s to compile, a map ~0.3 s)
    @staticmethod
    def cache_dir():
        d = os.environ.get("PSB_JIT_CACHE")
        if d is None:
            d = os.path.join(os.path.expanduser("~"), ".cache", "pystella_b200", "jit")
        if d in ("", "0", "off"):
            return None
        try:
            os.makedirs(d, exist_ok=True)
        except OSError:
            return None
        return d

    @classmethod
    def _cache_key(cls, source, name, options, extra_files=()):
        import hashlib
        h = hashlib.sha256()
        for part in (source, name, repr(list(options))):
            h.update(part.encode() if isinstance(part, str) else part)
            h.update(b"\0")
        lib_path = os.path.join(HERE, "_C", "libpystella_b200_jit.so")
        for path in list(extra_files) + [lib_path]:
            try:
                with open(path, "rb") as fh:
                    h.update(hashlib.sha256(fh.read()).digest())
            except OSError:
                h.update(b"?")
        return h.hexdigest()

    @classmethod
    def cached(cls, source, name, options, extra_files, build):
        """``build() -> handle`` unless a cubin for this key is on disk; stores new cubins."""
        lib = cls.get()
        d = cls.cache_dir()
        if d is None:
            return build()
        key = cls._cache_key(source, name, options, extra_files)
        cubin_path, sym_path = os.path.join(d, key + ".cubin"), os.path.join(d, key + ".sym")
        if os.path.exists(cubin_path) and os.path.exists(sym_path):
            try:
                blob = open(cubin_path, "rb").read()
                sym = open(sym_path, "rb").read()
                h = lib.psb_jit_load_cubin(blob, len(blob), sym)
                if h != 0:
                    return h
            except OSError:
                pass
        h = build()
        try:
            tmp = cubin_path + f".{os.getpid()}.tmp"
            if lib.psb_jit_write_cubin(h, tmp.encode()) == 0:
                sym = lib.psb_jit_kernel_symbol(h)
                with open(sym_path + f".{os.getpid()}.tmp", "wb") as fh:
                    fh.write(sym)
                os.replace(sym_path + f".{os.getpid()}.tmp", sym_path)
                os.replace(tmp, cubin_path)
        except OSError:
            pass
        return h

    @classmethod
    def compile(cls, ksrc, arch="sm_100a"):
        from pystella_b200._lib import PsbError
        lib = cls.get()

        def build():
            h = lib.psb_jit_compile(ksrc.source.encode(), ksrc.name.encode(), arch.encode())
            if h == 0:
                raise PsbError("NVRTC compilation of a generated kernel failed:\n"
                               + lib.psb_jit_last_log().decode(errors="replace")
                               + "\n--- source ---\n" + ksrc.source)
            return h
        return cls.cached(ksrc.source, ksrc.name, [arch], (), build)


def infer_layout(ksrc, view_shapes, rank_shape, halo_of):
    """
    -> ``(rank_shape, [halo per array slot])`` for arrays of the given 3-D shapes.  The
    domain is ``rank_shape`` if given, else the interior of the first Field operand, else
    the shape of a Field-less array written at offset (0, 0, 0); the padding of Field-less
    arrays is inferred from their shapes.
    """
    shape = tuple(rank_shape) if rank_shape is not None else None
    if shape is None:
        for vs, (name, sub, field) in zip(view_shapes, ksrc.slots):
            if field is not None:
                h = halo_of(field)
                shape = tuple(s - 2 * h for s in vs)
                break
    if shape is None and ksrc.domain_slots:
        shape = tuple(view_shapes[ksrc.domain_slots[0]])
    if shape is None:
        raise ValueError("cannot infer the domain: pass rank_shape=")
    halos = []
    for vs, (name, sub, field) in zip(view_shapes, ksrc.slots):
        if field is not None:
            h = halo_of(field)
        else:
            h, rem = divmod(vs[0] - shape[0], 2)
            if rem or h < 0:
                raise ValueError(f"'{name}' has shape {tuple(vs)}, not the domain "
                                 f"{shape} plus symmetric padding")
        if tuple(s - 2 * h for s in vs) != shape:
            raise ValueError(f"'{name}' has rank shape {tuple(s - 2 * h for s in vs)}, "
                             f"expected {shape}")
        halos.append(h)
    return shape, halos


class GenericMap:
    """
    A compiled element-wise map: the run-time generated counterpart of
    ``pystella.ElementWiseMap`` (``elementwise.py:237-353``).  Call with the arrays
    and scalars by keyword.
    """

    def __init__(self, map_instructions, tmp_instructions=None, dtype="float64",
                 halo_shape=None, rank_shape=None):
        self.dtype = np.dtype(dtype)
        ctype = {"float64": "double", "float32": "float"}[self.dtype.name]
        self.ksrc = generate_kernel(map_instructions, tmp_instructions, ctype,
                                    halo_shape=halo_shape)
        self.halo_shape = halo_shape
        self.rank_shape = tuple(rank_shape) if rank_shape is not None else None
        self._handle = None

    def _halo_of(self, field):
        from pystella_b200.field import evaluate
        off = field.base_offset[0] if field.base_offset else 0
        if is_constant(off):
            return int(off)
        if self.halo_shape is None:
            raise ValueError("pass halo_shape= for arrays with a symbolic offset")
        h = self.halo_shape if not isinstance(self.halo_shape, (tuple, list)) \
            else self.halo_shape[0]
        return int(evaluate(off, {"h": h, "hx": h, "hy": h, "hz": h}))

    def __call__(self, queue=None, filter_args=False, **kwargs):
        import torch
        from pystella_b200 import _lib
        if self._handle is None:
            self._handle = JitLibrary.compile(self.ksrc)
        args = JitArgs()
        shape = self.rank_shape
        want = getattr(torch, self.dtype.name)
        views = []
        for name, sub, field in self.ksrc.slots:
            if name not in kwargs:
                raise TypeError(f"missing array argument '{name}'")
            t = _lib.require_cuda_tensor(kwargs[name], name)
            if t.dtype != want:
                raise ValueError(f"'{name}' has dtype {t.dtype}; this map was built for "
                                 f"{self.dtype.name}")
            view = t[sub] if sub else t
            if view.dim() != 3:
                raise ValueError(f"'{name}{list(sub)}' is not a 3-D array")
            if not view.is_contiguous():
                raise ValueError(f"'{name}{list(sub)}' is not contiguous (z fastest)")
            views.append(view)
        shape, halos = infer_layout(self.ksrc, [tuple(v.shape) for v in views], shape,
                                    self._halo_of)
        for n, view in enumerate(views):
            args.arrays[n] = view.data_ptr()
            args.halos[n] = halos[n]
        for n, sname in enumerate(self.ksrc.scalars):
            base, _, idx = sname.partition("__")
            if base not in kwargs:
                raise TypeError(f"missing scalar argument '{base}'")
            v = kwargs[base]
            v = v.cpu().numpy() if isinstance(v, torch.Tensor) else v
            v = np.asarray(v)
            if idx:
                v = v[tuple(int(i) for i in idx.split("_"))]
            args.scalars[n] = float(v.reshape(-1)[0])
        args.Nx, args.Ny, args.Nz = shape
        lib = JitLibrary.get()
        rc = lib.psb_jit_launch(self._handle, C.byref(args), (shape[2] + 63) // 64,
                                (shape[1] + 3) // 4, shape[0], 64, 4, 1,
                                _lib.stream_ptr(queue))
        if rc != 0:
            raise _lib.PsbError("launch of a generated kernel failed: "
                                + lib.psb_jit_last_log().decode(errors="replace"))
        return None, {}


def _bind_operands(ksrc, kwargs, dtype, rank_shape, halo_of, args):
    """Fill array pointers, halos, scalars and the domain of ``args`` from keyword
    arguments (shared by the generated maps and histograms); -> rank shape."""
    import torch
    from pystella_b200 import _lib
    want = getattr(torch, dtype.name)
    views = []
    for name, sub, field in ksrc.slots:
        if name not in kwargs:
            raise TypeError(f"missing array argument '{name}'")
        t = _lib.require_cuda_tensor(kwargs[name], name)
        if t.dtype != want:
            raise ValueError(f"'{name}' has dtype {t.dtype}; this kernel was built for "
                             f"{dtype.name}")
        view = t[sub] if sub else t
        if view.dim() != 3:
            raise ValueError(f"'{name}{list(sub)}' is not a 3-D array")
        if not view.is_contiguous():
            raise ValueError(f"'{name}{list(sub)}' is not contiguous (z fastest)")
        views.append(view)
    shape, halos = infer_layout(ksrc, [tuple(v.shape) for v in views], rank_shape, halo_of)
    for n, view in enumerate(views):
        args.arrays[n] = view.data_ptr()
        args.halos[n] = halos[n]
    for n, sname in enumerate(ksrc.scalars):
        base, _, idx = sname.partition("__")
        if base not in kwargs:
            if base == "pi":
                args.scalars[n] = float(np.pi)
                continue
            raise TypeError(f"missing scalar argument '{base}'")
        v = kwargs[base]
        v = v.cpu().numpy() if isinstance(v, torch.Tensor) else v
        v = np.asarray(v)
        if idx:
            v = v[tuple(int(i) for i in idx.split("_"))]
        args.scalars[n] = float(v.reshape(-1)[0])
    args.Nx, args.Ny, args.Nz = shape
    return shape, views


class GeneratedHistogram:
    """
    The compiled kernel behind :class:`pystella_b200.Histogrammer` (the reference builds
    its histogram kernels with loopy, ``pystella/histogram.py:60-170``): all histograms of
    one ``Histogrammer`` in a single pass.  ``__call__`` returns a float64 CUDA tensor
    ``(num_hists, num_bins)`` on the launch stream.
    """

    def __init__(self, histograms, num_bins, dtype="float64", halo_shape=None,
                 rank_shape=None):
        self.dtype = np.dtype(dtype)
        ctype = {"float64": "double", "float32": "float"}[self.dtype.name]
        self.ksrc = generate_histogram_kernel(histograms, num_bins, ctype,
                                              halo_shape=halo_shape)
        self.halo_shape = halo_shape
        self.rank_shape = tuple(rank_shape) if rank_shape is not None else None
        self._handle = None

    _halo_of = GenericMap._halo_of

    def __call__(self, queue=None, blocks=0, **kwargs):
        import contextlib
        import torch
        from pystella_b200 import _lib
        if self._handle is None:
            self._handle = JitLibrary.compile(self.ksrc)
        if not self.ksrc.slots and self.rank_shape is None:
            raise ValueError("no array argument: pass rank_shape= to the constructor")
        args = JitArgs()
        shape, views = _bind_operands(self.ksrc, kwargs, self.dtype, self.rank_shape,
                                      self._halo_of, args)
        device = views[0].device if views else torch.device("cuda",
                                                            torch.cuda.current_device())
        nh, nb = self.ksrc.num_hists, self.ksrc.num_bins
        with torch.cuda.stream(queue) if queue is not None else contextlib.nullcontext():
            hist = torch.zeros((nh, nb), dtype=torch.float64, device=device)
        args.arrays[self.ksrc.hist_slot] = hist.data_ptr()
        ngroups = shape[0] * ((shape[1] + HIST_WARPS - 1) // HIST_WARPS)
        if blocks <= 0:
            blocks = 4 * torch.cuda.get_device_properties(device).multi_processor_count
        blocks = max(1, min(int(blocks), ngroups))
        lib = JitLibrary.get()
        rc = lib.psb_jit_launch(self._handle, C.byref(args), blocks, 1, 1, 32, HIST_WARPS, 1,
                                _lib.stream_ptr(queue))
        if rc != 0:
            raise _lib.PsbError("launch of a generated histogram kernel failed: "
                                + lib.psb_jit_last_log().decode(errors="replace"))
        return hist


class TiledStencilMap(GenericMap):
    """
    A compiled stencil map whose ``prefetch_args`` are staged in shared memory
    (:func:`generate_stencil_kernel`): what ``pystella.Stencil`` / ``StreamingStencil``
    ask loopy for (``stencil.py:66-84, 113-133``).  Same call contract as
    :class:`GenericMap`.
    """

    PLANES_PER_CHUNK = 64

    def __init__(self, map_instructions, tmp_instructions=None, dtype="float64",
                 halo_shape=None, rank_shape=None, prefetch_args=()):
        self.dtype = np.dtype(dtype)
        ctype = {"float64": "double", "float32": "float"}[self.dtype.name]
        self.ksrc = generate_stencil_kernel(map_instructions, tmp_instructions, ctype,
                                            halo_shape=halo_shape,
                                            prefetch_args=prefetch_args)
        self.halo_shape = halo_shape
        self.rank_shape = tuple(rank_shape) if rank_shape is not None else None
        self._handle = None

    @staticmethod
    def check_reach(prefetch, halos):
        """The stencil offsets of a staged array must stay inside its padding."""
        for p in prefetch:
            H = halos[p["slot"]]
            lo, hi = ((0, 2 * H) if p["absolute"] else (-H, H))
            if min(p["lo"]) < lo or max(p["hi"]) > hi:
                raise ValueError(
                    f"'{p['name']}' is read at offsets {p['lo']}..{p['hi']}, outside its "
                    f"padding of {H}")

    @classmethod
    def launch_grid(cls, shape, sms):
        """-> (gx, gy, gz): (8 x 64) columns in (y, z), x cut into chunks of ~64 planes
        (more, shorter chunks while the grid would not fill the GPU twice over)."""
        gx = (shape[2] + TILE_TZ - 1) // TILE_TZ
        gy = (shape[1] + TILE_TY * TILE_RY - 1) // (TILE_TY * TILE_RY)
        gz = max(1, (shape[0] + cls.PLANES_PER_CHUNK - 1) // cls.PLANES_PER_CHUNK)
        while gx * gy * gz < 2 * sms and shape[0] // (2 * gz) >= 8:
            gz *= 2
        return gx, gy, min(gz, 65535)

    def __call__(self, queue=None, filter_args=False, **kwargs):
        import torch
        from pystella_b200 import _lib
        if self._handle is None:
            self._handle = JitLibrary.compile(self.ksrc)
        args = JitArgs()
        shape, views = _bind_operands(self.ksrc, kwargs, self.dtype, self.rank_shape,
                                      self._halo_of, args)
        self.check_reach(self.ksrc.prefetch, [args.halos[n]
                                              for n in range(len(self.ksrc.slots))])
        sms = torch.cuda.get_device_properties(views[0].device).multi_processor_count
        gx, gy, gz = self.launch_grid(shape, sms)
        lib = JitLibrary.get()
        params = (C.c_void_p * 1)(C.addressof(args))
        rc = lib.psb_jit_launch_params(self._handle, params, gx, gy, gz, TILE_TZ, TILE_TY, 1,
                                       self.ksrc.smem_bytes, _lib.stream_ptr(queue))
        if rc != 0:
            raise _lib.PsbError("launch of a generated stencil kernel failed: "
                                + lib.psb_jit_last_log().decode(errors="replace"))
        return None, {}
