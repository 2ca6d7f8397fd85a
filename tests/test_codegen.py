"""
Run-time code generation (experimental JIT path), everything that needs no GPU:
the expression -> C translation is checked numerically by compiling the generated
expression as a HOST function with gcc, and generated kernels are compiled with NVRTC
for sm_100a through libpystella_b200_jit.so.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

import pystella_b200 as ps
from pystella_b200 import codegen
from pystella_b200.field import evaluate, var

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host_eval(exprs, names, values):
    """Compile `exprs` (over Variables `names`) into host C and evaluate at `values`."""
    def leaf(node):
        n = codegen._scalar_name(node)
        return f"v[{names.index(n)}]" if n in names else None
    body = "\n".join(f"  out[{i}] = {codegen.expr_to_c(e, leaf)};" for i, e in enumerate(exprs))
    src = ("#include <math.h>\ntypedef double T;\n"
           "void f(const double* v, double* out) {\n" + body + "\n}\n")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "e.c")
        so = os.path.join(d, "e.so")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-O1", "-ffp-contract=off", "-shared", "-fPIC", c,
                               "-o", so, "-lm"])
        lib = ctypes.CDLL(so)
        v = (ctypes.c_double * len(values))(*values)
        out = (ctypes.c_double * len(exprs))()
        lib.f(v, out)
        return list(out)


def test_expression_translation_matches_evaluate():
    x, y, z = var("x"), var("y"), var("z")
    fabs, sqrt, cos, exp = var("fabs"), var("sqrt"), var("cos"), var("exp")
    exprs = [
        x + 2 * y - z / 3,
        (x - y)**2 * z + x**3 - y**-2,
        x**2.5 + y**z,
        fabs(x - y) * sqrt(z) + cos(x * y) - exp(-z),
        -(x * y) / (1 + z**2) + 0.31,
        ((x + y) * (y + z) * (z + x))**4,
    ]
    names = ["x", "y", "z"]
    rng = np.random.default_rng(0)
    for _ in range(3):
        vals = list(rng.random(3) + 0.5)
        got = _host_eval(exprs, names, vals)
        ctx = dict(zip(names, vals))
        for e, g in zip(exprs, got):
            assert np.isclose(g, evaluate(e, ctx), rtol=1e-14), str(e)


def _lsrk_statements(rhs_dict, A, B):
    """The reference's make_steps (step.py:473-491) with this package's Fields."""
    from pystella_b200.step import gen_tmp_name
    tmps, maps = [], []
    for i, (key, val) in enumerate(rhs_dict.items()):
        tmps.append((var(gen_tmp_name(key, suffix=f"_rhs_{i}")), val))
    for i, key in enumerate(rhs_dict):
        fld = key.aggregate if hasattr(key, "aggregate") else key
        twin = ps.Field(gen_tmp_name(fld), offset=fld.offset, shape=fld.shape,
                        indices=fld.indices)
        k = twin[key.index] if hasattr(key, "aggregate") else twin
        rhs = var(gen_tmp_name(key, suffix=f"_rhs_{i}"))
        maps.append((k, A * k + var("dt") * rhs))
        maps.append((key, key + B * k))
    return maps, tmps


def test_generated_kernel_structure():
    f = ps.DynamicField("f", offset="h")
    maps, tmps = _lsrk_statements({f: f.dot, f.dot: f.lap}, -0.4, 0.3)
    k = codegen.generate_kernel(maps, tmps)
    assert [s[0] for s in k.slots] == ["dfdt", "lap_f", "_f_tmp", "f", "_dfdt_tmp"]
    assert k.scalars == ["dt"]
    src = k.source
    # every array element is loaded once and the four unknown/twin arrays stored once
    assert src.count("= ((T*)args.arrays[") == 5
    assert sum(1 for ln in src.splitlines() if ln.strip().startswith("((T*)args.arrays[")) == 4
    # the update of f uses the NEW twin (statement order), not a re-load
    assert "a3 = (a3 + ((T)(0.2" in src.replace("0.29999999999999999", "0.2")

    # hazards are refused
    g = ps.Field("g", offset="h")
    from pystella_b200.field import shift_fields
    with pytest.raises(NotImplementedError):
        codegen.generate_kernel([(g, shift_fields(g, (1, 0, 0)) + 1)])
    with pytest.raises(ValueError):
        codegen.generate_kernel([(var("x"), g)])


def _compile(ksrc):
    h = codegen.JitLibrary.compile(ksrc)
    assert codegen.JitLibrary.get().psb_jit_cubin_size(h) > 1000
    codegen.JitLibrary.get().psb_jit_release(h)


def test_nvrtc_compiles_generated_kernels():
    """NVRTC needs no GPU: the generated sources must compile for sm_100a."""
    # wave equation, one low-storage stage
    f = ps.DynamicField("f", offset="h")
    maps, tmps = _lsrk_statements({f: f.dot, f.dot: f.lap}, -0.4, 0.3)
    _compile(codegen.generate_kernel(maps, tmps, "double"))
    # two scalars with a NON-polynomial potential (outside the hand-written menu), fp32
    g = ps.DynamicField("g", offset="h", shape=(2,))
    cos = var("cos")
    sec = ps.ScalarSector(2, f=g, potential=lambda p: 1 - cos(p[0]) + p[0]**2 * p[1]**2 / 2)
    H, a = ps.Field("hubble", indices=[]), ps.Field("a", indices=[])
    sin = var("sin")
    rhs = {g[0]: g.dot[0], g.dot[0]: g.lap[0] - 2 * H * g.dot[0] - a**2 * (sin(g[0]) + g[0] * g[1]**2),
           g[1]: g.dot[1], g.dot[1]: g.lap[1] - 2 * H * g.dot[1] - a**2 * (g[0]**2 * g[1])}
    maps, tmps = _lsrk_statements(rhs, -1.2, 0.7)
    k = codegen.generate_kernel(maps, tmps, "float")
    assert set(k.scalars) == {"hubble", "a", "dt"} and sec.nscalars == 2
    _compile(k)
    # a stencil map: 2nd-order Laplacian + gradient-squared into separate arrays
    y = ps.Field("y", offset="h")
    lap = sum(ps.SecondCenteredDifference(1)(y, d + 1) * (1 / 0.1**2) for d in range(3))
    gsq = sum((ps.FirstCenteredDifference(1)(y, d + 1) * (1 / 0.1))**2 for d in range(3))
    k = codegen.generate_kernel([(ps.Field("lap"), lap), (ps.Field("gsq"), gsq)])
    assert len(k.slots) == 3 and k.scalars == []
    _compile(k)
    # compiler errors surface with the log
    bad = codegen.KernelSource("extern \"C\" __global__ void k() { undeclared = 1; }", "k",
                               [], [], "double")
    with pytest.raises(ps.PsbError) as err:
        codegen.JitLibrary.compile(bad)
    assert "undeclared" in str(err.value)


def _run_generated_on_host(ksrc, arrays, halos, scalars, shape):
    """
    Execute a generated kernel on the HOST: the CUDA source is turned into plain C by
    replacing the thread-index prologue with loop indices (test-only transformation), so
    that indexing, halos, shifts and statement order can be checked against NumPy here.
    """
    src = ksrc.source
    head = f'extern "C" __global__ void __launch_bounds__(256) {ksrc.name}(const psb_jit_args args) {{'
    assert head in src
    src = src.replace(head, "static void site(const psb_jit_args args, long long i, "
                            "long long j, long long k) {")
    for line in ("  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;\n",
                 "  const long long j = (long long)blockIdx.y * blockDim.y + threadIdx.y;\n",
                 "  const long long i = blockIdx.z;\n"):
        assert line in src
        src = src.replace(line, "")
    src = src.replace("struct psb_jit_args {", "typedef struct {").replace(
        "  double scalars[32]; long long Nx, Ny, Nz;\n};",
        "  double scalars[32]; long long Nx, Ny, Nz;\n} psb_jit_args;")
    src = "#include <math.h>\n" + src + (
        "void run(const psb_jit_args* a) {\n"
        "  for (long long i = 0; i < a->Nx; ++i) for (long long j = 0; j < a->Ny; ++j)\n"
        "    for (long long k = 0; k < a->Nz; ++k) site(*a, i, j, k);\n}\n")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "k.c")
        so = os.path.join(d, "k.so")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-O1", "-ffp-contract=off", "-shared", "-fPIC", c, "-o",
                               so, "-lm"])
        lib = ctypes.CDLL(so)
        args = codegen.JitArgs()
        for n, (name, sub, _) in enumerate(ksrc.slots):
            view = arrays[name][sub] if sub else arrays[name]
            assert view.flags["C_CONTIGUOUS"]
            args.arrays[n] = view.ctypes.data
            args.halos[n] = halos[name]
        for n, s in enumerate(ksrc.scalars):
            args.scalars[n] = scalars[s]
        args.Nx, args.Ny, args.Nz = shape
        lib.run(ctypes.byref(args))


def test_generated_lsrk_stage_semantics_on_host():
    """The generated stage == the oracle's scalar-sector stage (two fields, potential)."""
    import oracle as O
    rng = np.random.default_rng(3)
    N, h = (6, 5, 8), 2
    pad = tuple(n + 2 * h for n in N)
    f = rng.random((2,) + pad); d = rng.random((2,) + pad)
    kf = rng.random((2,) + pad); kd = rng.random((2,) + pad)
    lap = rng.random((2,) + N)
    terms = [(0.5, (2, 0)), (0.5, (0, 2)), (0.5, (2, 2))]
    sec = ps.ScalarSector(2, potential=lambda g: g[0]**2 / 2 + g[1]**2 / 2
                          + g[0]**2 * g[1]**2 / 2)
    dt, a, H = 0.01, 1.3, 0.2
    stepper = ps.LowStorageRK54(sec, dt=dt, halo_shape=h, jit=True)
    A, B = stepper._A[1], stepper._B[1]
    maps, tmps = stepper._jit_statements(1)          # what the product hands the generator
    ksrc = codegen.generate_kernel(maps, tmps)
    arrays = {"f": f.copy(), "dfdt": d.copy(), "_f_tmp": kf.copy(), "_dfdt_tmp": kd.copy(),
              "lap_f": lap}
    halos = {"f": h, "dfdt": h, "_f_tmp": h, "_dfdt_tmp": h, "lap_f": 0}
    _run_generated_on_host(ksrc, arrays, halos, {"dt": dt, "a": a, "hubble": H}, N)

    fr, dr, kfr, kdr = f.copy(), d.copy(), kf.copy(), kd.copy()
    fi = [O.interior(fr[i], h) for i in range(2)]
    di = [O.interior(dr[i], h) for i in range(2)]
    rhs = [np.array(r) for r in O.scalar_rhs(fi, di, [lap[0], lap[1]], a, H, terms)]
    unk, tmp = [], []
    for i in range(2):
        unk += [fi[i], di[i]]
        tmp += [O.interior(kfr[i], h), O.interior(kdr[i], h)]
    O.lsrk_stage(A, B, dt, unk, tmp, rhs)
    for got, ref in ((arrays["f"], fr), (arrays["dfdt"], dr), (arrays["_f_tmp"], kfr),
                     (arrays["_dfdt_tmp"], kdr)):
        assert np.allclose(got, ref, rtol=1e-14, atol=1e-15)     # halos untouched too


def test_generated_stencil_map_on_host():
    """Shifted reads (custom stencils) index correctly: == the oracle's lap / gradient."""
    import oracle as O
    rng = np.random.default_rng(4)
    N, h = (5, 6, 7), 2
    dx = (0.1, 0.2, 0.3)
    y = ps.Field("y", offset="h")
    lap_expr = sum(ps.SecondCenteredDifference(h)(y, d + 1) * dx[d]**-2 for d in range(3))
    pdz_expr = ps.FirstCenteredDifference(h)(y, 3) * (1 / dx[2])
    ksrc = codegen.generate_kernel([(ps.Field("lap"), lap_expr), (ps.Field("pdz"), pdz_expr)])
    yy = rng.random(tuple(n + 2 * h for n in N))
    arrays = {"y": yy, "lap": np.zeros(N), "pdz": np.zeros(N)}
    _run_generated_on_host(ksrc, arrays, {"y": h, "lap": 0, "pdz": 0}, {}, N)
    assert np.allclose(arrays["lap"], O.laplacian(yy, h, dx), rtol=1e-13, atol=1e-10)
    assert np.allclose(arrays["pdz"], O.partial(yy, h, dx, 2), rtol=1e-13, atol=1e-12)


class _OneSided:
    """A custom (non-centred) first-derivative stencil: (-3f(x) + 4f(x+1) - f(x+2)) / 2."""

    def __call__(self, f, direction):
        from pystella_b200.derivs import expand_stencil
        def off(s):
            o = [0, 0, 0]
            o[direction - 1] = s
            return tuple(o)
        return expand_stencil(f, {off(0): -1.5, off(1): 2.0, off(2): -0.5})


def _one_sided_numpy(yy, h, N, axis, dxi):
    def sl(s):
        idx = [slice(h, h + n) for n in N]
        idx[axis] = slice(h + s, h + s + N[axis])
        return yy[tuple(idx)]
    return (-1.5 * sl(0) + 2.0 * sl(1) - 0.5 * sl(2)) / dxi


def test_custom_stencil_differencer_kernels_on_host():
    """FiniteDifferencer(first_stencil=custom, jit=True) generates the reference's kernels
    (derivs.py:254-263, 291-337) from the callable; run them on the host."""
    import oracle as O
    rng = np.random.default_rng(5)
    N, h = (6, 5, 4), 2
    dx = (0.1, 0.2, 0.3)
    decomp = ps.DomainDecomposition((1, 1, 1), h, N)
    with pytest.raises(NotImplementedError, match="custom stencils"):
        ps.FiniteDifferencer(decomp, h, dx, first_stencil=_OneSided(), jit=False)
    fd = ps.FiniteDifferencer(decomp, h, dx, first_stencil=_OneSided())   # default: generated
    yy = rng.random(tuple(n + 2 * h for n in N))

    gm = fd._generated[("lap", "pdx", "pdy", "pdz")]
    arrays = {"fx": yy, **{n: np.zeros(N) for n in ("lap", "pdx", "pdy", "pdz")}}
    halos = {"fx": h, "lap": 0, "pdx": 0, "pdy": 0, "pdz": 0}
    _run_generated_on_host(gm.ksrc, arrays, halos, {}, N)
    assert np.allclose(arrays["lap"], O.laplacian(yy, h, dx), rtol=1e-13, atol=1e-10)
    for ax, n in enumerate(("pdx", "pdy", "pdz")):
        assert np.allclose(arrays[n], _one_sided_numpy(yy, h, N, ax, dx[ax]),
                           rtol=1e-13, atol=1e-11), n

    # the accumulate kernels of divergence (derivs.py:321-337)
    gm = fd._generated[("pdy", "incr")]
    div0 = rng.random(N)
    arrays = {"fx": yy, "div": div0.copy()}
    _run_generated_on_host(gm.ksrc, arrays, {"fx": h, "div": 0}, {}, N)
    assert np.allclose(arrays["div"], div0 + _one_sided_numpy(yy, h, N, 1, dx[1]),
                       rtol=1e-13, atol=1e-11)


@pytest.mark.parametrize("name", ["RungeKutta4", "RungeKutta3Heun", "RungeKutta3Nystrom",
                                  "RungeKutta3Ralston", "RungeKutta3SSP",
                                  "RungeKutta2Midpoint", "RungeKutta2Heun",
                                  "RungeKutta2Ralston"])
def test_generated_classical_rk_stages_on_host(name):
    """Classical RK through the generator (step.py:181-239): copies axis first, the RHS
    reads copy q, `lap_f` ignores the prepend (field/__init__.py:253-257)."""
    import oracle as O
    from pystella_b200.field import var
    rng = np.random.default_rng(6)
    N, h = (4, 5, 6), 1
    pad = tuple(n + 2 * h for n in N)
    dt = 0.05
    f_ = ps.DynamicField("f", offset="h")
    sin = var("sin")
    stepper = getattr(ps, name)({f_: f_.dot, f_.dot: f_.lap - sin(f_)}, dt=dt, halo_shape=h,
                                jit=True)
    nc = stepper.num_copies
    f, d = rng.random((nc,) + pad), rng.random((nc,) + pad)
    fr, dr = f.copy(), d.copy()
    for stage in range(stepper.num_stages):
        lap = rng.random(N)
        maps, tmps = stepper._jit_statements(stage)
        ksrc = codegen.generate_kernel(maps, tmps)
        arrays = {"f": f, "dfdt": d, "lap_f": lap}
        _run_generated_on_host(ksrc, arrays, {"f": h, "dfdt": h, "lap_f": 0}, {"dt": dt}, N)

        q = O.rk_rhs_copy(stage)
        fi = np.stack([O.interior(fr[c], h) for c in range(nc)])
        di = np.stack([O.interior(dr[c], h) for c in range(nc)])
        rhs_f, rhs_d = di[q].copy(), lap - np.sin(fi[q])
        O.rk_stage(name, stage, dt, fi, rhs_f)
        O.rk_stage(name, stage, dt, di, rhs_d)
        for c in range(nc):
            O.interior(fr[c], h)[...] = fi[c]
            O.interior(dr[c], h)[...] = di[c]
        assert np.allclose(f, fr, rtol=1e-14, atol=1e-15), (name, stage)
        assert np.allclose(d, dr, rtol=1e-14, atol=1e-15), (name, stage)


def test_generated_reducers_on_host():
    """Reducers outside the polynomial menu (reduction.py:80-190 accepts any expression):
    evaluated by a generated kernel into scratch arrays which the fold kernel then reads."""
    from pystella_b200.field import var
    rng = np.random.default_rng(7)
    N, h = (5, 4, 6), 2
    f, g, a = ps.Field("f", offset="h"), ps.Field("g"), var("a")
    decomp = ps.DomainDecomposition((1, 1, 1), h, N)
    reducers = {"c": [var("cos")(f) * g / a, (f**0.5, "max")], "p": [f * g]}
    with pytest.raises(NotImplementedError):
        ps.Reduction(decomp, reducers, halo_shape=h, jit=False)
    red = ps.Reduction(decomp, reducers, halo_shape=h)   # default: generated evaluation
    assert red._scratch_names == ["_psb_reducer_0", "_psb_reducer_1"]
    # the fold sees: scratch_0 (avg), scratch_1 (max), f*g (avg)
    assert [sorted(ref.name for mono in poly for ref, _ in mono) for poly in red._polys] == [
        ["_psb_reducer_0"], ["_psb_reducer_1"], ["f", "g"]]
    assert red.reduction_ops == ["avg", "max", "avg"]

    fh = rng.random(tuple(n + 2 * h for n in N)) + 0.1
    gh = rng.random(N)
    arrays = {"f": fh, "g": gh, "_psb_reducer_0": np.zeros(N), "_psb_reducer_1": np.zeros(N)}
    halos = {"f": h, "g": 0, "_psb_reducer_0": 0, "_psb_reducer_1": 0}
    _run_generated_on_host(red._generated_map.ksrc, arrays, halos, {"a": 1.7}, N)
    fi = fh[h:-h, h:-h, h:-h]
    assert np.allclose(arrays["_psb_reducer_0"], np.cos(fi) * gh / 1.7, rtol=1e-14)
    assert np.allclose(arrays["_psb_reducer_1"], fi**0.5, rtol=1e-14)


@pytest.mark.parametrize("h", [1, 2])
@pytest.mark.parametrize("cls", ["Stencil", "StreamingStencil"])
def test_stencil_map_with_raw_subscripts_on_host(h, cls):
    """test/test_stencil.py:36-90 verbatim in form: plain variables subscripted with
    ``i + h +- h``; padding and domain are inferred from the arrays."""
    from pystella_b200.field import var
    rng = np.random.default_rng(9)
    x, y = var("x"), var("y")
    i, j, k = var("i"), var("j"), var("k")
    map_dict = {y[i, j, k]: (x[i + h + h, j + h, k + h] + x[i + h, j + h + h, k + h]
                             + x[i + h, j + h, k + h + h] + x[i - h + h, j + h, k + h]
                             + x[i + h, j - h + h, k + h] + x[i + h, j + h, k - h + h])}
    with pytest.raises(NotImplementedError, match="jit"):
        getattr(ps, cls)(map_dict, h, prefetch_args=["x"], jit=False)
    with pytest.raises(TypeError):
        ps.Stencil(map_dict)
    # the reference's call, no flag (test/test_stencil.py:37-90)
    stencil_map = getattr(ps, cls)(map_dict, h, prefetch_args=["x"])
    assert stencil_map.prefetch_args == ["x"]
    # default: x staged in shared memory (executed in tests/test_cuda_emu.py); the direct
    # kernel (tile=False) is the one this per-site host transformation can run
    assert isinstance(stencil_map._generic, codegen.TiledStencilMap)
    assert stencil_map._generic.ksrc.prefetch[0]["hi"] == (2 * h, 2 * h, 2 * h)
    stencil_map = getattr(ps, cls)(map_dict, h, prefetch_args=["x"], tile=False)
    assert type(stencil_map._generic) is codegen.GenericMap
    ksrc = stencil_map._generic.ksrc
    N = (6, 5, 7)
    x_h = rng.random(tuple(n + 2 * h for n in N))
    y_h = rng.random(N)
    shape, halos = codegen.infer_layout(ksrc, [x_h.shape, y_h.shape], None, None)
    assert shape == N and halos == [h, 0]
    _run_generated_on_host(ksrc, {"x": x_h, "y": y_h}, {"x": h, "y": 0}, {}, N)
    y_true = (x_h[2*h:, h:-h, h:-h] + x_h[h:-h, 2*h:, h:-h] + x_h[h:-h, h:-h, 2*h:]
              + x_h[:-2*h, h:-h, h:-h] + x_h[h:-h, :-2*h, h:-h] + x_h[h:-h, h:-h, :-2*h])
    assert O_get_errs(y_true, y_h) < 5e-14
    with pytest.raises(ValueError, match="symmetric padding"):
        codegen.infer_layout(ksrc, [(N[0] + 3,) + x_h.shape[1:], y_h.shape], None, None)
    with pytest.raises(NotImplementedError):
        ps.StreamingStencil(map_dict, h, prefetch_args=["x", "y"], jit=True)


def O_get_errs(a, b):
    import oracle as O
    return O.get_errs(a, b)[0]


def _random_expression(rng, depth, leaves):
    from pystella_b200.field import var
    if depth == 0 or rng.random() < 0.2:
        if rng.random() < 0.3:
            return float(np.round(rng.uniform(-2, 2), 3))
        return leaves[rng.integers(len(leaves))]
    kind = rng.integers(7)
    a = _random_expression(rng, depth - 1, leaves)
    b = _random_expression(rng, depth - 1, leaves)
    if kind == 0:
        return a + b
    if kind == 1:
        return a - b
    if kind == 2:
        return a * b
    if kind == 3:
        return a / (1.5 + b * b)                   # keep denominators away from zero
    if kind == 4:
        return a ** int(rng.integers(0, 5))
    if kind == 5:
        return var(["sin", "cos", "tanh", "exp"][rng.integers(4)])(a * 0.3)
    return var("sqrt")(1.0 + a * a) + var("fabs")(b)


def test_random_expression_trees_translate_exactly():
    """60 random expression trees (sums, products, quotients, integer powers, elementary
    functions): the generated C evaluates to what `evaluate` gives, to rounding."""
    from pystella_b200.field import var
    rng = np.random.default_rng(2024)
    names = ["x", "y", "z"]
    leaves = [var(n) for n in names]
    exprs = [_random_expression(rng, 4, leaves) for _ in range(60)]
    exprs = [e for e in exprs if not isinstance(e, float)]
    vals = [0.37, -1.21, 0.83]
    got = _host_eval(exprs, names, vals)
    ctx = dict(zip(names, vals))
    for e, g in zip(exprs, got):
        want = float(evaluate(e, ctx))
        assert np.isclose(g, want, rtol=1e-12, atol=1e-13), (str(e), g, want)
