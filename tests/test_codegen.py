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
