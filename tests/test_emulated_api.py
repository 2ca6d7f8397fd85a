"""
The public API of the diagnostics and stencil maps driven END TO END on host tensors: the
product's own launch code (marshalling, layout inference, grid sizes) with the CUDA libraries
replaced by the CPU execution model of the same kernel sources (tests/emu_device.py).  What
remains for the GPU box (tests/test_zz_gpu_diagnostics.py) is the hardware: NVRTC's code for
sm_100a, real atomics, launch limits.
"""
import numpy as np
import pytest
import torch

import pystella_b200 as ps
from emu_device import NUM_SMS, emulated_device
from pystella_b200.field import var

# a kernel that hangs on the execution model (a protocol bug) must fail the run, not stall it
pytestmark = pytest.mark.timeout(600, method="thread")


def test_histogrammer_through_the_emulated_device(monkeypatch):
    """test/test_histogram.py of the reference: two histograms sharing a bin expression."""
    rng = np.random.default_rng(15)
    grid, num_bins = (5, 9, 40), 23
    decomp = ps.DomainDecomposition((1, 1, 1), 0, grid_shape=grid)
    _fx = ps.Field("fx")
    hist = ps.Histogrammer(decomp, {"count": (var("abs")(_fx) * num_bins, 1),
                                    "squared": (var("abs")(_fx) * num_bins, _fx**2)},
                           num_bins, "float64", rank_shape=grid)
    fx = rng.random(grid)
    with emulated_device(monkeypatch) as dev:
        out = hist(None, fx=torch.from_numpy(fx))
    assert len(dev.jit.launches) == 1          # both histograms: one kernel, one pass
    name, g, b, smem = dev.jit.launches[0]
    assert b == (32, 8, 1) and g == (min(4 * NUM_SMS, 5 * 2), 1, 1)
    idx = np.floor(np.abs(fx) * num_bins).astype(int).reshape(-1)
    assert np.array_equal(out["count"], np.bincount(idx, minlength=num_bins))
    assert np.allclose(out["squared"],
                       np.bincount(idx, weights=fx.reshape(-1)**2, minlength=num_bins),
                       rtol=1e-13)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_field_histogrammer_through_the_emulated_device(monkeypatch, dtype):
    rng = np.random.default_rng(16)
    grid, h, num_bins = (4, 6, 37), 2, 11
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
    fh = ps.FieldHistogrammer(decomp, num_bins, "float64", halo_shape=h)
    pad = tuple(n + 2 * h for n in grid)
    f = (rng.random((2,) + pad) * 4.2 - 1.2).astype(dtype)
    with emulated_device(monkeypatch) as dev:
        res = fh(torch.from_numpy(f))
    assert len(dev.jit.launches) == 2          # one kernel per outer index
    for s in range(2):
        fi = f[s][h:-h, h:-h, h:-h]
        assert res["linear_bins"][s][0] == fi.min() and res["linear_bins"][s][-1] == fi.max()
        assert np.isclose(res["log_bins"][s][-1], np.abs(fi).max(), rtol=1e-14)
        assert np.isclose(res["log_bins"][s][0], np.abs(fi).min(), rtol=1e-13)
        assert res["linear"][s].sum() == np.prod(grid) == res["log"][s].sum()
        np_res = np.histogram(fi, bins=res["linear_bins"][s])[0]
        assert np.sum(np.abs(res["linear"][s] - np_res)) <= 4
        np_log = np.histogram(np.abs(fi), bins=res["log_bins"][s])[0]
        assert np.sum(np.abs(res["log"][s] - np_log)) <= 4


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_power_spectra_through_the_emulated_device(monkeypatch, dtype):
    """PowerSpectra.__call__ / bin_power: the kernel path == the per-mode tables of the host
    path (pystella/fourier/spectra.py:132-215), r2c layout, anisotropic box."""
    rng = np.random.default_rng(14)
    grid, h = (8, 6, 20), 1
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
    fft = ps.DFT(decomp, None, None, grid, dtype, backend="fftw")
    L = (5., 6., 7.)
    dk = tuple(2 * np.pi / Li for Li in L)
    spectra = ps.PowerSpectra(decomp, fft, dk, float(np.prod(L)))
    fx = rng.random(tuple(n + 2 * h for n in grid)).astype(dtype)
    ref = spectra(fx, k_power=2)                          # host tables
    fk = fft.dft(fx).clone()
    with emulated_device(monkeypatch):
        got = spectra(fx, k_power=2)
        got_frac = spectra.bin_power(fk, k_power=1.5)
    ok = np.isfinite(ref)
    rtol = 1e-12 if dtype == "float64" else 2e-5
    assert np.allclose(got[ok], ref[ok], rtol=rtol)
    ref_frac = spectra.bin_power(fk, k_power=1.5)
    assert np.allclose(got_frac[ok], ref_frac[ok], rtol=rtol)


@pytest.mark.parametrize("cls", ["Stencil", "StreamingStencil"])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_stencil_call_through_the_emulated_device(monkeypatch, cls, dtype):
    """test/test_stencil.py:36-90 with the reference's call; `x` staged in shared memory."""
    rng = np.random.default_rng(9)
    h, N = 2, (19, 11, 70)
    x, y = var("x"), var("y")
    i, j, k = var("i"), var("j"), var("k")
    map_dict = {y[i, j, k]: (x[i + h + h, j + h, k + h] + x[i + h, j + h + h, k + h]
                             + x[i + h, j + h, k + h + h] + x[i - h + h, j + h, k + h]
                             + x[i + h, j - h + h, k + h] + x[i + h, j + h, k - h + h])}
    stencil_map = getattr(ps, cls)(map_dict, h, prefetch_args=["x"], dtype=dtype)
    x_h = rng.random(tuple(n + 2 * h for n in N)).astype(dtype)
    y_t = torch.zeros(N, dtype=getattr(torch, dtype))
    with emulated_device(monkeypatch) as dev:
        stencil_map(None, x=torch.from_numpy(x_h), y=y_t)
        name, g, b, smem = dev.jit.launches[-1]
        assert b == (64, 4, 1) and g == (2, 2, 2) and smem == stencil_map._generic.ksrc.smem_bytes
        with pytest.raises(ValueError, match="outside its padding"):
            stencil_map(None, x=torch.from_numpy(np.ascontiguousarray(x_h[1:-1, 1:-1, 1:-1])),
                        y=torch.zeros(tuple(n + 2 for n in N), dtype=getattr(torch, dtype)))
    y_true = (x_h[2*h:, h:-h, h:-h] + x_h[h:-h, 2*h:, h:-h] + x_h[h:-h, h:-h, 2*h:]
              + x_h[:-2*h, h:-h, h:-h] + x_h[h:-h, :-2*h, h:-h] + x_h[h:-h, h:-h, :-2*h])
    tol = 5e-14 if dtype == "float64" else 1e-5
    assert np.max(np.abs(y_t.numpy() - y_true)) <= tol * np.max(np.abs(y_true))


def test_elementwise_map_call_through_the_emulated_device(monkeypatch):
    """The generic ElementWiseMap launch path (GenericMap.__call__) on the same harness."""
    rng = np.random.default_rng(2)
    N, h = (3, 5, 70), 1
    f = ps.Field("f", offset="h")
    out = ps.Field("out", offset=0)
    m = ps.ElementWiseMap({out: var("c") * f**2 + 1}, halo_shape=h)
    f_h = rng.random(tuple(n + 2 * h for n in N))
    o = torch.zeros(N, dtype=torch.float64)
    with emulated_device(monkeypatch):
        m(None, f=torch.from_numpy(f_h), out=o, c=2.5)
    assert np.allclose(o.numpy(), 2.5 * f_h[h:-h, h:-h, h:-h]**2 + 1, rtol=1e-15)


def test_bench_diagnostics_record_on_the_emulated_device(monkeypatch, capsys):
    """bench.py --diagnostics (the extras record that times the new kernels on the GPU box)
    executed on a 16^3 lattice of host tensors: every section reports ok."""
    import importlib
    import os
    import sys
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    bench = importlib.import_module("bench")

    def timed(fn, reps=1, warm=0):
        fn()
        return 1.0

    with emulated_device(monkeypatch):
        rec = bench.run_diagnostics(types.SimpleNamespace(grid=16), dev=torch.device("cpu"),
                                    timed=timed, dft_backend="fftw")
    capsys.readouterr()
    for name in ("field_histogram", "bin_power", "stencil_map"):
        assert "error" not in rec[name], rec[name]
        assert rec[name]["ok"] is True, rec[name]
    assert rec["field_histogram"]["sites_counted"] == [16.**3, 16.**3]
    assert rec["bin_power"]["kshape"] == [16, 16, 9]


def test_stencil_autotune_keeps_the_faster_kernel(monkeypatch):
    """Stencil's first call times the shared-memory and the direct kernel (idempotent maps
    only) and keeps the faster one; in-place maps are never run twice."""
    import time
    import types

    class FakeEvent:
        def __init__(self, enable_timing=False):
            self.t = None

        def record(self, stream=None):
            self.t = time.perf_counter()

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    rng = np.random.default_rng(4)
    h, N = 1, (4, 8, 64)
    x, y = var("x"), var("y")
    i, j, k = var("i"), var("j"), var("k")
    md = {y[i, j, k]: x[i + h + h, j + h, k + h] - x[i, j + h, k + h]}
    st = ps.Stencil(md, h, prefetch_args=["x"])
    assert set(st._candidates) == {"tiled", "direct"}
    x_h = rng.random(tuple(n + 2 * h for n in N))
    y_t = torch.zeros(N, dtype=torch.float64)
    with emulated_device(monkeypatch) as dev:
        monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
        monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
        monkeypatch.setattr(torch.cuda, "current_stream", lambda: types.SimpleNamespace())
        monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
        st(None, x=torch.from_numpy(x_h), y=y_t)
        assert st._candidates is None and st.tuning["chosen"] in ("tiled", "direct")
        assert len(dev.jit.launches) == 2 * 4 + 1      # (1 + 3) launches each, then the call
        assert st._generic is {"tiled": st._generic, "direct": st._generic}[st.tuning["chosen"]]
    assert np.allclose(y_t.numpy(), x_h[2*h:, h:-h, h:-h] - x_h[:-2*h, h:-h, h:-h], rtol=1e-15)

    # an in-place map (y read and written) must not be autotuned
    f = ps.Field("f", offset="h")
    acc = ps.Field("acc", offset=0)
    from pystella_b200.field import shift_fields
    st2 = ps.Stencil({acc: acc + shift_fields(f, (1, 0, 0))}, h, prefetch_args=["f"])
    assert st2._candidates is None
    from pystella_b200.codegen import TiledStencilMap
    assert isinstance(st2._generic, TiledStencilMap) and not st2._generic.ksrc.idempotent


def test_example_rho_diagnostics_through_the_emulated_device(monkeypatch):
    """examples/scalar_preheating.py (reference lines 139-163): rho = T_00 / (3 mpl^2 H^2 /
    8 pi) through a generated ElementWiseMap, then its FieldHistogrammer -- the example's
    own construction, executed on host tensors."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "examples", "scalar_preheating.py")
    spec = importlib.util.spec_from_file_location("example_scalar_preheating", path)
    example = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(example)

    rng = np.random.default_rng(8)
    grid, h, ns = (4, 8, 40), 2, 2
    p = example.parser.parse_args([])
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
    sector = ps.ScalarSector(ns, potential=example.make_potential(p))
    hist, compute_rho = example.make_rho_diagnostics(decomp, sector, p.mpl, h, grid, "float64")
    pad = tuple(n + 2 * h for n in grid)
    f = 0.2 + 0.05 * rng.standard_normal((ns,) + pad)
    dfdt = -0.1 + 0.05 * rng.standard_normal((ns,) + pad)
    dfdx = 0.05 * rng.standard_normal((ns, 3) + grid)
    a, hubble = 1.3, 0.41
    rho = torch.zeros(grid, dtype=torch.float64)
    with emulated_device(monkeypatch):
        compute_rho(None, a=np.array(a), hubble=np.array(hubble), rho=rho,
                    f=torch.from_numpy(f), dfdt=torch.from_numpy(dfdt),
                    dfdx=torch.from_numpy(dfdx))
        res = hist(rho)
    fi = f[:, h:-h, h:-h, h:-h]
    di = dfdt[:, h:-h, h:-h, h:-h]
    # T_00 = sum_n dfdt_n^2 + g_00 L,  L = -kinetic/2 - V,  g_00 = -a^2
    kinetic = sum(-di[n]**2 / a**2 + (dfdx[n]**2).sum(axis=0) / a**2 for n in range(ns))
    V = example.make_potential(p)([fi[0], fi[1]])
    T00 = (di**2).sum(axis=0) - a**2 * (-kinetic / 2 - V)
    ref = T00 / (3 * p.mpl**2 * hubble**2 / 8 / np.pi)
    assert np.allclose(rho.numpy(), ref, rtol=1e-13)
    assert res["linear"].sum() == np.prod(grid) == res["log"].sum()
    assert res["linear"].shape == (1000,) and res["linear_bins"].shape == (1001,)
    assert np.sum(np.abs(res["linear"] - np.histogram(ref, bins=res["linear_bins"])[0])) <= 4
