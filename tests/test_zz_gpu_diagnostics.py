"""
GPU tests of the diagnostics next to the time loop (row f-4 of the scope table): the
spectrum binning kernel (`psb_bin_power`, csrc/diag.cu), the generated Histogrammer kernels
(codegen.generate_histogram_kernel) and the cuFFT-backed transforms.  The kernels' logic is
also executed on the CPU model of tests/cuda_emu (tests/test_cuda_emu.py).

This file sorts after the hot-path parity files on purpose: the stage / halo / reduction
gates run first.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # script mode
import pystella_b200 as ps  # noqa: E402
from helpers import to_dev  # noqa: E402

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu]


def test_fourier_on_device():
    """cuFFT path of DFT / RayleighGenerator / PowerSpectra == the host backend."""
    rng = np.random.default_rng(14)
    grid, h = (16, 12, 20), 2
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
    dev = ps.DFT(decomp, None, None, grid, "float64")
    host = ps.DFT(decomp, None, None, grid, "float64", backend="fftw")
    assert dev.fk.is_cuda and not host.fk.is_cuda
    pad = tuple(n + 2 * h for n in grid)
    fx = rng.random(pad)
    fk_d = dev.dft(to_dev(fx)).cpu().numpy()
    fk_h = host.dft(fx).numpy()
    assert np.allclose(fk_d, fk_h, rtol=1e-12, atol=1e-10)
    back = torch.zeros(pad, dtype=torch.float64, device="cuda")
    dev.idft(to_dev(fk_h), back)
    assert np.allclose(back.cpu().numpy()[h:-h, h:-h, h:-h],
                       np.prod(grid) * fx[h:-h, h:-h, h:-h], rtol=1e-11)
    assert float(back[0].abs().max()) == 0.

    L = (5., 6., 7.)
    dk = tuple(2 * np.pi / Li for Li in L)
    rands = rng.random((4,) + host.shape(True))
    kw = dict(norm=1.44e-12, omega_k=lambda k: np.sqrt(k**2 + 0.7), hubble=0.31)
    out = []
    for fft in (dev, host):
        modes = ps.RayleighGenerator(None, fft, dk, float(np.prod(L)), seed=1)
        modes._uniform = lambda shape, fft=fft: torch.from_numpy(rands.copy()).to(fft.device)
        out.append([x.cpu().numpy() for x in modes.generate_WKB(None, **kw)])
    assert np.allclose(out[0][0], out[1][0], rtol=1e-12, atol=1e-24)
    assert np.allclose(out[0][1], out[1][1], rtol=1e-12, atol=1e-24)
    s_d = ps.PowerSpectra(decomp, dev, dk, float(np.prod(L)))(to_dev(fx))
    s_h = ps.PowerSpectra(decomp, host, dk, float(np.prod(L)))(fx)
    ok = np.isfinite(s_h)
    assert np.allclose(s_d[ok], s_h[ok], rtol=1e-11)


def test_histograms_on_device():
    """Histogrammer / FieldHistogrammer on CUDA tensors == numpy (test/test_histogram.py)."""
    from pystella_b200.field import var
    rng = np.random.default_rng(15)
    grid, h, num_bins = (24, 16, 20), 1, 123
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
    _fx = ps.Field("fx")
    hist = ps.Histogrammer(decomp, {"count": (var("abs")(_fx) * num_bins, 1),
                                    "squared": (var("abs")(_fx) * num_bins, _fx**2)},
                           num_bins, "float64", rank_shape=grid)
    fx = rng.random(grid)
    out = hist(None, fx=to_dev(fx))
    idx = np.floor(np.abs(fx) * num_bins).astype(int).reshape(-1)
    assert np.array_equal(out["count"], np.bincount(idx, minlength=num_bins))
    assert np.allclose(out["squared"],
                       np.bincount(idx, weights=fx.reshape(-1)**2, minlength=num_bins),
                       rtol=1e-12)
    fh = ps.FieldHistogrammer(decomp, num_bins, "float64", rank_shape=grid, halo_shape=h)
    pad = tuple(n + 2 * h for n in grid)
    f = rng.random((2,) + pad) * 4.2 - 1.2
    res = fh(to_dev(f))
    for s in range(2):
        fi = f[s][h:-h, h:-h, h:-h]
        assert res["linear"][s].sum() == np.prod(grid)
        np_res = np.histogram(fi, bins=res["linear_bins"][s])[0]
        assert np.sum(np.abs(res["linear"][s] - np_res)) <= 4
        assert res["log"][s].sum() == np.prod(grid)


def test_distributed_fourier_nccl():
    """Slab DFT over NCCL all-to-all == numpy on the gathered array (2 GPUs)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from test_distributed import _run
    _run("gpu_fourier", 2)


def _numpy_spectrum(fk, kx, ky, kz, dk, bw, nb, k_power, is_real, nz):
    kv = np.meshgrid(kx, ky, kz, indexing="ij")
    kmag = np.sqrt(sum((d * k.astype(np.float64))**2 for d, k in zip(dk, kv)))
    b = np.floor(kmag / bw + .5).astype(int)
    mult = (np.where((kv[2] > 0) & (kv[2] < nz / 2), 2., 1.) if is_real
            else np.ones_like(kmag))
    w = mult * kmag**k_power * (fk.real.astype(np.float64)**2 + fk.imag.astype(np.float64)**2)
    ok = b.ravel() < nb
    return np.bincount(b.ravel()[ok], weights=w.ravel()[ok], minlength=nb)[:nb]


@pytest.mark.parametrize("grid, is_real, dtype, k_power, bw_scale", [
    ((32, 24, 40), True, "f64", 3, 1.0),       # r2c layout, per-warp shared copies
    ((32, 24, 40), False, "f64", 3, 1.0),      # c2c: |k| not monotone along z
    ((16, 20, 70), True, "f32", 2.5, 1.0),     # complex64, pow(), Kz = 36
    ((16, 20, 140), False, "f64", 0, 1.0),     # Kz > 64 with a tail
    ((24, 24, 24), True, "f64", 1, 0.002),     # ~10^4 bins > 48 KB: global atomics only
])
def test_bin_power_kernel(grid, is_real, dtype, k_power, bw_scale):
    """psb_bin_power through the C ABI == numpy (pystella/fourier/spectra.py:103-131)."""
    from pystella_b200 import _kernels
    rng = np.random.default_rng(21)
    rd = np.float64 if dtype == "f64" else np.float32
    cd = np.complex128 if dtype == "f64" else np.complex64
    kx = np.fft.fftfreq(grid[0], 1 / grid[0]).astype(rd)
    ky = np.fft.fftfreq(grid[1], 1 / grid[1]).astype(rd)
    kz = (np.fft.rfftfreq(grid[2], 1 / grid[2]) if is_real
          else np.fft.fftfreq(grid[2], 1 / grid[2])).astype(rd)
    K = (len(kx), len(ky), len(kz))
    fk = (rng.standard_normal(K) + 1j * rng.standard_normal(K)).astype(cd)
    dk = (0.3, 0.41, 0.37)
    bw = min(dk) * bw_scale
    kmax = np.sqrt(sum((d * np.abs(k).max())**2 for d, k in zip(dk, (kx, ky, kz))))
    nb = int(kmax / bw + .5) + 1
    ref = _numpy_spectrum(fk, kx, ky, kz, dk, bw, nb, k_power, is_real, grid[2])
    for blocks in (0, 3):
        got = _kernels.bin_power(to_dev(fk), [to_dev(k) for k in (kx, ky, kz)], dk, bw,
                                 k_power, grid[2] / 2, nb, is_real, blocks=blocks)
        got = got.cpu().numpy()
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_power_spectrum_of_a_plane_wave():
    """Size-independent property at a configuration-sized lattice (256^3): a single plane
    wave of wavenumber (3, 4, 12) puts all its power into bin round(|k| / dk) = 13, and
    Parseval: sum_k count |f_k|^2 = N^3 sum_x f^2 (spectra.py:165-215)."""
    n, h = 256, 2
    grid = (n, n, n)
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
    fft = ps.DFT(decomp, None, None, grid, "float64")
    L = 2 * np.pi
    dk = (1., 1., 1.)
    x = torch.arange(n, dtype=torch.float64, device="cuda") * (L / n)
    f = torch.zeros(tuple(m + 2 * h for m in grid), dtype=torch.float64, device="cuda")
    f[h:-h, h:-h, h:-h] = torch.cos(3 * x[:, None, None] + 4 * x[None, :, None]
                                    + 12 * x[None, None, :])
    spectra = ps.PowerSpectra(decomp, fft, dk, L**3)
    fk = fft.dft(f)
    raw = np.nan_to_num(spectra.bin_power(fk, k_power=0)) * spectra.bin_counts
    assert np.argmax(raw) == 13
    assert raw[13] / raw.sum() > 1 - 1e-12
    total = float((f[h:-h, h:-h, h:-h]**2).sum()) * n**3
    assert abs(raw.sum() / total - 1) < 1e-12


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_generated_histograms_match_numpy(dtype):
    """Two histograms in one generated kernel, tails in y and z, out-of-range bins dropped."""
    from pystella_b200.field import var
    rng = np.random.default_rng(33)
    grid, h, nb = (10, 13, 75), 2, 64
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
    f = ps.Field("f", offset="h")
    hist = ps.Histogrammer(decomp, {"a": (f * var("scale"), 1),
                                    "b": (var("abs")(f) * var("scale"), f * f)},
                           nb, "float64", halo_shape=h)
    pad = tuple(n + 2 * h for n in grid)
    arr = (rng.standard_normal(pad) * 2).astype(dtype)
    scale = 17.3
    out = hist(None, f=to_dev(arr), scale=scale)
    x = arr[h:-h, h:-h, h:-h]
    T = np.dtype(dtype).type
    for name, b in (("a", np.floor(x * T(scale))), ("b", np.floor(np.abs(x) * T(scale)))):
        ok = (b >= 0) & (b < nb)
        w = np.ones_like(x, dtype=np.float64) if name == "a" else (x * x).astype(np.float64)
        ref = np.bincount(b[ok].astype(int), weights=w[ok], minlength=nb)
        assert np.allclose(out[name], ref, rtol=1e-12 if dtype == "float64" else 1e-6,
                           atol=0)


def test_field_histogram_bounds_and_counts_at_config_size():
    """FieldHistogrammer on a 256^3 fp64 field: bounds from psb_reduce == torch, every site
    counted exactly once in both histograms (histogram.py:210-350)."""
    n, h, nb = 256, 2, 432
    grid = (n, n, n)
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
    gen = torch.Generator(device="cuda").manual_seed(7)
    f = torch.randn((1,) + tuple(m + 2 * h for m in grid), dtype=torch.float64,
                    device="cuda", generator=gen)
    fh = ps.FieldHistogrammer(decomp, nb, "float64", halo_shape=h)
    res = fh(f)
    x = f[0, h:-h, h:-h, h:-h]
    assert res["linear_bins"][0][0] == float(x.min())
    assert res["linear_bins"][0][-1] == float(x.max())
    assert np.isclose(res["log_bins"][0][-1], float(x.abs().max()), rtol=1e-14)
    assert res["linear"][0].sum() == n**3 and res["log"][0].sum() == n**3
    ref = torch.histc(x, bins=nb, min=float(x.min()), max=float(x.max())).cpu().numpy()
    assert np.sum(np.abs(res["linear"][0] - ref)) <= 8


# ---------------------------------------------------------------------------
# Stencil / StreamingStencil with prefetch_args staged in shared memory (row a5)
# ---------------------------------------------------------------------------

def _six_point(h):
    from pystella_b200.field import var
    x, y = var("x"), var("y")
    i, j, k = var("i"), var("j"), var("k")
    return {y[i, j, k]: (x[i + h + h, j + h, k + h] + x[i + h, j + h + h, k + h]
                         + x[i + h, j + h, k + h + h] + x[i - h + h, j + h, k + h]
                         + x[i + h, j - h + h, k + h] + x[i + h, j + h, k - h + h])}


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("cls", ["Stencil", "StreamingStencil"])
@pytest.mark.parametrize("h, N", [(1, (16, 12, 40)), (2, (37, 19, 150))])
def test_tiled_stencil_map(cls, dtype, h, N):
    """test/test_stencil.py:36-90 with the reference's call and tolerances: `x` staged in
    shared memory (tails in y and z, several x chunks)."""
    from pystella_b200.codegen import TiledStencilMap
    rng = np.random.default_rng(16)
    stencil_map = getattr(ps, cls)(_six_point(h), h, prefetch_args=["x"], dtype=dtype,
                                   autotune=False)
    assert isinstance(stencil_map._generic, TiledStencilMap)
    x_h = rng.random(tuple(n + 2 * h for n in N)).astype(dtype)
    yd = torch.zeros(N, dtype=getattr(torch, dtype), device="cuda")
    stencil_map(None, x=to_dev(x_h), y=yd)
    y_true = (x_h[2*h:, h:-h, h:-h] + x_h[h:-h, 2*h:, h:-h] + x_h[h:-h, h:-h, 2*h:]
              + x_h[:-2*h, h:-h, h:-h] + x_h[h:-h, :-2*h, h:-h] + x_h[h:-h, h:-h, :-2*h])
    tol = 5e-14 if dtype == "float64" else 1e-5
    assert np.max(np.abs(yd.cpu().numpy() - y_true)) <= tol * np.max(np.abs(y_true))


def test_tiled_stencil_equals_direct_kernel_at_config_size():
    """256^3 fp64, h = 2: the shared-memory-tile kernel and the direct kernel agree bit for
    bit (same expression, same operand order), and both equal shifted tensor sums."""
    h, n = 2, 256
    N = (n, n, n)
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(tuple(m + 2 * h for m in N), dtype=torch.float64, device="cuda",
                   generator=gen)
    y_tile = torch.zeros(N, dtype=torch.float64, device="cuda")
    y_direct = torch.zeros_like(y_tile)
    ps.Stencil(_six_point(h), h, prefetch_args=["x"], autotune=False)(None, x=x, y=y_tile)
    ps.Stencil(_six_point(h), h, prefetch_args=["x"], tile=False)(None, x=x, y=y_direct)
    assert torch.equal(y_tile, y_direct)
    # the default: the first call times both kernels on these arrays and keeps the faster
    tuned = ps.Stencil(_six_point(h), h, prefetch_args=["x"])
    y_tuned = torch.zeros_like(y_tile)
    tuned(None, x=x, y=y_tuned)
    assert tuned.tuning["chosen"] in ("tiled", "direct") and tuned._candidates is None
    assert tuned.tuning["tiled"] > 0 and tuned.tuning["direct"] > 0
    tuned(None, x=x, y=y_tuned)
    assert torch.equal(y_tuned, y_direct)
    c = slice(h, -h)
    ref = ((((x[2*h:, c, c] + x[c, 2*h:, c]) + x[c, c, 2*h:]) + x[:-2*h, c, c])
           + x[c, :-2*h, c]) + x[c, c, :-2*h]
    assert torch.equal(y_tile, ref)


# ---------------------------------------------------------------------------
# fused stepper: non-quadratic polynomial potentials on layouts without a TMA description
# ---------------------------------------------------------------------------

@pytest.mark.parametrize("h, N", [(1, (12, 10, 16)), (2, (12, 10, 17)), (3, (8, 12, 14))])
def test_fused_two_fields_on_non_tma_layouts(h, N):
    """ScalarSector(2) with the preheating potential under FusedStencilStage prefers the
    run-time generated instance of the TMA pipeline; for odd radii / odd rows that pipeline
    does not exist and the table-driven stage kernel must serve the system (found by running
    the stage on the CPU execution model: the stepper used to raise "served by the
    cooperative loader").  Five stages against the oracle, in-kernel halos included."""
    import oracle as O
    from test_gpu_api import _example_potential
    from helpers import scale_err
    rng = np.random.default_rng(49279)
    pad = tuple(n + 2 * h for n in N)
    dx = tuple(5 / n for n in N)
    dt = 0.1 * min(dx)
    potential, terms = _example_potential()
    f0 = (np.array([.193, 0.])[:, None, None, None] + 1e-3 * rng.standard_normal((2,) + pad))
    d0 = (np.array([-.142231, 0.])[:, None, None, None]
          + 1e-3 * rng.standard_normal((2,) + pad))
    a, hubble = 1.07, 0.45
    fr, dr = f0.copy(), d0.copy()
    kfr, kdr = np.zeros_like(fr), np.zeros_like(fr)
    lap = np.zeros((2,) + N)
    for s in range(5):
        O.reference_schedule_stage(s, fr, dr, kfr, kdr, lap, h, dx, dt, a=a, hubble=hubble,
                                   potential_terms=terms)
    sector = ps.ScalarSector(2, potential=potential)
    decomp = ps.DomainDecomposition((1, 1, 1), h, N)
    stepper = ps.LowStorageRK54([sector], halo_shape=h, rank_shape=N, dt=dt,
                                MapKernel=ps.FusedStencilStage, dx=dx, decomp=decomp)
    assert getattr(stepper.systems[0], "tables", None) is not None   # generated preferred
    f, dfdt = to_dev(f0), to_dev(d0)
    for s in range(5):
        stepper(s, f=f, dfdt=dfdt, a=np.array(a), hubble=np.array([hubble]))
    for got, ref in ((f, fr), (dfdt, dr)):
        for i in range(2):
            assert scale_err(O.interior(got[i].cpu().numpy(), h),
                             O.interior(ref[i], h)) < 1e-12


@pytest.mark.parametrize("N, h, dtype", [
    ((10, 12, 18), 2, np.float32),     # Nz + 2h = 22: rows of 88 B, no tensor map possible
    ((10, 12, 16), 2, np.float32),     # Nz + 2h = 20: 80 B rows, tensor map possible
    ((8, 9, 14), 4, np.float64),       # odd Ny, radius 4
])
def test_config4_stage_on_rows_without_tma_description(N, h, dtype):
    """Two scalars + tensor sector under FusedStencilStage on lattices whose rows are not
    multiples of 16 bytes in fp32: the fused tensor stage cannot be described to TMA and the
    element-wise tensor stage must serve them (found by the randomised sweep on the CPU
    execution model: the stepper used to raise "layout not describable by TMA")."""
    from helpers import gw_stage_case
    gw_stage_case(N, h, dtype, 3, np.random.default_rng(77), to_dev)
