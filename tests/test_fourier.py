"""
Fourier-space companions on the HOST backend
(``backend="fftw"``, the analogue of the reference's pDFT): conventions against NumPy
(test/test_dft.py), the reference's own make_hermitian / zero_corner_modes outputs
(tests/golden), mode generation against the oracle on identical deviates, and the
known-answer tests of test/test_rayleigh.py and test/test_spectra.py.
"""
import numpy as np
import pytest
import torch

import oracle as O
import pystella_b200 as ps
from oracle import fourier_oracle as F
from pystella_b200 import fourier


def make_fft(grid_shape, dtype="float64", h=1):
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid_shape)
    return decomp, ps.DFT(decomp, None, None, grid_shape, dtype, backend="fftw")


@pytest.mark.parametrize("dtype", ["float64", "complex128", "float32"])
@pytest.mark.parametrize("grid_shape", [(16, 12, 8), (8, 8, 8)])
def test_dft_conventions(dtype, grid_shape):
    """test/test_dft.py:36-104: DFT == numpy's (r)fftn, IDFT(DFT(f)) == grid_size * f."""
    rng = np.random.default_rng(1)
    decomp, fft = make_fft(grid_shape, dtype)
    fx = rng.random(grid_shape) + 1e-2
    if not fft.is_real:
        fx = fx + 1j * rng.random(grid_shape)
    fx = fx.astype(dtype)
    rtol = 1e-11 if fft.rdtype == np.float64 else 2e-3
    np_dft = np.fft.rfftn if fft.is_real else np.fft.fftn
    fk = fft.dft(fx).numpy().copy()
    assert fk.shape == fft.shape(True) and fk.dtype == fft.cdtype
    assert O.get_errs(np_dft(fx.astype(fft.cdtype if not fft.is_real else "float64")), fk)[0] \
        < rtol
    fx2 = fft.idft(fk).numpy().copy()      # attached buffer: copy (dft.py:160-166)
    assert O.get_errs(fx, fx2 / np.prod(grid_shape))[0] < rtol

    # every array flavour of test/test_dft.py:106-140: unpadded / padded, tensor / NumPy
    h = 1
    pad = tuple(n + 2 * h for n in grid_shape)
    fx_pad = np.zeros(pad, dtype)
    fx_pad[h:-h, h:-h, h:-h] = fx
    for arr in (fx_pad, torch.from_numpy(fx_pad), torch.from_numpy(fx)):
        out = np.empty(fft.shape(True), fft.cdtype)
        assert fft.dft(arr, out) is out
        assert np.array_equal(out, fk)
    back = np.full(pad, 7, dtype)
    fft.idft(fk, back)
    assert np.array_equal(back[h:-h, h:-h, h:-h], fx2)
    assert np.all(back[0] == 7)                      # halos untouched (restore_halos)
    with pytest.raises(ValueError):
        fft.dft(np.zeros((3, 3, 3), dtype))


def test_backend_selection():
    decomp = ps.DomainDecomposition((1, 1, 1), 0, grid_shape=(4, 4, 4))
    with pytest.raises(NotImplementedError):
        ps.DFT(decomp, None, None, (4, 4, 4), "float64", backend="mkl")
    if not torch.cuda.is_available():
        with pytest.raises(ps.PsbError, match="no CUDA device"):
            ps.DFT(decomp, None, None, (4, 4, 4), "float64")


def _cplx(pair):
    return np.array(pair[0]) + 1j * np.array(pair[1])


def test_hermitian_and_corner_modes_golden(golden):
    """== the reference's make_hermitian / zero_corner_modes run by make_golden.py."""
    g = golden["fourier"]
    grid = tuple(g["grid_shape"])
    _, fft = make_fft(grid)
    fk = _cplx(g["fk"])
    for flavour in (lambda a: a, torch.from_numpy):
        herm = fourier.make_hermitian(flavour(fk.copy()))
        assert np.array_equal(np.asarray(herm), _cplx(g["make_hermitian"]))
        zi = fft.zero_corner_modes(herm, only_imag=True)
        assert np.array_equal(np.asarray(zi), _cplx(g["zero_corner_modes_only_imag"]))
        z = fft.zero_corner_modes(flavour(fk.copy()))
        assert np.array_equal(np.asarray(z), _cplx(g["zero_corner_modes"]))
    for n, ref in golden["fftfreq"].items():
        assert fourier.fftfreq(int(n)).tolist() == ref
    wk = np.array(g["ps_wrapper"]["wk"])
    km = wk.copy()
    km[0, 0, 0] = 0.
    power = ps.RayleighGenerator._ps_wrapper(lambda w: 1 / 2 / w, wk, km)
    assert np.array_equal(power, np.array(g["ps_wrapper"]["power"]))
    assert np.array_equal(wk, np.array(g["ps_wrapper"]["wk_after"]))


@pytest.mark.parametrize("grid_shape", [(16, 16, 16), (12, 8, 10)])
def test_generate_wkb_matches_oracle(grid_shape):
    """Identical uniform deviates -> identical modes and fields as the oracle's restatement
    of rayleigh.py:57-82, 307-383."""
    rng = np.random.default_rng(3)
    L = (5., 6., 7.)
    dk = tuple(2 * np.pi / Li for Li in L)
    volume = float(np.prod(L))
    decomp, fft = make_fft(grid_shape, h=2)
    modes = ps.RayleighGenerator(None, fft, dk, volume, seed=1)
    kshape = fft.shape(True)
    rands = rng.random((4,) + kshape)
    modes._uniform = lambda shape: torch.from_numpy(rands.copy())
    kw = dict(norm=1.44e-12, omega_k=lambda k: np.sqrt(k**2 + 0.7), hubble=0.31)
    fk, dfk = modes.generate_WKB(None, **kw)
    fk_o, dfk_o = F.generate_wkb(rands, grid_shape, dk, volume, **kw)
    assert O.get_errs(fk_o, fk.numpy())[0] < 1e-12
    assert np.allclose(dfk.numpy(), dfk_o, rtol=1e-12, atol=1e-22)

    pad = tuple(n + 4 for n in grid_shape)
    f, d = np.zeros(pad), np.zeros(pad)
    modes.init_WKB_fields(f, d, **kw)
    fx_o, dfx_o = F.init_wkb_fields(rands, grid_shape, dk, volume, **kw)
    assert np.allclose(f[2:-2, 2:-2, 2:-2], fx_o, rtol=1e-11, atol=1e-20)
    assert np.allclose(d[2:-2, 2:-2, 2:-2], dfx_o, rtol=1e-11, atol=1e-20)
    # Hermitian modes: the backward transform of the full-complex spectrum is real too
    full = np.fft.ifftn(_full_spectrum(fk.numpy(), grid_shape))
    assert np.max(np.abs(full.imag)) < 1e-12 * np.max(np.abs(full.real))

    # non-WKB generation on identical deviates
    r2 = rng.random((2,) + kshape)
    modes._uniform = lambda shape: torch.from_numpy(r2.copy())
    g = modes.generate(None, norm=2.0, field_ps=lambda k: k**-2)
    g_o = F.generate(r2, grid_shape, dk, volume, norm=2.0, field_ps=lambda k: k**-2)
    assert O.get_errs(g_o, g.numpy())[0] < 1e-12


def _full_spectrum(fk, grid_shape):
    """Half-complex -> full complex array using f(-k) = conj f(k)."""
    nx, ny, nz = grid_shape
    full = np.zeros(grid_shape, complex)
    full[:, :, :nz // 2 + 1] = fk
    for k in range(nz // 2 + 1, nz):
        src = fk[:, :, nz - k]
        full[:, :, k] = np.conj(np.roll(np.roll(src[::-1, ::-1], 1, axis=0), 1, axis=1))
    return full


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
@pytest.mark.parametrize("random", [True, False])
def test_generate_spectrum(dtype, random):
    """test/test_rayleigh.py:63-117: generated modes carry the requested power law, and
    the field is not skewed."""
    grid_shape = (32, 32, 32)
    decomp, fft = make_fft(grid_shape, dtype)
    num_bins = int(sum(n**2 for n in grid_shape)**.5 / 2 + .5) + 1
    L = (10,) * 3
    volume = np.prod(L)
    dk = tuple(2 * np.pi / Li for Li in L)
    spectra = ps.PowerSpectra(decomp, fft, dk, volume)
    modes = ps.RayleighGenerator(None, fft, dk, volume, seed=5123)
    kbins = min(dk) * np.arange(0, num_bins)
    test_norm = 1 / 2 / np.pi**2 / np.prod(grid_shape)**2
    for exp in (-1, -2, -3):
        fk = modes.generate(None, random=random, norm=1, field_ps=lambda k, e=exp: k**e)
        spectrum = spectra.norm * spectra.bin_power(fk, k_power=3)[1:-1]
        true_spectrum = test_norm * kbins[1:-1]**3 * kbins[1:-1]**exp
        err = np.abs(1 - spectrum / true_spectrum)
        tol = .1 if num_bins < 64 else .3
        assert np.max(err[num_bins // 3:-num_bins // 3]) < tol and np.average(err[1:]) < tol
        if random:
            fx = fft.idft(fk).numpy().real
            avg = np.mean(fx)
            var = np.mean(fx**2) - avg**2
            skew = (np.mean(fx**3) - 3 * avg * var - avg**3) / var**1.5
            assert skew < tol


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
@pytest.mark.parametrize("L", [(10,) * 3, (10, 7, 8), (3, 8, 19), (13.2, 5.71, 9.4),
                               (11, 11, 4), (4, 11, 11), (11, 4, 11)])
def test_spectra(dtype, L):
    """test/test_spectra.py:48-121 against numpy.histogram."""
    rng = np.random.default_rng(4)
    grid_shape = (16, 16, 16)
    decomp, fft = make_fft(grid_shape, dtype)
    dk = tuple(2 * np.pi / Li for Li in L)
    spec = ps.PowerSpectra(decomp, fft, dk, np.prod(L), bin_width=min(dk) + .001)
    assert int(np.sum(spec.bin_counts)) == np.prod(grid_shape)
    k_power = 2.
    kshape = fft.shape(True)
    fk = (rng.random(kshape) + 1j * rng.random(kshape)).astype(fft.cdtype)
    spectrum = spec.bin_power(torch.from_numpy(fk), k_power=k_power)
    bins = np.arange(-.5, spec.num_bins + .5) * spec.bin_width
    kvecs = np.meshgrid(*[k.numpy() for k in fft.sub_k.values()], indexing="ij")
    kmags = np.sqrt(sum((dki * ki)**2 for dki, ki in zip(dk, kvecs)))
    if fft.is_real:
        counts = 2. * np.ones_like(kmags)
        counts[kvecs[2] == 0] = 1
        counts[kvecs[2] == grid_shape[-1] // 2] = 1
    else:
        counts = np.ones_like(kmags)
    ones = spec.bin_power(np.ones_like(fk), k_power=0)
    ok = spec.bin_counts > 0
    max_err, avg_err = O.get_errs(ones[ok], np.ones_like(ones[ok]))
    assert max_err < 1e-8 and avg_err < 1e-11
    hist = np.histogram(kmags, bins=bins, weights=np.abs(fk)**2 * counts * kmags**k_power)[0]
    with np.errstate(invalid="ignore"):
        hist = hist / spec.bin_counts
    sel = ok[1:-2]                       # anisotropic boxes leave some shells empty
    max_err, avg_err = O.get_errs(spectrum[1:-2][sel], hist[1:-2][sel])
    assert max_err < 1e-8 and avg_err < 1e-11

    # position-space entry point: outer axes looped over, normalisation applied
    fx = rng.random((2,) + grid_shape)
    if not fft.is_real:
        fx = fx + 0j
    out = spec(fx.astype(dtype), k_power=k_power)
    assert out.shape == (2, spec.num_bins)
    np_dft = np.fft.rfftn if fft.is_real else np.fft.fftn
    direct = spec.norm * spec.bin_power(np_dft(fx[1]), k_power=k_power)
    assert np.allclose(out[1][ok], direct[ok], rtol=1e-12)


def test_example_initial_conditions_match_oracle(golden, monkeypatch):
    """The IC section of examples/scalar_preheating.py (host backend) == the oracle's
    restatement of the reference example (tests/helpers.py) on identical deviates."""
    import importlib.util
    import os
    from helpers import example_initial_conditions, example_parameters
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        "example_preheating", os.path.join(root, "examples", "scalar_preheating.py"))
    example = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(example)

    grid = (16, 16, 16)
    par = example_parameters(golden)
    ic = example_initial_conditions(par, grid, seed=77)

    p = example.parser.parse_args(["-grid", "16", "16", "16", "--seed", "77"])
    p.grid_shape = tuple(p.grid_shape)
    h = p.halo_shape
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
    # the same NumPy stream as the oracle helper: one (4, kshape) draw per field
    rng = np.random.default_rng(77)
    monkeypatch.setattr(ps.RayleighGenerator, "_uniform",
                        lambda self, shape: torch.from_numpy(rng.random(shape)))
    f0, df0 = [.193 * p.mpl, 0], [-.142231 * p.mpl, 0]
    pad = tuple(n + 2 * h for n in grid)
    f, d = np.zeros((2,) + pad), np.zeros((2,) + pad)
    # background energy as the reference computes it (homogeneous fields)
    V = example.make_potential(p)
    kin = sum(x**2 / 2 for x in df0)
    total = kin + float(V(f0))
    energy = {"total": total, "pressure": kin - float(V(f0))}
    expand = ps.Expansion(total, ps.LowStorageRK54, mpl=p.mpl)
    eff_mass = example.init_rayleigh(p, decomp, f, d, f0, V, energy, expand, backend="fftw")
    assert np.allclose(eff_mass, ic["eff_mass"], rtol=1e-12)
    for i in range(2):
        f[i] += f0[i]
        d[i] += df0[i]
    inner = (slice(None),) + (slice(h, -h),) * 3
    assert np.allclose(f[inner], ic["f"][inner], rtol=1e-12, atol=1e-18)
    assert np.allclose(d[inner], ic["dfdt"][inner], rtol=1e-12, atol=1e-18)
