"""
Projector: the known-answer tests of test/test_projectors.py
on the host backend, with the oracle's finite differences standing in for
FiniteDifferencer (which is CUDA-only).
"""
import numpy as np
import pytest
import torch

import oracle as O
import pystella_b200 as ps
from pystella_b200 import fourier
from pystella_b200.sectors import tensor_index as tid

GRID = (16, 12, 20)
L = (10, 8, 11.5)
H = 2


@pytest.fixture(scope="module")
def setup():
    decomp = ps.DomainDecomposition((1, 1, 1), H, grid_shape=GRID)
    fft = ps.DFT(decomp, None, None, GRID, "float64", backend="fftw")
    dx = tuple(Li / Ni for Li, Ni in zip(L, GRID))
    dk = tuple(2 * np.pi / Li for Li in L)
    project = ps.Projector(fft, ps.FirstCenteredDifference(H).get_eigenvalues, dk, dx)
    return decomp, fft, dx, dk, project


def make_data(fft, rng):
    k = fft.shape(True)
    data = rng.random(k) + 1j * rng.random(k)
    data = fourier.make_hermitian(data)
    return torch.from_numpy(fft.zero_corner_modes(data))


def to_position(fft, fk):
    """-> padded array with halos shared (what FiniteDifferencer sees)."""
    pad = np.zeros(tuple(n + 2 * H for n in GRID))
    fft.idft(fk, pad)
    O.share_halos(pad, H)
    return pad


def divergence_error(fft, dx, vector):
    xs = [to_position(fft, vector[mu]) for mu in range(3)]
    parts = [O.partial(xs[mu], H, dx, mu) for mu in range(3)]
    div = sum(parts)
    norm = sum(np.abs(p) for p in parts)
    return np.max(np.abs(div)) / np.max(norm), np.sum(np.abs(div)) / np.sum(norm)


def test_effective_momenta():
    """test/test_projectors.py:38-66."""
    N, dx, dk = 128, 10 / 128, 2 * np.pi / 10
    kmag = dk * np.linspace(-N // 2 + 1, N // 2 + 1, 100)
    for h in (1, 2, 3, 4):
        st = ps.FirstCenteredDifference(h)
        diff = sum(c * np.exp(1j * kmag * dx * i) - c * np.exp(-1j * kmag * dx * i)
                   for i, c in st.coefs.items())
        assert np.max(np.abs(np.real(diff / dx / 1j) / st.get_eigenvalues(kmag, dx) - 1)) \
            < 1e-13
        st = ps.SecondCenteredDifference(h)
        diff = sum(c * np.exp(1j * kmag * dx * i) + (c * np.exp(-1j * kmag * dx * i)
                                                     if i > 0 else 0)
                   for i, c in st.coefs.items())
        assert np.max(np.abs(np.real(diff / dx**2) / st.get_eigenvalues(kmag, dx) - 1)) < 1e-11


def test_vector_projector(setup):
    """test/test_projectors.py:79-230."""
    decomp, fft, dx, dk, project = setup
    rng = np.random.default_rng(2)
    max_rtol, avg_rtol = 1e-11, 1e-13
    vector = torch.stack([make_data(fft, rng) for _ in range(3)])
    assert divergence_error(fft, dx, vector)[0] > 1e-3
    project.transversify(None, vector)
    max_err, avg_err = divergence_error(fft, dx, vector)
    assert max_err < max_rtol and avg_err < avg_rtol

    plus, minus = make_data(fft, rng), make_data(fft, rng)
    project.pol_to_vec(None, plus, minus, vector)
    max_err, avg_err = divergence_error(fft, dx, vector)
    assert max_err < max_rtol and avg_err < avg_rtol
    vector_2 = torch.zeros_like(vector)
    project.transversify(None, vector, vector_2)
    assert O.get_errs(vector.numpy(), vector_2.numpy())[0] < max_rtol

    plus1, minus1 = torch.zeros_like(plus), torch.zeros_like(minus)
    project.vec_to_pol(None, plus1, minus1, vector)
    # modes with vanishing effective momentum carry no polarisation
    nz = (plus1.abs() > 0).numpy()
    assert O.get_errs(plus.numpy()[nz], plus1.numpy()[nz])[0] < max_rtol
    assert O.get_errs(minus.numpy()[nz], minus1.numpy()[nz])[0] < max_rtol
    project.vec_to_pol(None, vector[0], vector[1], vector)          # in place
    assert O.get_errs(plus1.numpy()[nz], vector[0].numpy()[nz])[0] < max_rtol
    assert O.get_errs(minus1.numpy()[nz], vector[1].numpy()[nz])[0] < max_rtol

    # longitudinal component: lap(long) == div(vector) with the same stencils
    vector = torch.stack([make_data(fft, rng) for _ in range(3)])
    xs = np.stack([to_position(fft, vector[mu]) for mu in range(3)])
    div_true = O.divergence(xs, H, dx)
    lng = torch.zeros_like(minus)
    project.decompose_vector(None, vector, plus1, minus1, lng)
    long_x = to_position(fft, lng)
    grad = O.gradient(long_x, H, dx)
    grad_h = np.zeros((3,) + long_x.shape)
    for mu in range(3):
        O.interior(grad_h[mu], H)[...] = grad[mu]
        O.share_halos(grad_h[mu], H)
    div_long = O.divergence(grad_h, H, dx)
    scale = np.max(np.abs(div_true))
    assert np.max(np.abs(div_long - div_true)) < 1e-10 * scale

    # lng = -i k.v / k^2 pairs with decomp_to_vec(times_abs_k=True), which adds i k lng
    # (test/test_projectors.py:213-222); the |k|-weighted variant pairs the other way
    vector[...] = 0.
    project.decomp_to_vec(None, plus1, minus1, lng, vector, times_abs_k=True)
    xs = np.stack([to_position(fft, vector[mu]) for mu in range(3)])
    assert np.max(np.abs(O.divergence(xs, H, dx) - div_true)) < 1e-10 * scale
    vector = torch.stack([make_data(fft, rng) for _ in range(3)])
    project.transversify(None, vector, vector_2)
    lng_k = torch.zeros_like(lng)
    project.decompose_vector(None, vector, plus1, minus1, lng_k, times_abs_k=True)
    rebuilt = torch.zeros_like(vector)
    project.decomp_to_vec(None, plus1, minus1, lng_k, rebuilt)
    # v = transverse part (plus/minus) + longitudinal part
    project.transversify(None, rebuilt, vector_2)
    only_pol = torch.zeros_like(vector)
    project.pol_to_vec(None, plus1, minus1, only_pol)
    assert np.allclose(vector_2.numpy(), only_pol.numpy(), rtol=1e-10, atol=1e-12)
    keep = (project._k[0] != 0) & (project._k[1] != 0) & (project._k[2] != 0)
    assert np.allclose(rebuilt.numpy()[:, keep.numpy()], vector.numpy()[:, keep.numpy()],
                       rtol=1e-10, atol=1e-12)


def test_tensor_projector(setup):
    """test/test_projectors.py:239-330: transverse, traceless, polarisation round trip."""
    decomp, fft, dx, dk, project = setup
    rng = np.random.default_rng(3)
    hij = torch.stack([make_data(fft, rng) for _ in range(6)])
    project.transverse_traceless(None, hij)
    for i in range(1, 4):
        rows = [hij[tid(i, mu + 1)] for mu in range(3)]
        max_err, avg_err = divergence_error(fft, dx, rows)
        assert max_err < 1e-11 and avg_err < 1e-13
    h = hij.numpy()
    trace = sum(h[tid(i, i)] for i in range(1, 4))
    norm = np.sqrt(sum(np.abs(h[tid(i, i)])**2 for i in range(1, 4)))
    rel = np.abs(trace[norm != 0]) / norm[norm != 0]
    assert np.max(rel[rel < .9]) < 1e-12
    # idempotent
    again = hij.clone()
    project.transverse_traceless(None, hij, again)
    assert np.allclose(again.numpy(), h, rtol=1e-11, atol=1e-14)

    plus, minus = make_data(fft, rng), make_data(fft, rng)
    project.pol_to_tensor(None, plus, minus, hij)
    h = hij.numpy().copy()
    project.transverse_traceless(None, hij)
    assert np.allclose(hij.numpy(), h, rtol=1e-10, atol=1e-13)     # already TT
    p1, m1 = torch.zeros_like(plus), torch.zeros_like(minus)
    project.tensor_to_pol(None, p1, m1, hij)
    nz = (p1.abs() > 0).numpy()
    assert O.get_errs(plus.numpy()[nz], p1.numpy()[nz])[0] < 1e-10
    assert O.get_errs(minus.numpy()[nz], m1.numpy()[nz])[0] < 1e-10


def test_gw_and_polarization_spectra(setup):
    """PowerSpectra.gw == sum of the two polarisation spectra of the TT part; vector
    polarisation spectra of a purely longitudinal field vanish."""
    decomp, fft, dx, dk, project = setup
    rng = np.random.default_rng(4)
    spec = ps.PowerSpectra(decomp, fft, dk, float(np.prod(L)))
    hij = rng.standard_normal((6,) + GRID)
    gw = spec.gw(hij, project, hubble=np.array([0.3]))
    pol = spec.gw_polarization(hij, project, hubble=0.3)
    assert gw.shape == (spec.num_bins,) and pol.shape == (2, spec.num_bins)
    ok = spec.bin_counts > 0
    # sum_ij |h^TT_ij|^2 = |h_+|^2 + |h_-|^2 (unit polarisation tensors eps_i eps_j)
    assert np.allclose(gw[ok], (pol[0] + pol[1])[ok], rtol=1e-9, atol=1e-30)

    vec = rng.standard_normal((2, 3) + GRID)
    out = spec.polarization(vec, project)
    assert out.shape == (2, 2, spec.num_bins)
    dec = spec.vector_decomposition(vec, project)
    assert dec.shape == (2, 3, spec.num_bins)
    assert np.allclose(dec[:, :2][..., ok], out[..., ok], rtol=1e-11)
    assert np.all(dec[:, 2][..., ok][..., 1:-1] > 0)      # corner modes: k_eff = 0
