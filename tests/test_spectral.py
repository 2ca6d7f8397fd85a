"""SpectralCollocator / SpectralPoissonSolver on the host backend: plane-wave derivatives
(test/test_derivs.py with h = 0) and the Poisson known-answer test (test/test_poisson.py),
the oracle's stencils standing in for FiniteDifferencer when h > 0."""
import numpy as np
import pytest

import oracle as O
import pystella_b200 as ps

GRID = (16, 12, 20)
L = (3, 5, 7)


def make(h):
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=GRID)
    fft = ps.DFT(decomp, None, None, GRID, "float64", backend="fftw")
    dx = tuple(Li / Ni for Li, Ni in zip(L, GRID))
    dk = tuple(2 * np.pi / Li for Li in L)
    return decomp, fft, dx, dk


def test_spectral_derivatives_of_plane_waves():
    decomp, fft, dx, dk = make(0)
    derivs = ps.SpectralCollocator(fft, dk)
    modes = (-3, 2, 5)
    kvec = [dki * m for dki, m in zip(dk, modes)]
    x = np.meshgrid(*[np.arange(n) * d for n, d in zip(GRID, dx)], indexing="ij")
    phase = sum(k * xi for k, xi in zip(kvec, x))
    fx = np.stack([np.sin(phase), np.cos(phase)])
    lap = np.zeros_like(fx)
    grd = np.zeros((2, 3) + GRID)
    derivs(None, fx, lap=lap, grd=grd)
    ksq = sum(k**2 for k in kvec)
    assert np.allclose(lap, -ksq * fx, rtol=1e-10, atol=1e-10)
    for mu in range(3):
        assert np.allclose(grd[0, mu], kvec[mu] * np.cos(phase), atol=1e-11)
        assert np.allclose(grd[1, mu], -kvec[mu] * np.sin(phase), atol=1e-11)
    # single outputs and the reference's precedence: lap wins over a lone pdx
    pdy = np.zeros(GRID)
    derivs(None, fx[0], pdy=pdy)
    assert np.allclose(pdy, grd[0, 1], atol=1e-12)
    lap1, pdx = np.zeros(GRID), np.full(GRID, 7.)
    derivs(None, fx[0], lap=lap1, pdx=pdx)
    assert np.allclose(lap1, lap[0], atol=1e-12) and np.all(pdx == 7.)
    # divergence of the gradient == Laplacian
    div = np.zeros(GRID)
    derivs.divergence(None, grd[0], div)
    assert np.allclose(div, lap[0], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("h", [0, 1, 2])
def test_spectral_poisson(h):
    """test/test_poisson.py:38-95: lap(f) - m^2 f == rho with the matching Laplacian."""
    rng = np.random.default_rng(8)
    decomp, fft, dx, dk = make(h)
    if h == 0:
        evals = lambda k, dx: -k**2                                     # noqa: E731
        derivs = ps.SpectralCollocator(fft, dk)
    else:
        evals = ps.SecondCenteredDifference(h).get_eigenvalues
    solver = ps.SpectralPoissonSolver(fft, dk, dx, evals)
    rho = rng.random(GRID)
    rho -= rho.mean()
    fx = np.zeros(tuple(n + 2 * h for n in GRID))
    for m_squared in (0, 1.2, 19.2):
        solver(None, fx, rho, m_squared=m_squared)
        if h == 0:
            lap = np.zeros(GRID)
            derivs(None, fx, lap=lap)
            fi = fx
        else:
            O.share_halos(fx, h)
            lap = O.laplacian(fx, h, dx)
            fi = O.interior(fx, h)
        diff = np.abs(lap - rho - m_squared * fi)
        assert np.max(diff) / np.max(np.abs(rho)) < 1e-12
        assert np.sum(diff) / np.sum(np.abs(rho)) < 1e-13
