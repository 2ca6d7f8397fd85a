"""
TEST INFRASTRUCTURE ONLY -- NumPy restatement of the Fourier-space helpers that
initialise the preheating example (row f-4 of SURVEY.md section 8): the DFT
conventions, the momenta, and RayleighGenerator's WKB mode generation.

Every function cites the reference lines (relative to /root/reference) it follows.
The reference draws its uniform random numbers with pyopencl's Threefry generator
(``fourier/rayleigh.py:154``), which is not available here: callers pass the uniform
deviates in, so that oracle and product can be compared on identical draws, and the
stored example constraint (``test/test_examples.py:33``) can only be matched
statistically.
"""

from itertools import product

import numpy as np

__all__ = ["fftfreq", "momenta", "dft", "idft", "make_hermitian", "zero_corner_modes",
           "kmags", "ps_wrapper", "generate_wkb", "generate", "init_wkb_fields"]


def fftfreq(n):
    """Integer wavenumbers with the Nyquist mode made positive (fourier/dft.py:326-332)."""
    freq = np.fft.fftfreq(n, 1 / n)
    if n % 2 == 0:
        freq[n // 2] = np.abs(freq[n // 2])
    return freq


def momenta(grid_shape, real=True):
    """``sub_k`` of a single-rank transform (fourier/dft.py:335-351): the last axis of a
    real transform holds the non-negative half."""
    k = [fftfreq(n) for n in grid_shape]
    if real:
        k[-1] = np.fft.rfftfreq(grid_shape[-1], 1 / grid_shape[-1])
    return k


def dft(fx):
    """Unnormalised forward transform == ``numpy.fft.rfftn`` (test/test_dft.py:54-58,
    96-99)."""
    return np.fft.rfftn(fx)


def idft(fk, grid_shape):
    """Unnormalised backward transform: ``irfftn * grid_size`` (test/test_dft.py:92-94,
    101-104 divide by ``grid_size`` before comparing with NumPy)."""
    return np.fft.irfftn(fk, s=grid_shape, axes=(0, 1, 2)) * np.prod(grid_shape)


def make_hermitian(fk):
    """fourier/rayleigh.py:35-54 (applied for the single-device transforms,
    ``_post_process``, rayleigh.py:160-166).  Statement order kept: later assignments
    overwrite earlier ones exactly as in the reference."""
    grid_shape = list(fk.shape)
    grid_shape[-1] = 2 * (grid_shape[-1] - 1)
    pos = [np.arange(0, n // 2 + 1) for n in grid_shape]
    neg = [np.concatenate([np.array([0]), np.arange(n - 1, n // 2 - 1, -1)])
           for n in grid_shape]
    for k in [0, grid_shape[-1] // 2]:
        for n, p in zip(neg[0], pos[0]):
            fk[n, neg[1], k] = np.conj(fk[p, pos[1], k])
            fk[p, neg[1], k] = np.conj(fk[n, pos[1], k])
        for n, p in zip(neg[1], pos[1]):
            fk[neg[0], n, k] = np.conj(fk[pos[0], p, k])
            fk[neg[0], p, k] = np.conj(fk[pos[0], n, k])
    for i in [0, grid_shape[0] // 2]:
        for j in [0, grid_shape[1] // 2]:
            for k in [0, grid_shape[2] // 2]:
                fk[i, j, k] = np.real(fk[i, j, k])
    return fk


def zero_corner_modes(array, grid_shape, only_imag=False):
    """fourier/dft.py:292-323."""
    sub_k = [k.astype("int") for k in momenta(grid_shape)]
    where_to_zero = []
    for mu in range(3):
        kk = sub_k[mu]
        where_0 = np.argwhere(abs(kk) == 0).reshape(-1)
        where_n2 = np.argwhere(abs(kk) == grid_shape[mu] // 2).reshape(-1)
        where_to_zero.append(np.concatenate([where_0, where_n2]))
    for i, j, k in product(*where_to_zero):
        array[i, j, k] = array[i, j, k].real if only_imag else 0.
    return array


def kmags(grid_shape, dk):
    """``|k|`` on the half-complex grid (fourier/rayleigh.py:150-152)."""
    kvecs = np.meshgrid(*momenta(grid_shape), indexing="ij", sparse=False)
    return np.sqrt(sum((dki * ki)**2 for dki, ki in zip(dk, kvecs)))


def ps_wrapper(ps_func, wk, km):
    """Evaluate the spectrum away from the 1/0 of the homogeneous mode, whose power is
    then set to zero (fourier/rayleigh.py:168-179)."""
    found_zero = km[0, 0, 0] == 0.
    if found_zero:
        wk0 = wk[0, 0, 0]
        wk[0, 0, 0] = wk[0, 0, 1]
    power = ps_func(wk)
    if found_zero:
        power[0, 0, 0] = 0.
        wk[0, 0, 0] = wk0
    return power


def _post_process(fk, grid_shape):
    """fourier/rayleigh.py:160-171 for the single-device real transform."""
    fk = make_hermitian(fk)
    return zero_corner_modes(fk, grid_shape, only_imag=True)


def generate(rands, grid_shape, dk, volume, field_ps=lambda kmag: 1 / 2 / kmag, norm=1,
             window=lambda kmag: 1.):
    """fourier/rayleigh.py:84-97 (kernel) and :181-221; ``rands`` has shape
    ``(2,) + kshape``, uniform in (0, 1)."""
    km = kmags(grid_shape, dk)
    f_power = norm / volume * window(km)**2 * ps_wrapper(field_ps, km, km)
    amp = np.sqrt(-np.log(rands[0]))
    phs = np.exp(1j * 2. * np.pi * rands[1])
    return _post_process(phs * amp * np.sqrt(f_power), grid_shape)


def generate_wkb(rands, grid_shape, dk, volume, field_ps=lambda wk: 1 / 2 / wk, norm=1,
                 omega_k=lambda kmag: kmag, hubble=0., window=lambda kmag: 1.):
    """fourier/rayleigh.py:57-82 (kernel) and :307-362; ``rands`` has shape
    ``(4,) + kshape``, uniform in (0, 1)."""
    km = kmags(grid_shape, dk)
    wk = omega_k(km)
    f_power = norm / volume * window(km)**2 * ps_wrapper(field_ps, wk, km)
    amp_1 = np.sqrt(-np.log(rands[0]))
    amp_2 = np.sqrt(-np.log(rands[2]))
    phs_1 = np.exp(1j * 2. * np.pi * rands[1])
    phs_2 = np.exp(1j * 2. * np.pi * rands[3])
    lmode = phs_1 * amp_1 * np.sqrt(f_power)
    rmode = phs_2 * amp_2 * np.sqrt(f_power)
    sqrt2 = np.sqrt(2.)
    fk = (lmode + rmode) / sqrt2
    dfk = 1j * wk * (lmode - rmode) / sqrt2 - hubble * fk
    return _post_process(fk, grid_shape), _post_process(dfk, grid_shape)


def init_wkb_fields(rands, grid_shape, dk, volume, **kwargs):
    """-> ``(fx, dfx)`` unpadded (fourier/rayleigh.py:364-383)."""
    fk, dfk = generate_wkb(rands, grid_shape, dk, volume, **kwargs)
    return idft(fk, grid_shape), idft(dfk, grid_shape)
