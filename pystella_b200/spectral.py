"""
Derivatives and a Poisson solver by spectral collocation: :class:`SpectralCollocator`
(``pystella/fourier/derivs.py:29-205``, the ``halo_shape = 0`` alternative to
:class:`~pystella_b200.FiniteDifferencer` with the same call interface) and
:class:`SpectralPoissonSolver` (``pystella/fourier/poisson.py:33-125``).  Transforms are
:class:`~pystella_b200.DFT`'s (cuFFT, slab-distributed); the momentum-space multipliers are
broadcast tensor operations.
"""

from itertools import product

import numpy as np
import torch

__all__ = ["SpectralCollocator", "SpectralPoissonSolver"]

_VIEWS = ((-1, 1, 1), (1, -1, 1), (1, 1, -1))


class SpectralCollocator:
    """``SpectralCollocator(fft, dk)``; ``__call__(queue, fx, *, lap=None, pdx=None,
    pdy=None, pdz=None, grd=None)`` and ``divergence(queue, vec, div)`` as
    FiniteDifferencer.  First derivatives drop the zero and Nyquist modes
    (derivs.py:66-69)."""

    def __init__(self, fft, dk):
        self.fft = fft
        self.grid_size = float(np.prod(fft.grid_shape))
        self._k1, self._k2 = [], []
        for mu in range(3):
            kk = fft._k_host[mu].astype("int")
            full = dk[mu] * kk.astype(fft.rdtype)
            first = full.copy()
            first[abs(kk) == fft.grid_shape[mu] // 2] = 0.
            first[kk == 0] = 0.
            self._k2.append(torch.from_numpy(full).to(fft.device).view(_VIEWS[mu]))
            self._k1.append(torch.from_numpy(first).to(fft.device).view(_VIEWS[mu]))

    def _kmag_sq(self):
        return self._k2[0]**2 + self._k2[1]**2 + self._k2[2]**2

    def __call__(self, queue, fx, *, lap=None, pdx=None, pdy=None, pdz=None, grd=None,
                 allocator=None):
        if grd is not None:
            if isinstance(grd, (tuple, list)):
                pdx, pdy, pdz = grd
            else:
                pdx, pdy, pdz = (grd[..., i, :, :, :] for i in range(3))
        have_grad = pdx is not None and pdy is not None and pdz is not None
        for s in product(*[range(n) for n in fx.shape[:-3]]):
            fk = self.fft.dft(fx[s], queue=queue) * (1 / self.grid_size)
            # the same precedence as the reference's kernel choice (derivs.py:147-162)
            if lap is not None and have_grad:
                wanted = dict(lap=lap, pdx=pdx, pdy=pdy, pdz=pdz)
            elif have_grad:
                wanted = dict(pdx=pdx, pdy=pdy, pdz=pdz)
            elif lap is not None:
                wanted = dict(lap=lap)
            elif pdx is not None:
                wanted = dict(pdx=pdx)
            elif pdy is not None:
                wanted = dict(pdy=pdy)
            elif pdz is not None:
                wanted = dict(pdz=pdz)
            else:
                wanted = {}
            for name, arr in wanted.items():
                if name == "lap":
                    out_k = -self._kmag_sq() * fk
                else:
                    out_k = 1j * self._k1["xyz".index(name[2])] * fk
                self.fft.idft(out_k, arr[s], queue=queue)
        return None

    def divergence(self, queue, vec, div, allocator=None):
        for s in product(*[range(n) for n in vec.shape[:-4]]):
            div_k = 0
            for mu in range(3):
                fk = self.fft.dft(vec[s][mu], queue=queue) * (1 / self.grid_size)
                div_k = div_k + 1j * self._k1[mu] * fk
            self.fft.idft(div_k, div[s], queue=queue)
        return None


class SpectralPoissonSolver:
    """``SpectralPoissonSolver(fft, dk, dx, effective_k)`` solves ``lap f - m^2 f = rho``
    relative to the second-difference stencil whose eigenvalues ``effective_k(k, dx)``
    returns (e.g. ``SecondCenteredDifference(h).get_eigenvalues``); modes with a vanishing
    eigenvalue are set to zero (poisson.py:84-87)."""

    def __init__(self, fft, dk, dx, effective_k):
        self.fft = fft
        self.grid_size = float(np.prod(fft.grid_shape))
        self._ev = []
        for mu in range(3):
            kk = fft._k_host[mu].astype("int")
            ev = np.array(effective_k(dk[mu] * kk.astype(fft.rdtype), dx[mu]),
                          dtype=fft.rdtype)
            self._ev.append(torch.from_numpy(ev).to(fft.device).view(_VIEWS[mu]))

    def __call__(self, queue, fx, rho, m_squared=0, allocator=None):
        rhok = self.fft.dft(rho, queue=queue) * (1 / self.grid_size)
        minus_k_squared = self._ev[0] + self._ev[1] + self._ev[2]
        denom = minus_k_squared - float(m_squared)
        safe = torch.where(denom == 0, torch.ones_like(denom), denom)
        fk = torch.where(minus_k_squared < 0, rhok / safe, torch.zeros_like(rhok))
        self.fft.idft(fk, fx, queue=queue)
