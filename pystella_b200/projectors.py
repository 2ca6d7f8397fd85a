"""
:class:`Projector`: polarisation bases, transverse and transverse-traceless projections in
momentum space, with the call contract of ``pystella/fourier/projectors.py:29-464``.
Needed for gravitational-wave and vector-field spectra (``PowerSpectra.gw``,
``.gw_polarization``, ``.polarization``, ``.vector_decomposition``).

Projections are taken relative to a finite-difference stencil through its effective
momenta (``effective_k(k, dx)``, e.g. ``FirstCenteredDifference(h).get_eigenvalues``), so
that a projected field is transverse with respect to the *discrete* divergence.  Each
method is a handful of broadcast tensor operations over the rank-local momentum-space
arrays (they run once per output slice); every output is formed completely before it is
stored, so the in-place uses of the reference (``plus = vector[0]``...) work.
"""

import numpy as np
import torch

from pystella_b200.sectors import tensor_index as tid

__all__ = ["Projector"]


class Projector:
    """
    ``Projector(fft, effective_k, dk, dx)`` (projectors.py:68-95): ``effective_k`` is a
    callable ``(k, dx) -> k_eff``, or an int ``h`` (centred difference of that radius), or
    ``0`` (spectral, ``k_eff = k``).  Effective momenta of the zero and Nyquist modes are
    zero.
    """

    def __init__(self, fft, effective_k, dk, dx):
        self.fft = fft
        if not callable(effective_k):
            if effective_k != 0:
                from pystella_b200.derivs import FirstCenteredDifference
                effective_k = FirstCenteredDifference(effective_k).get_eigenvalues
            else:
                def effective_k(k, dx):  # noqa: ARG001
                    return k
        self.eff_mom = {}
        views = ((-1, 1, 1), (1, -1, 1), (1, 1, -1))
        self._k = []
        for mu, name in enumerate(("eff_mom_x", "eff_mom_y", "eff_mom_z")):
            kk = fft._k_host[mu].astype("int")
            eff = np.array(effective_k(dk[mu] * kk.astype(fft.rdtype), dx[mu]),
                           dtype=fft.rdtype)
            eff[abs(kk) == fft.grid_shape[mu] // 2] = 0.
            eff[kk == 0] = 0.
            t = torch.from_numpy(eff).to(fft.device)
            self.eff_mom[name] = t
            self._k.append(t.view(views[mu]))

    # -- pieces ---------------------------------------------------------------
    def _kvec_zero(self, tol=1e-14):
        kx, ky, kz = self._k
        return (kx.abs() < tol) & (ky.abs() < tol) & (kz.abs() < tol)

    def _eps(self):
        """Circular polarisation vector (projectors.py:135-155)."""
        kx, ky, kz = self._k
        kmag = torch.sqrt(kx**2 + ky**2 + kz**2)
        kappa = torch.sqrt(kx**2 + ky**2)
        kxy0 = (kx.abs() < 1e-10) & (ky.abs() < 1e-10)
        kz_nonzero = kz.abs() > 1e-10
        safe_kmag = torch.where(kmag == 0, torch.ones_like(kmag), kmag)
        safe_kappa = torch.where(kappa == 0, torch.ones_like(kappa), kappa)
        r2 = 2**.5
        ctype = getattr(torch, np.dtype(self.fft.cdtype).name)
        zero = torch.zeros((), dtype=ctype, device=kx.device)
        on_axis0 = torch.where(kz_nonzero, zero + 1 / r2, zero)
        on_axis1 = torch.where(kz_nonzero, zero + 1j / r2, zero)
        e0 = torch.where(kxy0, on_axis0,
                         torch.complex(kx * kz / safe_kmag, -ky + 0 * kz) / safe_kappa / r2)
        e1 = torch.where(kxy0, on_axis1,
                         torch.complex(ky * kz / safe_kmag, kx + 0 * kz) / safe_kappa / r2)
        e2 = torch.where(kxy0, zero, (-kappa / safe_kmag / r2).to(ctype))
        return e0, e1, e2

    @staticmethod
    def _store(dst, values):
        for d, v in zip(dst, values):
            d.copy_(v)

    # -- vectors --------------------------------------------------------------
    def transversify(self, queue, vector, vector_T=None):
        """Remove the longitudinal part: ``v - k (k.v)/k^2`` (projectors.py:118-124)."""
        if vector_T is None:
            vector_T = vector
        k = self._k
        ksq = k[0]**2 + k[1]**2 + k[2]**2
        safe = torch.where(ksq == 0, torch.ones_like(ksq), ksq)
        div = k[0] * vector[0] + k[1] * vector[1] + k[2] * vector[2]
        zero = self._kvec_zero()
        out = [torch.where(zero, torch.zeros_like(div), vector[mu] - k[mu] / safe * div)
               for mu in range(3)]
        self._store([vector_T[mu] for mu in range(3)], out)

    def pol_to_vec(self, queue, plus, minus, vector):
        """``v_mu = plus eps_mu + minus conj(eps_mu)`` (projectors.py:171-177)."""
        eps = self._eps()
        out = [plus * eps[mu] + minus * eps[mu].conj() for mu in range(3)]
        self._store([vector[mu] for mu in range(3)], out)

    def vec_to_pol(self, queue, plus, minus, vector):
        """``plus = v.conj(eps)``, ``minus = v.eps`` (projectors.py:158-169)."""
        eps = self._eps()
        p = sum(vector[mu] * eps[mu].conj() for mu in range(3))
        m = sum(vector[mu] * eps[mu] for mu in range(3))
        self._store([plus, minus], [p, m])

    def decompose_vector(self, queue, vector, plus, minus, lng, *, times_abs_k=False):
        """Polarisations and longitudinal component ``-i (k.v)/k^2`` (or ``/|k|``)
        (projectors.py:197-209)."""
        eps = self._eps()
        k = self._k
        ksq = k[0]**2 + k[1]**2 + k[2]**2
        safe = torch.where(ksq == 0, torch.ones_like(ksq), ksq)
        div = k[0] * vector[0] + k[1] * vector[1] + k[2] * vector[2]
        p = sum(vector[mu] * eps[mu].conj() for mu in range(3))
        m = sum(vector[mu] * eps[mu] for mu in range(3))
        denom = safe**.5 if times_abs_k else safe
        lg = torch.where(self._kvec_zero(), torch.zeros_like(div), -1j * div / denom)
        self._store([plus, minus, lng], [p, m, lg])

    def decomp_to_vec(self, queue, plus, minus, lng, vector, *, times_abs_k=False):
        """Inverse of :meth:`decompose_vector` (projectors.py:179-195)."""
        eps = self._eps()
        k = self._k
        kmag = torch.sqrt(k[0]**2 + k[1]**2 + k[2]**2)
        safe = torch.where(kmag == 0, torch.ones_like(kmag), kmag)
        zero = self._kvec_zero()
        out = []
        for mu in range(3):
            along = 1j * k[mu] * lng if times_abs_k else 1j * k[mu] / safe * lng
            out.append(plus * eps[mu] + minus * eps[mu].conj()
                       + torch.where(zero, torch.zeros_like(along), along))
        self._store([vector[mu] for mu in range(3)], out)

    # -- tensors --------------------------------------------------------------
    def transverse_traceless(self, queue, hij, hij_TT=None):
        """``(P_ac P_db - P_ab P_cd / 2) h_cd`` with ``P_ab = delta_ab - khat_a khat_b``
        (projectors.py:211-241)."""
        if hij_TT is None:
            hij_TT = hij
        k = self._k
        kmag = torch.sqrt(k[0]**2 + k[1]**2 + k[2]**2)
        safe = torch.where(kmag == 0, torch.ones_like(kmag), kmag)
        khat = [kk / safe for kk in k]
        P = {}
        for a in range(1, 4):
            for b in range(a, 4):
                P[tid(a, b)] = (1. if a == b else 0.) - khat[a - 1] * khat[b - 1]
        zero = self._kvec_zero()
        out = {}
        for a in range(1, 4):
            for b in range(a, 4):
                acc = 0
                for c in range(1, 4):
                    for d in range(1, 4):
                        acc = acc + (P[tid(a, c)] * P[tid(d, b)]
                                     - P[tid(a, b)] * P[tid(c, d)] / 2) * hij[tid(c, d)]
                out[tid(a, b)] = torch.where(zero, torch.zeros_like(acc), acc)
        for n, v in out.items():
            hij_TT[n].copy_(v)

    def tensor_to_pol(self, queue, plus, minus, hij):
        """projectors.py:243-252."""
        eps = self._eps()
        p = sum(hij[tid(c, d)] * eps[c - 1].conj() * eps[d - 1].conj()
                for c in range(1, 4) for d in range(1, 4))
        m = sum(hij[tid(c, d)] * eps[c - 1] * eps[d - 1]
                for c in range(1, 4) for d in range(1, 4))
        self._store([plus, minus], [p, m])

    def pol_to_tensor(self, queue, plus, minus, hij):
        """projectors.py:254-262."""
        eps = self._eps()
        out = {tid(a, b): plus * eps[a - 1] * eps[b - 1]
               + minus * eps[a - 1].conj() * eps[b - 1].conj()
               for a in range(1, 4) for b in range(a, 4)}
        for n, v in out.items():
            hij[n].copy_(v)
