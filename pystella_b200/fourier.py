"""
Fourier-space companions of the time loop: :class:`DFT`,
:class:`RayleighGenerator`, :class:`PowerSpectra` with the call contracts of
``pystella/fourier/dft.py``, ``rayleigh.py`` and ``spectra.py``.

These run once per output slice (or once per run), not per stage: the transforms are
cuFFT through ``torch.fft`` (a library call, like the reference's VkFFT/clFFT), the
mode generation and binning are a handful of device-wide tensor operations.  Conventions
follow the reference exactly: forward and backward transforms are both UNNORMALISED
(``test/test_dft.py:92-104``), the last axis of a real transform holds the non-negative
half, the Nyquist wavenumber is positive (``dft.py:326-332``).

``backend="fftw"`` selects host transforms on CPU tensors / NumPy arrays, the analogue of
the reference's ``pDFT`` (``dft.py:350-428``); every other backend value means cuFFT on
the current CUDA device.

Distributed transforms follow mpi4py-fft, which the reference uses (``dft.py:391-407``).
Slab decompositions ``proc_shape = (P, 1, 1)`` (the layout of the fused stepper): position
space split along x, momentum space along y; local 2-D transforms over (y, z), one
all-to-all transpose (NCCL on GPUs, gloo on the host backend), 1-D transforms over x.
Pencil decompositions ``(Px, Py, 1)``: position space split along x and y, momentum space
along y (over the ``Px`` ranks of a column) and z (over the ``Py`` ranks of a row); 1-D
transforms along z, y, x with a transpose inside the row and one inside the column
communicator in between.  ``sub_k`` and all k-space arrays are the rank-local slices.
"""

from itertools import product

import numpy as np
import torch
import torch.distributed as dist

from pystella_b200 import _lib
from pystella_b200._lib import PsbError

__all__ = ["DFT", "RayleighGenerator", "PowerSpectra", "fftfreq", "make_hermitian"]


def fftfreq(n):
    """Integer wavenumbers of an ``n``-point transform, Nyquist positive (dft.py:326-332)."""
    k = np.fft.fftfreq(n, 1 / n)
    if n % 2 == 0:
        k[n // 2] = abs(k[n // 2])
    return k


def _real_dtype(dtype):
    dtype = np.dtype(dtype)
    return np.dtype({"complex128": "float64", "complex64": "float32"}.get(dtype.name,
                                                                          dtype.name))


def _complex_dtype(dtype):
    return np.dtype({"float64": "complex128", "float32": "complex64"}[_real_dtype(dtype).name])


def _tdtype(dt):
    return getattr(torch, np.dtype(dt).name)


class DFT:
    """
    ``DFT(decomp, context, queue, grid_shape, dtype, backend=None)`` (dft.py:41-93;
    ``context``/``queue`` are accepted and ignored -- the stream is the current one).

    Attributes as in ``BaseDFT`` (dft.py:103-120): ``fx``, ``fk`` (default buffers),
    ``sub_k`` (``{"momenta_x": ..., "momenta_y": ..., "momenta_z": ...}``, device
    tensors of integer wavenumbers), ``grid_shape``, ``dtype``, ``rdtype``, ``cdtype``,
    ``is_real``.
    """

    def __init__(self, decomp, context=None, queue=None, grid_shape=None, dtype="float64",
                 backend=None, device=None, **kwargs):
        if grid_shape is None:
            raise TypeError("grid_shape is required")
        if "use_fftw" in kwargs and kwargs.pop("use_fftw"):
            backend = "fftw"
        if decomp.proc_shape[2] != 1:
            raise NotImplementedError("decomposition in z")
        if backend not in (None, "vkfft", "clfft", "cufft", "fftw"):
            raise NotImplementedError(f"{backend} backend for DFTs.")
        self.decomp = decomp
        self.backend = "fftw" if backend == "fftw" else "cufft"
        self.grid_shape = tuple(int(n) for n in grid_shape)
        self.dtype = np.dtype(dtype)
        self.is_real = self.dtype.kind == "f"
        self.rdtype = _real_dtype(self.dtype)
        self.cdtype = _complex_dtype(self.dtype)
        if device is not None:
            self.device = torch.device(device)
        elif self.backend == "fftw":
            self.device = torch.device("cpu")
        else:
            if not torch.cuda.is_available():
                raise PsbError("DFT: no CUDA device (pass backend='fftw' for host "
                               "transforms of NumPy arrays)")
            self.device = torch.device("cuda", torch.cuda.current_device())
        # position space: x split over Px, y over Py; momentum space: y split over Px (the
        # ranks of a column), z over Py (the ranks of a row); uneven rule of
        # decomp.py:322-337 (the first N % P ranks hold one more)
        self.nranks, self.rank = decomp.nranks, decomp.rank
        Px, Py = int(decomp.proc_shape[0]), int(decomp.proc_shape[1])
        self._px, self._py = Px, Py
        self._rx, self._ry = int(decomp.rank_tuple[0]), int(decomp.rank_tuple[1])

        def split(n, parts):
            q, r = divmod(n, parts)
            return [q + (1 if i < r else 0) for i in range(parts)]

        nzk = self.grid_shape[2] // 2 + 1 if self.is_real else self.grid_shape[2]
        self._x_sizes = split(self.grid_shape[0], Px)        # position x, by rx
        self._ypos_sizes = split(self.grid_shape[1], Py)     # position y, by ry
        self._y_sizes = split(self.grid_shape[1], Px)        # momentum y, by rx
        self._zk_sizes = split(nzk, Py)                      # momentum z, by ry
        self._y_start = sum(self._y_sizes[:self._rx])
        self._zk_start = sum(self._zk_sizes[:self._ry])
        if min(self._x_sizes + self._ypos_sizes + self._y_sizes + self._zk_sizes) == 0:
            raise ValueError(f"grid {self.grid_shape} too small for proc_shape "
                             f"{tuple(decomp.proc_shape)}")
        # communicators: the column (same ry, all rx) and, for pencils, the row
        self._col_group, self._row_group = decomp.comm, None
        if Py > 1:
            for ry in range(Py):
                g = dist.new_group([decomp.rankID(rx, ry, 0) for rx in range(Px)])
                if ry == self._ry:
                    self._col_group = g
            for rx in range(Px):
                g = dist.new_group([decomp.rankID(rx, ry, 0) for ry in range(Py)])
                if rx == self._rx:
                    self._row_group = g
        self.fx = torch.zeros(self.shape(False), dtype=_tdtype(self.dtype),
                              device=self.device)
        self.fk = torch.zeros(self.shape(True), dtype=_tdtype(self.cdtype),
                              device=self.device)
        k = [fftfreq(n) for n in self.grid_shape]
        if self.is_real:
            k[-1] = np.fft.rfftfreq(self.grid_shape[-1], 1 / self.grid_shape[-1])
        k[1] = k[1][self._y_start:self._y_start + self._y_sizes[self._rx]]
        k[2] = k[2][self._zk_start:self._zk_start + self._zk_sizes[self._ry]]
        self._k_host = [ki.astype(self.rdtype) for ki in k]
        self.sub_k = {name: torch.from_numpy(ki).to(self.device)
                      for name, ki in zip(("momenta_x", "momenta_y", "momenta_z"),
                                          self._k_host)}

    @property
    def proc_permutation(self):
        return tuple(range(len(self.grid_shape)))

    def shape(self, forward_output=True):
        """Rank-local shape of momentum-space (``True``) or position-space arrays
        (dft.py:415-416, 483-489)."""
        nx, ny, nz = self.grid_shape
        if not forward_output:
            return (self._x_sizes[self._rx], self._ypos_sizes[self._ry], nz)
        return (nx, self._y_sizes[self._rx], self._zk_sizes[self._ry])

    # -- transposes -----------------------------------------------------------
    def _redistribute(self, arr, group, me, full_axis, split_axis, full_sizes, split_sizes):
        """``arr`` is complete along ``full_axis`` and holds slice ``me`` of
        ``split_axis`` (which is cut into ``split_sizes`` over ``group``); returns the
        array complete along ``split_axis`` holding slice ``me`` of ``full_axis`` cut into
        ``full_sizes``.  One all-to-all; complex data travel as pairs of reals."""
        pieces, start = [], 0
        for n in full_sizes:
            pieces.append(torch.view_as_real(arr.narrow(full_axis, start, n).contiguous())
                          .reshape(-1))
            start += n
        send = torch.cat(pieces)
        shapes = []
        for n in split_sizes:
            shp = list(arr.shape)
            shp[full_axis] = full_sizes[me]
            shp[split_axis] = n
            shapes.append(shp)
        out_splits = [2 * int(np.prod(shp)) for shp in shapes]
        recv = torch.empty(sum(out_splits), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send, out_splits, [p.numel() for p in pieces],
                               group=group)
        parts, off = [], 0
        for shp, n in zip(shapes, out_splits):
            parts.append(torch.view_as_complex(recv[off:off + n].view(*shp, 2)))
            off += n
        return torch.cat(parts, dim=split_axis)

    def _forward(self, x):
        if self.nranks == 1:
            return (torch.fft.rfftn(x, dim=(-3, -2, -1)) if self.is_real
                    else torch.fft.fftn(x, dim=(-3, -2, -1)))
        if self._py == 1:       # slab: (nx_l, Ny, Nz) -> 2-D transform -> x <-> y
            a = (torch.fft.rfftn(x, dim=(-2, -1)) if self.is_real
                 else torch.fft.fftn(x, dim=(-2, -1)))
        else:                   # pencil: z, then y after the row transpose
            a = torch.fft.rfft(x, dim=-1) if self.is_real else torch.fft.fft(x, dim=-1)
            a = self._redistribute(a, self._row_group, self._ry, 2, 1, self._zk_sizes,
                                   self._ypos_sizes)
            a = torch.fft.fft(a, dim=1)
        a = self._redistribute(a, self._col_group, self._rx, 1, 0, self._y_sizes,
                               self._x_sizes)
        return torch.fft.fft(a, dim=0)

    def _backward(self, k):
        if self.nranks == 1:
            if self.is_real:
                return torch.fft.irfftn(k, s=self.grid_shape, dim=(-3, -2, -1),
                                        norm="forward")
            return torch.fft.ifftn(k, dim=(-3, -2, -1), norm="forward")
        a = torch.fft.ifft(k, dim=0, norm="forward")
        a = self._redistribute(a, self._col_group, self._rx, 0, 1, self._x_sizes,
                               self._y_sizes)
        if self._py == 1:
            if self.is_real:
                return torch.fft.irfftn(a, s=self.grid_shape[1:], dim=(-2, -1),
                                        norm="forward")
            return torch.fft.ifftn(a, dim=(-2, -1), norm="forward")
        a = torch.fft.ifft(a, dim=1, norm="forward")
        a = self._redistribute(a, self._row_group, self._ry, 1, 2, self._ypos_sizes,
                               self._zk_sizes)
        if self.is_real:
            return torch.fft.irfft(a, n=self.grid_shape[2], dim=-1, norm="forward")
        return torch.fft.ifft(a, dim=-1, norm="forward")

    # -- array plumbing -------------------------------------------------------
    def _halo(self):
        return tuple(self.decomp.halo_shape)

    def _position_input(self, fx, queue):
        """-> unpadded tensor on ``self.device`` holding ``fx`` (dft.py:171-184)."""
        if fx is None:
            return self.fx
        local = self.shape(False)
        if tuple(fx.shape) != local:
            padded = tuple(n + 2 * h for n, h in zip(local, self._halo()))
            if tuple(fx.shape) != padded:
                raise ValueError(f"array of shape {tuple(fx.shape)} is neither the rank's "
                                 f"grid {local} nor its padded grid {padded}")
            if tuple(self.fx.shape) != local:
                self.fx = torch.zeros(local, dtype=_tdtype(self.dtype), device=self.device)
            if isinstance(fx, torch.Tensor) and fx.is_cuda and self.fx.is_cuda:
                self.decomp.remove_halos(queue, fx, self.fx)
            else:
                inner = tuple(slice(h, h + n) for n, h in zip(local, self._halo()))
                self.fx.copy_(self._to_device(fx[inner]))
            return self.fx
        if isinstance(fx, torch.Tensor) and fx.device == self.device:
            return fx
        if tuple(self.fx.shape) != local:
            self.fx = torch.zeros(local, dtype=_tdtype(self.dtype), device=self.device)
        self.fx.copy_(self._to_device(fx))
        return self.fx

    def _to_device(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(self.device)
        return torch.from_numpy(np.ascontiguousarray(x)).to(self.device)

    @staticmethod
    def _store(dst, src):
        """``dst[...] = src`` across NumPy / CPU / CUDA."""
        if isinstance(dst, torch.Tensor):
            dst.copy_(src)
        else:
            dst[...] = src.cpu().numpy()
        return dst

    # -- transforms -----------------------------------------------------------
    def dft(self, fx=None, fk=None, queue=None, **kwargs):
        """Forward transform of ``fx`` (padded or not, device tensor or NumPy array) into
        ``fk`` or the attached :attr:`fk`; unnormalised (dft.py:138-206)."""
        out = self._forward(self._position_input(fx, queue))
        if fk is None:
            self.fk = out
            return self.fk
        return self._store(fk, out)

    def idft(self, fk=None, fx=None, queue=None, **kwargs):
        """Backward transform, unnormalised; a padded ``fx`` has its interior written
        (halos are not filled, as ``restore_halos``) (dft.py:208-290)."""
        out = self._backward(self.fk if fk is None else self._to_device(fk))
        if fx is None:
            self.fx = out
            return self.fx
        local = self.shape(False)
        if tuple(fx.shape) == local:
            return self._store(fx, out)
        padded = tuple(n + 2 * h for n, h in zip(local, self._halo()))
        if tuple(fx.shape) != padded:
            raise ValueError(f"array of shape {tuple(fx.shape)} is neither the rank's grid "
                             f"{local} nor its padded grid {padded}")
        if isinstance(fx, torch.Tensor) and fx.is_cuda and out.is_cuda:
            self.decomp.restore_halos(queue, out.contiguous(), fx)
        else:
            inner = tuple(slice(h, h + n) for n, h in zip(local, self._halo()))
            self._store(fx[inner], out)
        return fx

    def zero_corner_modes(self, array, only_imag=False):
        """Zero the modes whose integer wavenumbers are all 0 or Nyquist, or only their
        imaginary parts (dft.py:292-323)."""
        where = []
        for mu in range(3):
            kk = self._k_host[mu].astype("int")
            where.append([int(i) for i in np.flatnonzero(
                (abs(kk) == 0) | (abs(kk) == self.grid_shape[mu] // 2))])
        for i, j, k in product(*where):
            if only_imag:
                if isinstance(array, torch.Tensor):
                    array[i, j, k] = array[i, j, k].real.clone()
                else:
                    array[i, j, k] = array[i, j, k].real
            else:
                array[i, j, k] = 0.
        return array


def _mirror_planes(plane):
    """One ``k_z`` plane of :func:`make_hermitian`, in place, statements in the
    reference's order (later ones overwrite earlier ones)."""
    nx, ny = plane.shape
    px, py = np.arange(nx // 2 + 1), np.arange(ny // 2 + 1)
    mx, my = (-px) % nx, (-py) % ny              # the mirrored (negative) wavenumbers
    for a, b in zip(mx, px):
        plane[a, my] = plane[b, py].conj()
        plane[b, my] = plane[a, py].conj()
    for a, b in zip(my, py):
        plane[mx, a] = plane[px, b].conj()
        plane[mx, b] = plane[px, a].conj()


def make_hermitian(fk):
    """
    Enforce ``f(-k) = conj(f(k))`` on the two self-conjugate planes ``k_z = 0`` and
    ``k_z = N_z/2`` of a half-complex array and make the eight corner modes real
    (rayleigh.py:35-54).  Only those two planes are touched; they are processed on the
    host (2 N_x N_y numbers) and written back.
    """
    nz = 2 * (fk.shape[-1] - 1)
    is_tensor = isinstance(fk, torch.Tensor)
    for kz in (0, nz // 2):
        plane = fk[:, :, kz].cpu().numpy().copy() if is_tensor else fk[:, :, kz]
        _mirror_planes(plane)
        for i in (0, fk.shape[0] // 2):
            for j in (0, fk.shape[1] // 2):
                plane[i, j] = plane[i, j].real
        if is_tensor:
            fk[:, :, kz] = torch.from_numpy(plane).to(fk.device)
    return fk


class RayleighGenerator:
    """
    Gaussian random fields with a chosen power spectrum, drawn mode by mode from the
    Rayleigh distribution (rayleigh.py:57-395).  ``RayleighGenerator(context, fft, dk,
    volume, seed=13298)``; ``context`` is ignored.  Random numbers come from torch's
    Philox generator on the transform's device (the reference uses pyopencl's Threefry),
    so realisations differ from the reference's for the same seed.

    Generated modes are returned as tensors on the transform's device (the reference
    returns host arrays).
    """

    def __init__(self, context, fft, dk, volume, seed=13298):
        self.fft = fft
        self.dtype, self.rdtype, self.cdtype = fft.dtype, fft.rdtype, fft.cdtype
        self.volume = volume
        self.device = fft.device
        kvecs = np.meshgrid(*fft._k_host, indexing="ij", sparse=False)
        self.kmags = np.sqrt(sum((dki * ki)**2 for dki, ki in zip(dk, kvecs)))
        self.rng = torch.Generator(device=self.device)
        self.rng.manual_seed(int(seed))

    # test hook: tests substitute their own deviates to compare with the oracle
    def _uniform(self, shape):
        r = torch.rand(shape, generator=self.rng, dtype=_tdtype(self.rdtype),
                       device=self.device)
        return torch.where(r == 0, torch.full_like(r, 0.5), r)   # log(0) guard

    def _post_process(self, fk):
        """rayleigh.py:160-171: Hermitian symmetry on the self-conjugate planes, corner
        modes real."""
        if self.fft.is_real:
            if self.fft.nranks == 1:
                # single-device transforms need the symmetry imposed by hand; the slab
                # transform's final c2r pass discards the offending parts, as in the
                # reference, which symmetrises for pyclDFT only (rayleigh.py:162-166)
                fk = make_hermitian(fk)
            self.fft.zero_corner_modes(fk, only_imag=True)
        return fk

    @staticmethod
    def _ps_wrapper(ps_func, wk, kmags):
        """Evaluate the spectrum with the homogeneous mode's 1/0 avoided and its power
        set to zero (rayleigh.py:173-184)."""
        homogeneous = kmags[0, 0, 0] == 0.
        if homogeneous:
            saved = wk[0, 0, 0]
            # any valid neighbour (the local z extent can be 1 on a pencil)
            other = (0, 0, 1) if wk.shape[2] > 1 else (0, 1, 0) if wk.shape[1] > 1 \
                else (1, 0, 0)
            wk[0, 0, 0] = wk[other]
        power = np.asarray(ps_func(wk), dtype=np.float64) * np.ones_like(kmags)
        if homogeneous:
            power[0, 0, 0] = 0.
            wk[0, 0, 0] = saved
        return power

    def _dev(self, host):
        return torch.from_numpy(np.ascontiguousarray(host, dtype=self.rdtype)).to(self.device)

    def generate(self, queue=None, random=True, field_ps=lambda kmag: 1 / 2 / kmag, norm=1,
                 window=lambda kmag: 1.):
        """Fourier modes ``sqrt(P) * sqrt(-log u1) * exp(2 pi i u2)`` (rayleigh.py:84-97,
        186-226)."""
        rands = self._uniform((2,) + self.kmags.shape)
        if not random:
            rands[0] = np.exp(-1)
        f_power = (norm / self.volume * window(self.kmags)**2
                   * self._ps_wrapper(field_ps, self.kmags, self.kmags))
        amp = torch.sqrt(-torch.log(rands[0]))
        phs = torch.polar(torch.ones_like(rands[1]), 2. * np.pi * rands[1])
        fk = (phs * (amp * torch.sqrt(self._dev(f_power)))).to(_tdtype(self.cdtype))
        return self._post_process(fk)

    def init_field(self, fx, queue=None, **kwargs):
        """rayleigh.py:228-246."""
        self.fft.idft(self.generate(queue, **kwargs), fx, queue=queue)

    def generate_WKB(self, queue=None, random=True, field_ps=lambda wk: 1 / 2 / wk, norm=1,
                     omega_k=lambda kmag: kmag, hubble=0., window=lambda kmag: 1.):
        """Modes of a field and of its conformal-time derivative in the WKB approximation:
        left- and right-moving waves ``L, R``; ``f = (L + R)/sqrt2``,
        ``f' = i w (L - R)/sqrt2 - H f`` (rayleigh.py:57-82, 307-362)."""
        rands = self._uniform((4,) + self.kmags.shape)
        if not random:
            rands[0] = rands[2] = np.exp(-1)
        wk = np.array(omega_k(self.kmags), dtype=np.float64) * np.ones_like(self.kmags)
        f_power = (norm / self.volume * window(self.kmags)**2
                   * self._ps_wrapper(field_ps, wk, self.kmags))
        root_p = torch.sqrt(self._dev(f_power))
        one = torch.ones_like(rands[1])
        left = torch.polar(one, 2. * np.pi * rands[1]) * (torch.sqrt(-torch.log(rands[0]))
                                                         * root_p)
        right = torch.polar(one, 2. * np.pi * rands[3]) * (torch.sqrt(-torch.log(rands[2]))
                                                          * root_p)
        sqrt2 = float(np.sqrt(2.))
        fk = (left + right) / sqrt2
        dfk = 1j * self._dev(wk) * (left - right) / sqrt2 - float(hubble) * fk
        ctype = _tdtype(self.cdtype)
        return self._post_process(fk.to(ctype)), self._post_process(dfk.to(ctype))

    def init_WKB_fields(self, fx, dfx, queue=None, **kwargs):
        """rayleigh.py:364-383."""
        fk, dfk = self.generate_WKB(queue, **kwargs)
        self.fft.idft(fk, fx, queue=queue)
        self.fft.idft(dfk, dfx, queue=queue)


class PowerSpectra:
    """
    Binned power spectra (spectra.py:29-253): ``PowerSpectra(decomp, fft, dk, volume,
    bin_width=min(dk))``; ``__call__(fx, queue=None, k_power=3)`` loops over the outer
    axes of ``fx`` and returns a NumPy array of shape ``outer + (num_bins,)``;
    ``bin_power(fk, k_power=3)`` bins a momentum-space array; ``polarization``,
    ``vector_decomposition``, ``gw`` and ``gw_polarization`` take a
    :class:`~pystella_b200.Projector`.
    """

    def __init__(self, decomp, fft, dk, volume, **kwargs):
        self.decomp, self.fft = decomp, fft
        self.grid_shape = fft.grid_shape
        self.dtype, self.rdtype, self.cdtype = fft.dtype, fft.rdtype, fft.cdtype
        self.kshape = fft.shape(True)
        self.dk = dk
        self.bin_width = kwargs.pop("bin_width", min(dk))
        d3x = volume / np.prod(self.grid_shape)
        self.norm = (1 / 2 / np.pi**2 / volume) * d3x**2          # spectra.py:73-74
        kvecs = np.meshgrid(*fft._k_host, indexing="ij", sparse=False)
        kmags = np.sqrt(sum((dki * ki)**2 for dki, ki in zip(dk, kvecs)))
        if fft.is_real:                                           # spectra.py:80-85
            counts = 2. * np.ones_like(kmags)
            counts[kvecs[2] == 0] = 1.
            counts[kvecs[2] == self.grid_shape[-1] // 2] = 1.
        else:
            counts = np.ones_like(kmags)
        max_k = float(decomp.allreduce(float(np.max(kmags)), op="max"))   # spectra.py:87-91
        self.num_bins = int(max_k / self.bin_width + .5) + 1
        edges = np.arange(-.5, self.num_bins + .5) * self.bin_width
        self.bin_counts = decomp.allreduce(
            np.histogram(kmags, weights=counts, bins=edges)[0])
        # device arrays are binned by csrc/diag.cu, which forms |k|, the bin and the
        # multiplicity on the fly (spectra.py:103-117: bin = round(|k| / bin_width), C
        # rounding; weight = count * |k|**n * |f_k|**2, count = 2 strictly inside
        # 0 < k_z < N_z/2); the per-mode tables below serve host arrays only
        self._tables = None

    def _host_tables(self):
        if self._tables is None:
            fft = self.fft
            kvecs = np.meshgrid(*fft._k_host, indexing="ij", sparse=False)
            kmags = np.sqrt(sum((dki * ki)**2 for dki, ki in zip(self.dk, kvecs)))
            if fft.is_real:
                kz = kvecs[2]
                mult = np.where((kz > 0) & (kz < self.grid_shape[-1] / 2), 2., 1.)
            else:
                mult = np.ones_like(kmags)
            self._tables = (
                torch.from_numpy(np.floor(kmags / self.bin_width + .5).astype(np.int64)
                                 .reshape(-1)),
                torch.from_numpy(mult.reshape(-1)), torch.from_numpy(kmags.reshape(-1)))
        return self._tables

    def bin_power(self, fk, queue=None, k_power=3, allocator=None):
        """-> NumPy array: the unnormalised binned spectrum of ``fk`` divided by the bin
        multiplicities (spectra.py:132-163)."""
        k = self.fft._to_device(fk)
        if _lib.on_device(k):
            # one pass over fk, |k| formed on the fly from the 1-D momenta (csrc/diag.cu)
            from pystella_b200 import _kernels
            k = k.contiguous().reshape(self.kshape)
            spectrum = _kernels.bin_power(
                k, [self.fft.sub_k[n] for n in ("momenta_x", "momenta_y", "momenta_z")],
                self.dk, self.bin_width, k_power, self.grid_shape[-1] / 2, self.num_bins,
                self.fft.is_real, queue=queue)
        else:
            # host arrays (the decomposition's host test backend)
            k = k.reshape(-1)
            bins, mult, kmag = self._host_tables()
            weight = mult * kmag**k_power * (k.real.double()**2 + k.imag.double()**2)
            spectrum = torch.zeros(self.num_bins, dtype=torch.float64, device=k.device)
            spectrum.index_add_(0, bins, weight)
        if self.decomp.comm is not None:
            spectrum = self.decomp.allreduce(spectrum)
        with np.errstate(divide="ignore", invalid="ignore"):
            return spectrum.cpu().numpy() / self.bin_counts

    def __call__(self, fx, queue=None, k_power=3, allocator=None):
        """``Delta_f^2(k) = 1/(2 pi^2 V) \\int dOmega |k|^n |f(k)|^2`` (spectra.py:165-215)."""
        outer = tuple(fx.shape[:-3])
        out = np.zeros(outer + (self.num_bins,), dtype=np.float64)
        for s in product(*[range(n) for n in outer]):
            fk = self.fft.dft(fx[s], queue=queue)
            out[s] = self.bin_power(fk, queue, k_power) * self.norm
        return out

    def _transform_components(self, arrays, n, queue):
        ks = torch.empty((n,) + tuple(self.kshape), dtype=_tdtype(self.cdtype),
                         device=self.fft.device)
        for mu in range(n):
            self.fft.dft(arrays[mu], ks[mu], queue=queue)
        return ks

    def polarization(self, vector, projector, queue=None, k_power=3, allocator=None):
        """Spectra of the plus and minus polarisations of a vector field; outer axes are
        looped over: ``vector.shape[:-4] + (2, num_bins)`` (spectra.py:217-262)."""
        outer = tuple(vector.shape[:-4])
        out = np.zeros(outer + (2, self.num_bins), dtype=np.float64)
        for s in product(*[range(n) for n in outer]):
            vec_k = self._transform_components(vector[s], 3, queue)
            projector.vec_to_pol(queue, vec_k[0], vec_k[1], vec_k)
            out[s][0] = self.bin_power(vec_k[0], queue, k_power)
            out[s][1] = self.bin_power(vec_k[1], queue, k_power)
        return self.norm * out

    def vector_decomposition(self, vector, projector, queue=None, k_power=3,
                             allocator=None):
        """As :meth:`polarization` plus the longitudinal component (times ``|k|``):
        ``vector.shape[:-4] + (3, num_bins)`` (spectra.py:264-312)."""
        outer = tuple(vector.shape[:-4])
        out = np.zeros(outer + (3, self.num_bins), dtype=np.float64)
        for s in product(*[range(n) for n in outer]):
            vec_k = self._transform_components(vector[s], 3, queue)
            projector.decompose_vector(queue, vec_k, vec_k[0], vec_k[1], vec_k[2],
                                       times_abs_k=True)
            for n in range(3):
                out[s][n] = self.bin_power(vec_k[n], queue, k_power)
        return self.norm * out

    def gw(self, hij, projector, hubble, queue=None, k_power=3, allocator=None):
        """Spectral abundance of transverse-traceless gravitational waves,
        ``1/(24 pi^2 H^2 V) sum_ij int dOmega |k|^3 |h'_ij(k)|^2``; ``hij`` is the
        length-6 position-space array of time derivatives (spectra.py:314-358)."""
        hij_k = self._transform_components(hij, 6, queue)
        projector.transverse_traceless(queue, hij_k)
        spec = [self.bin_power(hij_k[mu], queue, k_power) for mu in range(6)]
        from pystella_b200.sectors import tensor_index
        total = sum(spec[tensor_index(i, j)] for i in range(1, 4) for j in range(1, 4))
        return self.norm / 12 / float(np.asarray(hubble).reshape(-1)[0])**2 * total

    def gw_polarization(self, hij, projector, hubble, queue=None, k_power=3,
                        allocator=None):
        """Gravitational-wave abundance on the circular polarisation basis, shape
        ``(2, num_bins)`` (spectra.py:360-419)."""
        hij_k = self._transform_components(hij, 6, queue)
        projector.tensor_to_pol(queue, hij_k[0], hij_k[1], hij_k)
        out = np.zeros((2, self.num_bins), dtype=np.float64)
        out[0] = self.bin_power(hij_k[0], queue, k_power)
        out[1] = self.bin_power(hij_k[1], queue, k_power)
        return self.norm / 12 / float(np.asarray(hubble).reshape(-1)[0])**2 * out
