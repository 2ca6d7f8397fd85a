"""
Thin Python wrappers around the C ABI: they only marshal tensors / scalars into
the argument structs of ``include/pystella_b200.h`` and raise on error.
"""

import contextlib
import ctypes as C

import torch

from pystella_b200 import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def fill_potential(args, dV_terms, nscalars):
    """
    ``dV_terms[n]`` = list of ``(coef, powers)`` for dV/df_n (polynomial in the
    fields).  Writes nterms/coef/power of a ScalarStageArgs.
    """
    for n in range(nscalars):
        terms = dV_terms[n] if dV_terms is not None else []
        if len(terms) > _lib.MAX_TERMS:
            raise ValueError(f"dV/df_{n} has {len(terms)} monomials; at most "
                             f"{_lib.MAX_TERMS} are supported per field")
        args.nterms[n] = len(terms)
        for t, (c, powers) in enumerate(terms):
            args.coef[n][t] = float(c)
            for m in range(nscalars):
                p = int(powers[m])
                if p < 0 or p > 127:
                    raise ValueError("monomial powers must be in 0..127")
                args.power[n][t][m] = p


def scalar_stage(*, fused, f_in, f_out, dfdt, kf, kdfdt, lap_f=None, h, dx=(1., 1., 1.),
                 dt, A, B, a=1., hubble=0., a_dev=None, hubble_dev=None,
                 dV_terms=None, wrap=(True, True, True), x_range=None,
                 queue=None, strict=False, xchunk=0, variant=0,
                 energy_partials=None, energy_slot_offset=0, V_terms=None,
                 count_slots_only=False, probe_plan=False, peer=None, generated=None,
                 tail=None,
                 prefetch=0, pws_sched=None, launch_flags=0):
    """
    One scalar-sector LSRK stage on padded arrays of shape ``(n, Nx+2h, Ny+2h, Nz+2h)``
    (or without the leading axis for a single field).
    """
    lib = _lib.load(strict)
    for name, t in (("f_in", f_in), ("f_out", f_out), ("dfdt", dfdt), ("kf", kf),
                    ("kdfdt", kdfdt)):
        _lib.require_cuda_tensor(t, name)
        if t.dtype != f_in.dtype or t.shape != f_in.shape:
            raise ValueError(f"{name} must match f in dtype and shape")
    if f_in.dim() == 3:
        nscalars = 1
    elif f_in.dim() == 4:
        nscalars = f_in.shape[0]
    else:
        raise ValueError("scalar fields must have shape (n, Nx+2h, Ny+2h, Nz+2h)")
    if not 1 <= nscalars <= _lib.MAX_SCALARS:
        raise NotImplementedError(
            f"{nscalars} scalar fields per launch (max {_lib.MAX_SCALARS})")
    Px, Py, Pz = f_in.shape[-3:]
    Nx, Ny, Nz = Px - 2 * h, Py - 2 * h, Pz - 2 * h

    args = _lib.ScalarStageArgs()
    args.dtype = _lib.torch_dtype_code(f_in)
    args.nscalars = nscalars
    args.h = h
    args.wrap_x, args.wrap_y, args.wrap_z = (int(bool(w)) for w in wrap)
    args.Nx, args.Ny, args.Nz = Nx, Ny, Nz
    if x_range is None:
        x_range = (0, Nx)
    if len(x_range) == 2 and isinstance(x_range[0], (tuple, list)):
        (args.x_begin, args.x_end), (args.x2_begin, args.x2_end) = x_range
    else:
        args.x_begin, args.x_end = x_range
    args.field_stride = Px * Py * Pz
    args.f_in = _ptr(f_in)
    args.f_out = _ptr(f_out)
    args.dfdt = _ptr(dfdt)
    args.kf = _ptr(kf)
    args.kdfdt = _ptr(kdfdt)
    if lap_f is not None:
        _lib.require_cuda_tensor(lap_f, "lap_f")
        if lap_f.shape[-3:] != (Nx, Ny, Nz) or lap_f.dtype != f_in.dtype:
            raise ValueError("lap_f must have the un-padded shape and dtype of f")
        args.lap_f = _ptr(lap_f)
        args.lap_stride = Nx * Ny * Nz
    for d in range(3):
        args.dx[d] = float(dx[d])
    args.dt, args.A, args.B = float(dt), float(A), float(B)
    args.a, args.hubble = float(a), float(hubble)
    args.a_dev = _ptr(a_dev)
    args.hubble_dev = _ptr(hubble_dev)
    if peer is not None:
        # dict(lo=ptr, hi=ptr, lo_Nx=.., lo_stride=.., hi_stride=..): peer-mapped f_out
        args.peer_lo_f_out = C.c_void_p(peer["lo"])
        args.peer_hi_f_out = C.c_void_p(peer["hi"])
        args.peer_lo_Nx = peer["lo_Nx"]
        args.peer_lo_stride = peer["lo_stride"]
        args.peer_hi_stride = peer["hi_stride"]
    fill_potential(args, dV_terms, nscalars)
    args.tune_xchunk = xchunk
    args.tune_variant = variant
    args.tune_prefetch = int(prefetch)
    args.launch_flags = int(launch_flags)
    if pws_sched is not None:
        args.pws_sched = _ptr(pws_sched)
    if tail is not None:
        if not fused:
            raise ValueError("the in-kernel stage tail exists for the fused stage only")
        args.tail = C.pointer(tail)
    if V_terms is not None:
        if len(V_terms) > _lib.MAX_TERMS:
            raise ValueError(f"V has {len(V_terms)} monomials; at most {_lib.MAX_TERMS}")
        args.nvterms = len(V_terms)
        for t, (c, powers) in enumerate(V_terms):
            args.vcoef[t] = float(c)
            for m in range(nscalars):
                pw = int(powers[m])
                if pw < 0 or pw > 127:
                    raise ValueError("monomial powers must be in 0..127")
                args.vpower[t][m] = pw
    if probe_plan:
        # can the TMA pipeline (and hence its run-time generated instances) serve this
        # layout?  Dry run of psb_scalar_stage_fused_plan: geometry only, no driver call
        raw = C.create_string_buffer(C.sizeof(_lib.FusedPlan) + 64)
        plan = _lib.FusedPlan.from_address((C.addressof(raw) + 63) // 64 * 64)
        plan.reserved = 1
        return lib.psb_scalar_stage_fused_plan(C.byref(args), C.byref(plan)) == 0
    if count_slots_only:
        n = lib.psb_scalar_stage_energy_slots(C.byref(args))
        if n < 0:
            _lib.check(lib, 1, "psb_scalar_stage_energy_slots")
        return n
    if energy_partials is not None:
        if not fused:
            raise ValueError("energy partials come from the fused stage only")
        args.energy_partials = _ptr(energy_partials)
        args.energy_slot_offset = energy_slot_offset
    if generated is not None:
        # EXPERIMENTAL: the same pipeline, compiled at run time with a generated RHS
        if not fused:
            raise ValueError("generated right-hand sides exist for the fused stage only")
        launcher, values = generated
        launcher.launch(lib, args, queue, values, strict)
        return
    fn = lib.psb_scalar_stage_fused if fused else lib.psb_scalar_stage_unfused
    rc = fn(C.byref(args), _lib.stream_ptr(queue))
    _lib.check(lib, rc, fn.__name__)


def finite_difference(fx, *, lap=None, pdx=None, pdy=None, pdz=None, h, halo=None,
                      dx, accumulate=False, queue=None, strict=False, variant=0):
    lib = _lib.load(strict)
    _lib.require_cuda_tensor(fx, "fx")
    halo = h if halo is None else halo
    Nx, Ny, Nz = (n - 2 * halo for n in fx.shape[-3:])
    for name, t in (("lap", lap), ("pdx", pdx), ("pdy", pdy), ("pdz", pdz)):
        if t is None:
            continue
        _lib.require_cuda_tensor(t, name)
        if tuple(t.shape) != (Nx, Ny, Nz) or t.dtype != fx.dtype:
            raise ValueError(
                f"{name} has shape {tuple(t.shape)}/{t.dtype}; expected {(Nx, Ny, Nz)}/"
                f"{fx.dtype}")
    dxa = (C.c_double * 3)(*[float(d) for d in dx])
    rc = lib.psb_finite_difference_v(_ptr(fx), _ptr(lap), _ptr(pdx), _ptr(pdy), _ptr(pdz),
                                     _lib.torch_dtype_code(fx), Nx, Ny, Nz, h, halo,
                                     C.byref(dxa), int(accumulate), int(variant),
                                     _lib.stream_ptr(queue))
    _lib.check(lib, rc, "psb_finite_difference")


def halo_wrap(arr, halo_shape, axes_mask=7, queue=None, strict=False):
    lib = _lib.load(strict)
    _lib.require_cuda_tensor(arr, "fx")
    hx, hy, hz = halo_shape
    Nx, Ny, Nz = (n - 2 * h for n, h in zip(arr.shape[-3:], halo_shape))
    rc = lib.psb_halo_wrap(_ptr(arr), _lib.torch_dtype_code(arr), Nx, Ny, Nz,
                           hx, hy, hz, axes_mask, _lib.stream_ptr(queue))
    _lib.check(lib, rc, "psb_halo_wrap")


def halo_pack(arr, buf, halo_shape, axis, queue=None, unpack=False, strict=False):
    lib = _lib.load(strict)
    _lib.require_cuda_tensor(arr, "fx")
    _lib.require_cuda_tensor(buf, "buf")
    hx, hy, hz = halo_shape
    Nx, Ny, Nz = (n - 2 * h for n, h in zip(arr.shape[-3:], halo_shape))
    if unpack:
        rc = lib.psb_halo_unpack(_ptr(arr), _ptr(buf), _lib.torch_dtype_code(arr),
                                 Nx, Ny, Nz, hx, hy, hz, axis, _lib.stream_ptr(queue))
    else:
        rc = lib.psb_halo_pack(_ptr(arr), _ptr(buf), _lib.torch_dtype_code(arr),
                               Nx, Ny, Nz, hx, hy, hz, axis, _lib.stream_ptr(queue))
    _lib.check(lib, rc, "psb_halo_unpack" if unpack else "psb_halo_pack")


def remove_halos(padded, unpadded, halo_shape, queue=None, restore=False, strict=False):
    lib = _lib.load(strict)
    _lib.require_cuda_tensor(padded, "padded")
    _lib.require_cuda_tensor(unpadded, "unpadded")
    hx, hy, hz = halo_shape
    Nx, Ny, Nz = unpadded.shape[-3:]
    if tuple(padded.shape[-3:]) != (Nx + 2 * hx, Ny + 2 * hy, Nz + 2 * hz):
        raise ValueError("padded/unpadded shapes do not differ by 2*halo_shape")
    if padded.dtype != unpadded.dtype:
        raise ValueError("in_array and out_array have different dtypes")
    code = _lib.torch_dtype_code(padded)
    if restore:
        rc = lib.psb_restore_halos(_ptr(unpadded), _ptr(padded), code, Nx, Ny, Nz,
                                   hx, hy, hz, _lib.stream_ptr(queue))
    else:
        rc = lib.psb_remove_halos(_ptr(padded), _ptr(unpadded), code, Nx, Ny, Nz,
                                  hx, hy, hz, _lib.stream_ptr(queue))
    _lib.check(lib, rc, "psb_restore_halos" if restore else "psb_remove_halos")


_reduce_ws = {}


def reduce_poly(arrays, halos, terms, ops, flags, shape, queue=None, strict=False,
                variant=0):
    """
    ``arrays``: list of tensors; ``halos``: their padding; ``terms``: list of
    ``(output, coef, powers)``; ``ops``: per-output op code (0 sum, 1 max, 2 min).
    Returns a float64 device tensor with one value per output.
    """
    lib = _lib.load(strict)
    if len(arrays) > _lib.MAX_RED_ARRAYS:
        raise NotImplementedError("too many distinct arrays in one reduction")
    if len(terms) > _lib.MAX_RED_TERMS or len(ops) > _lib.MAX_RED_OUTPUTS:
        raise NotImplementedError("too many terms/outputs in one reduction")
    dev = arrays[0].device
    args = _lib.ReduceArgs()
    args.dtype = _lib.torch_dtype_code(arrays[0])
    args.narrays, args.nterms, args.noutputs = len(arrays), len(terms), len(ops)
    args.Nx, args.Ny, args.Nz = shape
    for m, (t, h) in enumerate(zip(arrays, halos)):
        _lib.require_cuda_tensor(t, f"array {m}")
        if t.dtype != arrays[0].dtype:
            raise ValueError("all arrays of a reduction must share one dtype")
        if tuple(t.shape[-3:]) != tuple(n + 2 * h for n in shape):
            raise ValueError(f"array {m} has shape {tuple(t.shape)}, expected padding {h}")
        args.arrays[m] = t.data_ptr()
        args.halo[m] = h
    for t, (out, coef, powers) in enumerate(terms):
        args.term_output[t] = out
        args.coef[t] = float(coef)
        for m, p in enumerate(powers):
            args.power[t][m] = int(p)
    for r, (op, fl) in enumerate(zip(ops, flags)):
        args.op[r] = op
        args.flags[r] = fl
    key = (dev.index, _lib.stream_ptr(queue))   # the stream the kernels actually run on
    if key not in _reduce_ws:
        n = lib.psb_reduce_workspace_doubles(_lib.MAX_RED_OUTPUTS)
        _reduce_ws[key] = torch.empty(n, dtype=torch.float64, device=dev)
    out = torch.empty(len(ops), dtype=torch.float64, device=dev)
    args.tune_variant = int(variant)
    args.partials = _reduce_ws[key].data_ptr()
    args.out = out.data_ptr()
    rc = lib.psb_reduce(C.byref(args), _lib.stream_ptr(queue))
    _lib.check(lib, rc, "psb_reduce")
    return out


def bin_power(fk, momenta, dk, bin_width, k_power, nyquist, num_bins, is_real,
              queue=None, blocks=0):
    """
    Binning pass of ``PowerSpectra.bin_power`` (``psb_bin_power``): ``fk`` is a contiguous
    complex CUDA tensor ``(Kx, Ky, Kz)``, ``momenta`` the three 1-D real mode-number tensors
    of ``DFT.sub_k``.  Returns a float64 device tensor of ``num_bins`` sums.
    """
    lib = _lib.load(False)
    _lib.require_cuda_tensor(fk, "fk")
    if fk.dim() != 3:
        raise ValueError("fk must be a 3-D momentum-space array")
    if fk.dtype == torch.complex128:
        code, rdtype = _lib.PSB_F64, torch.float64
    elif fk.dtype == torch.complex64:
        code, rdtype = _lib.PSB_F32, torch.float32
    else:
        raise TypeError(f"fk has dtype {fk.dtype} (complex128/complex64 only)")
    args = _lib.BinPowerArgs()
    args.dtype, args.is_real = code, int(bool(is_real))
    args.Kx, args.Ky, args.Kz = fk.shape
    args.fk = fk.data_ptr()
    stream_ctx = torch.cuda.stream(queue) if queue is not None else contextlib.nullcontext()
    keep = []
    with stream_ctx:       # temporaries are allocated on the stream the kernel runs on
        for name, m, n in zip(("momenta_x", "momenta_y", "momenta_z"), momenta, fk.shape):
            _lib.require_cuda_tensor(m, name)
            if m.dtype != rdtype:
                m = m.to(rdtype)
                keep.append(m)
            if m.numel() != n:
                raise ValueError(f"{name} has {m.numel()} entries, fk has extent {n}")
            setattr(args, name, m.data_ptr())
        spectrum = torch.zeros(int(num_bins), dtype=torch.float64, device=fk.device)
    for mu in range(3):
        args.dk[mu] = float(dk[mu])
    args.bin_width, args.k_power, args.nyquist = float(bin_width), float(k_power), float(nyquist)
    args.num_bins, args.tune_blocks = int(num_bins), int(blocks)
    args.spectrum = spectrum.data_ptr()
    rc = lib.psb_bin_power(C.byref(args), _lib.stream_ptr(queue))
    _lib.check(lib, rc, "psb_bin_power")
    return spectrum


ENERGY_SUMS_DOUBLES = 16 + 64 * 16 + 2   # PSB_ENERGY_SUMS_DOUBLES of the header


def sentinel_rows(nrows, device):
    """``(nrows, 16)`` float64 rows holding PSB_TAIL_SENTINEL: the state in which the
    in-kernel stage tail expects (and leaves) its rows."""
    rows = torch.empty((int(nrows), _lib.ENERGY_STRIDE), dtype=torch.float64, device=device)
    rows.view(torch.int64).fill_(_lib.TAIL_SENTINEL)
    return rows


def energy_sums(partials, nslots, nvals, sums, queue=None, strict=False):
    """``sums``: zero-initialised float64 buffer of ENERGY_SUMS_DOUBLES (result in [:16])."""
    lib = _lib.load(strict)
    if sums.dtype != torch.float64 or sums.numel() < ENERGY_SUMS_DOUBLES:
        raise ValueError(f"sums must be a float64 buffer of {ENERGY_SUMS_DOUBLES} elements "
                         "(result + scratch rows + ticket)")
    if partials.shape[0] < nslots:
        raise ValueError("fewer partial rows than nslots")
    rc = lib.psb_energy_sums(_ptr(partials), nslots, nvals, _ptr(sums),
                             _lib.stream_ptr(queue))
    _lib.check(lib, rc, "psb_energy_sums")


def expansion_stage(sums, nscalars, grid_size, mpl, dt, A, B, state, energies, queue=None,
                    strict=False):
    lib = _lib.load(strict)
    rc = lib.psb_expansion_stage(_ptr(sums), nscalars, float(grid_size), float(mpl),
                                 float(dt), float(A), float(B), _ptr(state),
                                 _ptr(energies), _lib.stream_ptr(queue))
    _lib.check(lib, rc, "psb_expansion_stage")


def expansion_init(sums, nscalars, grid_size, mpl, a0, state, energies, queue=None,
                   strict=False):
    lib = _lib.load(strict)
    rc = lib.psb_expansion_init(_ptr(sums), nscalars, float(grid_size), float(mpl),
                                float(a0), _ptr(state), _ptr(energies),
                                _lib.stream_ptr(queue))
    _lib.check(lib, rc, "psb_expansion_init")


def expansion_stage_peer(peer_args, sums, nscalars, grid_size, mpl, dt, A, B, state,
                         energies, queue=None, strict=False, launch_flags=0):
    """All-gather of the sums over peer memory + barrier + expansion stage (one kernel);
    ``nscalars == 0``: barrier only."""
    lib = _lib.load(strict)
    rc = lib.psb_expansion_stage_peer_v(C.byref(peer_args), _ptr(sums), nscalars,
                                        float(grid_size), float(mpl), float(dt), float(A),
                                        float(B), _ptr(state), _ptr(energies),
                                        int(launch_flags), _lib.stream_ptr(queue))
    _lib.check(lib, rc, "psb_expansion_stage_peer_v")
