"""
Tensor-perturbation (gravitational-wave) sector on the device: recognises a
:class:`~pystella_b200.TensorPerturbationSector` sourced by scalar sectors
(``pystella/sectors.py:170-208``) and advances it with ``psb_tensor_stage_unfused``.
"""

import ctypes as C
import math
import os

import torch

from pystella_b200 import _lib

__all__ = ["TensorSystem", "tensor_system_from_sector"]


class TensorSystem:
    """hij / dhijdt pairs sourced by ``16 pi sum_n d_i f_n d_j f_n``."""

    def __init__(self, sector, grad_name, nscalars, source_name="f"):
        self.source_name = source_name
        hij = sector.hij
        self.h_name = hij.name
        self.dot_name = hij.dot.name
        self.lap_name = hij.lap.name
        self.grad_name = grad_name
        self.nscalars = nscalars
        self.source_coef = 16 * math.pi

    # -- fused path: stencils evaluated in-kernel (default for FusedStencilStage) ---------
    def _fused_ok(self, stepper, kwargs):
        """The fused tensor stage (``psb_tensor_stage_fused``) applies when the stepper is a
        ``FusedStencilStage`` one, the radius is even, rows are even and the rank layout is
        single-rank or x-slabs with peer-memory halos.  ``PSB_FUSED_GW=0`` forces the
        element-wise stage (which reads ``lap_hij`` / ``dfdx`` from FiniteDifferencer)."""
        if os.environ.get("PSB_FUSED_GW", "1") == "0" or not stepper.fused:
            return False
        decomp = stepper.decomp
        if decomp is not None and decomp.nranks > 1 \
                and getattr(stepper, "_halo_mode", None) != "peer":
            return False
        h = stepper.halo_shape
        hij = kwargs.get(self.h_name)
        if not (h in (2, 4) and hij is not None and hij.shape[-1] % 2 == 0
                and self.nscalars in (1, 2)):
            return False
        # what cuTensorMapEncodeTiled needs (csrc/common.cu, encode_field_tensor_map): rows,
        # planes and components that are multiples of 16 bytes (fp32: Nz + 2h a multiple
        # of 4) and 16-byte aligned arrays; other layouts take the element-wise stage
        es = hij.element_size()
        Px, Py, Pz = hij.shape[-3:]
        arrays = [hij, kwargs.get(self.dot_name)]
        return ((Pz * es) % 16 == 0 and (Py * Pz * es) % 16 == 0
                and (Px * Py * Pz * es) % 16 == 0
                and all(t is not None and t.data_ptr() % 16 == 0 for t in arrays))

    def _run_fused(self, stepper, queue, kwargs, dt, A, B):
        from pystella_b200 import _kernels, fused
        from pystella_b200.step import _scalar_value, gen_tmp_name
        hij, dhij = kwargs[self.h_name], kwargs[self.dot_name]
        for name, t in ((self.h_name, hij), (self.dot_name, dhij)):
            _lib.require_cuda_tensor(t, name)
        if hij.dim() != 4 or hij.shape[0] != 6 or dhij.shape != hij.shape:
            raise ValueError("hij / dhijdt must have shape (6, Nx+2h, Ny+2h, Nz+2h)")
        kh = stepper.tmp_arrays[gen_tmp_name(self.h_name)]
        kd = stepper.tmp_arrays[gen_tmp_name(self.dot_name)]
        # the scalar stage of THIS stage has just swapped its buffers: the twin holds the
        # OLD f, whose gradients source the tensor modes (all right-hand sides are taken from
        # the old state, pystella/step.py:476-485)
        if stepper.__dict__.get("_stage_serial", {}).get(self.source_name) \
                != stepper.__dict__.get("_call_serial"):
            raise ValueError(
                "the fused tensor stage needs the scalar sector's stage of the same call to "
                "have run first: list the ScalarSector before the TensorPerturbationSector")
        old_f = stepper._twins[self.source_name]
        if old_f.dtype != hij.dtype or old_f.shape[-3:] != hij.shape[-3:]:
            raise ValueError("hij and the source fields differ in dtype or padded shape")
        h = stepper.halo_shape
        decomp = stepper.decomp
        multi = decomp is not None and decomp.nranks > 1
        key = self.h_name
        if multi:
            if key not in stepper._peer_state:
                fused._setup_peer_buffers(stepper, key, hij, h)
            elif hij.data_ptr() not in stepper._peer_state[key]["ids"]:
                fused._adopt_into_peer_buffers(stepper, key, hij)
        twin = stepper._twins.get(key)
        if twin is None or twin.shape != hij.shape or twin.dtype != hij.dtype:
            twin = hij.clone()
            stepper._twins[key] = twin
            stepper.__dict__.setdefault("_fresh", {}).pop(key, None)
        if not fused.halos_are_fresh(stepper, key, hij):
            if multi and stepper._peer_barrier_pending:
                stepper._peer.barrier_kernel(queue, stepper.strict_fp, stepper.launch_flags)
                stepper._peer_barrier_pending = False
            fused.fill_halos(stepper, hij, h, queue)
        Nx, Ny, Nz = (n - 2 * h for n in hij.shape[-3:])
        args = _lib.TensorFusedArgs()
        args.dtype = _lib.torch_dtype_code(hij)
        args.nscalars, args.h = self.nscalars, h
        args.wrap_x, args.wrap_y, args.wrap_z = (0 if multi else 1), 1, 1
        args.Nx, args.Ny, args.Nz = Nx, Ny, Nz
        args.hij_in, args.hij_out = hij.data_ptr(), twin.data_ptr()
        args.dhijdt, args.khij, args.kdhijdt = dhij.data_ptr(), kh.data_ptr(), kd.data_ptr()
        args.src = old_f.data_ptr()
        args.src_stride = old_f.shape[-3] * old_f.shape[-2] * old_f.shape[-1]
        for d in range(3):
            args.dx[d] = float(stepper.dx[d])
        args.dt, args.A, args.B = dt, A, B
        args.source_coef = self.source_coef
        tune = stepper.tune or {}
        args.tune_xchunk = tune.get("gw_xchunk", 0)
        args.tune_variant = tune.get("gw_variant", 0)
        args.tune_prefetch = tune.get("gw_prefetch", tune.get("prefetch", 0))
        args.launch_flags = stepper.launch_flags
        hub = kwargs["hubble"]
        if isinstance(hub, torch.Tensor):
            args.hubble_dev = hub.data_ptr()
        else:
            args.hubble = _scalar_value(hub)
        if multi:
            ps = stepper._peer_state[key]
            out_id = ps["ids"][twin.data_ptr()]
            args.peer_lo_hij_out = ps["ptrs"][out_id][ps["lo"]]
            args.peer_hi_hij_out = ps["ptrs"][out_id][ps["hi"]]
            args.peer_lo_Nx = ps["lo_Nx"]
            args.peer_lo_stride, args.peer_hi_stride = ps["lo_stride"], ps["hi_stride"]
            stepper._peer_barrier_pending = True
        lib = _lib.load(stepper.strict_fp)
        rkey = ("gw_rows", tuple(hij.shape), hij.dtype, args.tune_xchunk, args.tune_variant)
        cache = stepper.__dict__.setdefault("_energy_layout", {})
        if rkey not in cache:
            cache[rkey] = lib.psb_tensor_stage_fused_rows(C.byref(args))
            if cache[rkey] < 0:
                _lib.check(lib, 1, "psb_tensor_stage_fused_rows")
        tail = stepper._make_tail(fold=False, dt=dt, nrows=cache[rkey], device=hij.device)
        if tail is not None:
            args.tail = C.pointer(tail)
        rc = lib.psb_tensor_stage_fused(C.byref(args), _lib.stream_ptr(queue))
        _lib.check(lib, rc, "psb_tensor_stage_fused")
        stepper._tail_issued(tail)
        fused._swap_storage(hij, twin)
        fused.mark_halos_fresh(stepper, key, hij)

    def run_stage(self, stepper, stage, queue, kwargs, dt, A, B):
        from pystella_b200.step import _scalar_value, gen_tmp_name
        if "hubble" not in kwargs:
            # the reference's kernel would fail on the missing argument as well
            raise TypeError("missing scalar argument 'hubble' of the tensor sector stage")
        if self._fused_ok(stepper, kwargs):
            return self._run_fused(stepper, queue, kwargs, dt, A, B)
        names = (self.h_name, self.dot_name, self.lap_name, self.grad_name)
        for n in names:
            if kwargs.get(n) is None:
                raise TypeError(
                    f"missing array argument '{n}' (the tensor sector stage is "
                    "element-wise: pass lap_hij and dfdx computed by FiniteDifferencer)")
        hij, dhij, lap, grd = (kwargs[n] for n in names)
        kh = stepper.tmp_arrays[gen_tmp_name(self.h_name)]
        kd = stepper.tmp_arrays[gen_tmp_name(self.dot_name)]
        h = stepper.halo_shape
        for name, t in zip(names, (hij, dhij, lap, grd)):
            _lib.require_cuda_tensor(t, name)
            if t.dtype != hij.dtype:
                raise ValueError(f"'{name}' has dtype {t.dtype}, expected {hij.dtype}")
        if hij.dim() != 4 or hij.shape[0] != 6 or dhij.shape != hij.shape:
            raise ValueError("hij / dhijdt must have shape (6, Nx+2h, Ny+2h, Nz+2h)")
        Nx, Ny, Nz = (n - 2 * h for n in hij.shape[-3:])
        if tuple(lap.shape) != (6, Nx, Ny, Nz):
            raise ValueError(f"lap_hij must have shape {(6, Nx, Ny, Nz)}")
        if tuple(grd.shape) != (self.nscalars, 3, Nx, Ny, Nz):
            raise ValueError(f"dfdx must have shape {(self.nscalars, 3, Nx, Ny, Nz)}")

        args = _lib.TensorStageArgs()
        args.dtype = _lib.torch_dtype_code(hij)
        args.nscalars = self.nscalars
        args.h = h
        args.Nx, args.Ny, args.Nz = Nx, Ny, Nz
        args.hij, args.dhijdt = hij.data_ptr(), dhij.data_ptr()
        args.khij, args.kdhijdt = kh.data_ptr(), kd.data_ptr()
        args.lap_hij, args.dfdx = lap.data_ptr(), grd.data_ptr()
        args.dt, args.A, args.B = dt, A, B
        args.source_coef = self.source_coef
        hub = kwargs["hubble"]
        if isinstance(hub, torch.Tensor):
            args.hubble_dev = hub.data_ptr()
        else:
            args.hubble = _scalar_value(hub)
        lib = _lib.load(stepper.strict_fp)
        rc = lib.psb_tensor_stage_unfused(C.byref(args), _lib.stream_ptr(queue))
        _lib.check(lib, rc, "psb_tensor_stage_unfused")


def tensor_system_from_sector(sector):
    from pystella_b200.sectors import ScalarSector
    if len(sector.sectors) != 1 or not isinstance(sector.sectors[0], ScalarSector):
        raise NotImplementedError(
            "device tensor sector: exactly one ScalarSector source is supported")
    src = sector.sectors[0]
    return TensorSystem(sector, src.f.pd.name, src.nscalars, source_name=src.f.name)
