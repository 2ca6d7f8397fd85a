"""
Tensor-perturbation (gravitational-wave) sector on the device: recognises a
:class:`~pystella_b200.TensorPerturbationSector` sourced by scalar sectors
(``pystella/sectors.py:170-208``) and advances it with ``psb_tensor_stage_unfused``.
"""

import ctypes as C
import math

import torch

from pystella_b200 import _lib

__all__ = ["TensorSystem", "tensor_system_from_sector"]


class TensorSystem:
    """hij / dhijdt pairs sourced by ``16 pi sum_n d_i f_n d_j f_n``."""

    def __init__(self, sector, grad_name, nscalars):
        hij = sector.hij
        self.h_name = hij.name
        self.dot_name = hij.dot.name
        self.lap_name = hij.lap.name
        self.grad_name = grad_name
        self.nscalars = nscalars
        self.source_coef = 16 * math.pi

    def run_stage(self, stepper, stage, queue, kwargs, dt, A, B):
        from pystella_b200.step import _scalar_value, gen_tmp_name
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
        hub = kwargs.get("hubble", 0.)
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
    return TensorSystem(sector, src.f.pd.name, src.nscalars)
