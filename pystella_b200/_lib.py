"""
ctypes binding of ``libpystella_b200.so`` (the C ABI declared in
``include/pystella_b200.h``).  There is deliberately no fallback: if the shared
object is missing or a call fails, an exception is raised.
"""

import ctypes as C
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

PSB_F64, PSB_F32 = 0, 1
MAX_SCALARS = 4
MAX_TERMS = 16
MAX_RED_ARRAYS = 8
MAX_RED_TERMS = 32
MAX_RED_OUTPUTS = 16
ENERGY_STRIDE = 16

c_void_p, c_int, c_int32, c_int64, c_double = (
    C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_double)


class PeerArgs(C.Structure):
    _fields_ = [("nranks", c_int32), ("rank", c_int32), ("peer_slots", c_void_p),
                ("peer_flags", c_void_p), ("epoch", c_void_p)]


TAIL_FOLD, TAIL_EXPAND, TAIL_PEER = 1, 2, 4
LAUNCH_PDL = 1          # PSB_LAUNCH_PDL
#: bit pattern (a NaN payload) of a not-yet-published value in a stage-tail row
TAIL_SENTINEL = 0x7ff8dead0000beef


class StageTail(C.Structure):
    """psb_stage_tail: what CTA 0 of a fused stage launch does once every CTA has published
    its row (fold energy sums, peer barrier, background stage)."""
    _fields_ = [("mode", c_int32), ("nscalars", c_int32), ("rows", c_void_p),
                ("nrows", c_int64), ("sums", c_void_p),
                ("grid_size", c_double), ("mpl", c_double), ("dt", c_double),
                ("A", c_double), ("B", c_double),
                ("state", c_void_p), ("energies", c_void_p), ("peers", PeerArgs)]


class ScalarStageArgs(C.Structure):
    _fields_ = [
        ("dtype", c_int32), ("nscalars", c_int32), ("h", c_int32),
        ("wrap_x", c_int32), ("wrap_y", c_int32), ("wrap_z", c_int32),
        ("Nx", c_int64), ("Ny", c_int64), ("Nz", c_int64),
        ("x_begin", c_int64), ("x_end", c_int64),
        ("x2_begin", c_int64), ("x2_end", c_int64),
        ("field_stride", c_int64), ("lap_stride", c_int64),
        ("f_in", c_void_p), ("f_out", c_void_p), ("dfdt", c_void_p),
        ("kf", c_void_p), ("kdfdt", c_void_p), ("lap_f", c_void_p),
        ("dx", c_double * 3),
        ("dt", c_double), ("A", c_double), ("B", c_double),
        ("a", c_double), ("hubble", c_double),
        ("a_dev", c_void_p), ("hubble_dev", c_void_p),
        ("peer_lo_f_out", c_void_p), ("peer_hi_f_out", c_void_p),
        ("peer_lo_Nx", c_int64), ("peer_lo_stride", c_int64), ("peer_hi_stride", c_int64),
        ("nterms", c_int32 * MAX_SCALARS),
        ("coef", (c_double * MAX_TERMS) * MAX_SCALARS),
        ("power", ((C.c_int8 * MAX_SCALARS) * MAX_TERMS) * MAX_SCALARS),
        ("energy_partials", c_void_p),
        ("energy_slot_offset", c_int64),
        ("nvterms", c_int32),
        ("vcoef", c_double * MAX_TERMS),
        ("vpower", (C.c_int8 * MAX_SCALARS) * MAX_TERMS),
        ("tune_xchunk", c_int32), ("tune_variant", c_int32),
        ("tail", C.POINTER(StageTail)),
        ("tune_prefetch", c_int32), ("launch_flags", c_int32),
        ("pws_sched", c_void_p),
    ]


class RKStageArgs(C.Structure):
    _fields_ = [("dtype", c_int32), ("nscalars", c_int32), ("h", c_int32),
                ("ncopies", c_int32),
                ("Nx", c_int64), ("Ny", c_int64), ("Nz", c_int64),
                ("f", c_void_p), ("dfdt", c_void_p), ("lap_f", c_void_p),
                ("q_read", c_int32), ("nstatements", c_int32),
                ("dst", c_int32 * 2), ("ncomb", c_int32 * 2),
                ("src", (c_int32 * 2) * 2), ("coef", (c_double * 2) * 2),
                ("dt_coef", c_double * 2),
                ("dt", c_double), ("a", c_double), ("hubble", c_double),
                ("nterms", c_int32 * MAX_SCALARS),
                ("pcoef", (c_double * MAX_TERMS) * MAX_SCALARS),
                ("power", ((C.c_int8 * MAX_SCALARS) * MAX_TERMS) * MAX_SCALARS)]


class TensorStageArgs(C.Structure):
    _fields_ = [("dtype", c_int32), ("nscalars", c_int32), ("h", c_int32),
                ("reserved", c_int32),
                ("Nx", c_int64), ("Ny", c_int64), ("Nz", c_int64),
                ("hij", c_void_p), ("dhijdt", c_void_p), ("khij", c_void_p),
                ("kdhijdt", c_void_p), ("lap_hij", c_void_p), ("dfdx", c_void_p),
                ("dt", c_double), ("A", c_double), ("B", c_double), ("hubble", c_double),
                ("source_coef", c_double), ("hubble_dev", c_void_p)]


class TensorFusedArgs(C.Structure):
    _fields_ = [("dtype", c_int32), ("nscalars", c_int32), ("h", c_int32),
                ("wrap_x", c_int32), ("wrap_y", c_int32), ("wrap_z", c_int32),
                ("Nx", c_int64), ("Ny", c_int64), ("Nz", c_int64),
                ("hij_in", c_void_p), ("hij_out", c_void_p), ("dhijdt", c_void_p),
                ("khij", c_void_p), ("kdhijdt", c_void_p), ("src", c_void_p),
                ("src_stride", c_int64), ("dx", c_double * 3),
                ("dt", c_double), ("A", c_double), ("B", c_double), ("hubble", c_double),
                ("source_coef", c_double), ("hubble_dev", c_void_p),
                ("peer_lo_hij_out", c_void_p), ("peer_hi_hij_out", c_void_p),
                ("peer_lo_Nx", c_int64), ("peer_lo_stride", c_int64),
                ("peer_hi_stride", c_int64),
                ("tail", C.POINTER(StageTail)),
                ("tune_xchunk", c_int32), ("tune_variant", c_int32),
                ("tune_prefetch", c_int32), ("launch_flags", c_int32)]


class ReduceArgs(C.Structure):
    _fields_ = [
        ("dtype", c_int32), ("narrays", c_int32), ("nterms", c_int32),
        ("noutputs", c_int32),
        ("Nx", c_int64), ("Ny", c_int64), ("Nz", c_int64),
        ("arrays", c_void_p * MAX_RED_ARRAYS),
        ("halo", c_int32 * MAX_RED_ARRAYS),
        ("coef", c_double * MAX_RED_TERMS),
        ("power", (C.c_int8 * MAX_RED_ARRAYS) * MAX_RED_TERMS),
        ("term_output", c_int32 * MAX_RED_TERMS),
        ("op", c_int32 * MAX_RED_OUTPUTS),
        ("flags", c_int32 * MAX_RED_OUTPUTS),
        ("partials", c_void_p), ("out", c_void_p),
        ("tune_variant", c_int32), ("reserved0", c_int32),
    ]


class BinPowerArgs(C.Structure):
    """psb_bin_power_args: binning pass of PowerSpectra.bin_power."""
    _fields_ = [
        ("dtype", c_int32), ("is_real", c_int32),
        ("Kx", c_int64), ("Ky", c_int64), ("Kz", c_int64),
        ("fk", c_void_p), ("momenta_x", c_void_p), ("momenta_y", c_void_p),
        ("momenta_z", c_void_p),
        ("dk", c_double * 3),
        ("bin_width", c_double), ("k_power", c_double), ("nyquist", c_double),
        ("num_bins", c_int32), ("tune_blocks", c_int32),
        ("spectrum", c_void_p),
    ]


#: every symbol include/pystella_b200.h declares: (restype, argtypes)
class FusedPlan(C.Structure):
    """psb_fused_plan: launch parameters of the fused stage for an external kernel."""
    _fields_ = [
        ("dtype", c_int32), ("h", c_int32), ("nscalars", c_int32),
        ("by", c_int32), ("tz", c_int32), ("p", c_int32), ("minb", c_int32),
        ("vec", c_int32), ("energy", c_int32),
        ("grid", C.c_uint32 * 3), ("threads", C.c_uint32), ("smem_bytes", C.c_uint32),
        ("params_bytes", C.c_uint32), ("reserved", c_int32),
        ("tensor_maps", (C.c_ubyte * 128) * 4),
        ("params", C.c_ubyte * 4096),
    ]


SYMBOLS = {
    "psb_last_error": (C.c_char_p, []),
    "psb_abi_version": (c_int, []),
    "psb_strict_fp": (c_int, []),
    "psb_scalar_stage_fused": (c_int, [C.POINTER(ScalarStageArgs), c_void_p]),
    "psb_scalar_stage_unfused": (c_int, [C.POINTER(ScalarStageArgs), c_void_p]),
    "psb_scalar_rk_stage": (c_int, [C.POINTER(RKStageArgs), c_void_p]),
    "psb_tensor_stage_fused": (c_int, [C.POINTER(TensorFusedArgs), c_void_p]),
    "psb_tensor_stage_fused_rows": (c_int64, [C.POINTER(TensorFusedArgs)]),
    "psb_tensor_stage_unfused": (c_int, [C.POINTER(TensorStageArgs), c_void_p]),
    "psb_finite_difference": (c_int, [c_void_p] * 5 + [c_int, c_int64, c_int64, c_int64,
                                                        c_int, c_int,
                                                        C.POINTER(c_double * 3), c_int,
                                                        c_void_p]),
    "psb_finite_difference_v": (c_int, [c_void_p] * 5 + [c_int, c_int64, c_int64, c_int64,
                                                          c_int, c_int,
                                                          C.POINTER(c_double * 3), c_int,
                                                          c_int, c_void_p]),
    "psb_halo_wrap": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64,
                              c_int, c_int, c_int, c_int, c_void_p]),
    "psb_halo_pack": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64,
                              c_int, c_int, c_int, c_int, c_void_p]),
    "psb_halo_unpack": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64,
                                c_int, c_int, c_int, c_int, c_void_p]),
    "psb_remove_halos": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64,
                                 c_int, c_int, c_int, c_void_p]),
    "psb_restore_halos": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64,
                                  c_int, c_int, c_int, c_void_p]),
    "psb_reduce_workspace_doubles": (c_int64, [c_int]),
    "psb_reduce": (c_int, [C.POINTER(ReduceArgs), c_void_p]),
    "psb_bin_power": (c_int, [C.POINTER(BinPowerArgs), c_void_p]),
    "psb_scalar_stage_energy_slots": (c_int64, [C.POINTER(ScalarStageArgs)]),
    "psb_scalar_stage_fused_plan": (c_int, [C.POINTER(ScalarStageArgs),
                                            C.POINTER(FusedPlan)]),
    "psb_energy_sums": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "psb_expansion_stage": (c_int, [c_void_p, c_int, c_double, c_double, c_double,
                                    c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "psb_expansion_stage_peer": (c_int, [C.POINTER(PeerArgs), c_void_p, c_int, c_double,
                                         c_double, c_double, c_double, c_double, c_void_p,
                                         c_void_p, c_void_p]),
    "psb_expansion_stage_peer_v": (c_int, [C.POINTER(PeerArgs), c_void_p, c_int, c_double,
                                           c_double, c_double, c_double, c_double,
                                           c_void_p, c_void_p, c_int, c_void_p]),
    "psb_expansion_init": (c_int, [c_void_p, c_int, c_double, c_double, c_double,
                                   c_void_p, c_void_p, c_void_p]),
}


class PsbError(RuntimeError):
    pass


_libs = {}


def lib_path(strict=False):
    name = "libpystella_b200_strict.so" if strict else "libpystella_b200.so"
    if not strict and os.environ.get("PSB_DEV_LIB"):   # development A/B builds
        name = os.environ["PSB_DEV_LIB"]
    return os.path.join(HERE, "_C", name)


def load(strict=False):
    """Load (once) and return the shared library; raises if it is not built."""
    strict = bool(strict) or os.environ.get("PSB_STRICT_FP", "0") == "1"
    if strict not in _libs:
        path = lib_path(strict)
        if not os.path.exists(path):
            raise PsbError(
                f"{path} not found: build the CUDA extension first "
                "(python -m pystella_b200.build); there is no CPU fallback")
        lib = C.CDLL(path)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.psb_abi_version() != 2:
            raise PsbError("ABI version mismatch")
        _libs[strict] = lib
    return _libs[strict]


def check(lib, rc, what):
    if rc != 0:
        msg = lib.psb_last_error().decode(errors="replace")
        raise PsbError(f"{what} failed (code {rc}): {msg}")


def dtype_code(dtype):
    if dtype in (torch.float64, np.float64) or np.dtype(dtype) == np.float64:
        return PSB_F64
    if dtype in (torch.float32, np.float32) or np.dtype(dtype) == np.float32:
        return PSB_F32
    raise TypeError(f"unsupported dtype {dtype}")


def torch_dtype_code(t):
    if t.dtype == torch.float64:
        return PSB_F64
    if t.dtype == torch.float32:
        return PSB_F32
    raise TypeError(f"unsupported dtype {t.dtype} (float64/float32 only)")


def stream_ptr(queue=None):
    """``queue`` is a torch.cuda.Stream (the analogue of a cl.CommandQueue) or None."""
    if queue is None:
        return torch.cuda.current_stream().cuda_stream
    if isinstance(queue, torch.cuda.Stream):
        return queue.cuda_stream
    raise TypeError("queue must be a torch.cuda.Stream or None")


def on_device(t):
    """True for tensors the kernels can take (CUDA tensors).  One place, so that the CPU
    test-suite can run the launch code against the execution model of tests/cuda_emu."""
    return bool(t.is_cuda)


def require_cuda_tensor(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor on a CUDA device, got {type(t)}")
    if not on_device(t):
        raise PsbError(f"{name} lives on {t.device}: the hot path runs on the GPU only "
                       "(no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t
