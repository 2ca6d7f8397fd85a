"""
Test infrastructure: an "emulated device" for the CPU test-suite.  The product's launch code
(argument marshalling, layout inference, grid sizes, the ctypes structs) runs unchanged on
HOST tensors; where it would call into the CUDA libraries, stand-ins execute the very same
kernel source on the CPU execution model of tests/cuda_emu/cuda_emu.h:

* generated kernels (`libpystella_b200_jit.so`: psb_jit_launch / psb_jit_launch_params)
  -> the generated source compiled with g++ and run block by block;
* `psb_bin_power` -> csrc/diag_kernels.cuh through tests/cuda_emu/emu_diag.cpp;
* `psb_reduce` (validated on the GPU in tests/test_gpu_*.py) -> a NumPy restatement of its
  contract, enough for the callers' argument tables to be checked.

Nothing here is reachable from the product.
"""
import contextlib
import ctypes as C
import os
import types

import numpy as np
import torch

from test_cuda_emu import EMU, build_emu

NUM_SMS = 3      # small, so that the persistent kernels loop over several row groups per CTA


class EmuJitLibrary:
    """Stands in for libpystella_b200_jit.so: handles are KernelSource objects."""

    def __init__(self):
        self.launches = []

    def _run(self, ksrc, args_ref, grid, block, smem):
        driver = (
            '#include "cuda_emu.h"\n' + ksrc.source +
            'extern "C" void run(const psb_jit_args* a, int gx, int gy, int gz, int bx, '
            'int by, int bz, long long smem) {\n'
            '  emu::launch_smem(dim3(gx, gy, gz), dim3(bx, by, bz), (size_t)smem, '
            f'[&] {{ {ksrc.name}(*a); }});\n}}\n')
        lib = build_emu(driver, "emu_jit")
        self.launches.append((ksrc.name, grid, block, smem))
        lib.run(args_ref, *[C.c_int(int(v)) for v in grid + block], C.c_longlong(int(smem)))
        return 0

    def psb_jit_launch(self, handle, args_ref, gx, gy, gz, bx, by, bz, stream):
        return self._run(handle, args_ref, (gx, gy, gz), (bx, by, bz), 0)

    def psb_jit_launch_params(self, handle, params, gx, gy, gz, bx, by, bz, smem, stream):
        from pystella_b200.codegen import JitArgs
        args = C.cast(params[0], C.POINTER(JitArgs))
        return self._run(handle, args, (gx, gy, gz), (bx, by, bz), smem)

    def psb_jit_last_log(self):
        return b""


class EmuMainLibrary:
    """Stands in for libpystella_b200.so (the entry points the diagnostics use)."""

    def __init__(self):
        src = open(os.path.join(EMU, "emu_diag.cpp")).read()
        self._diag = build_emu(src, "emu_diag", extra=["-I" + os.path.join(
            os.path.dirname(EMU), "..", "include")])

    def psb_bin_power(self, args_ref, stream):
        return self._diag.emu_psb_bin_power(args_ref, C.c_int(NUM_SMS))

    def psb_last_error(self):
        return b""


def numpy_reduce_poly(arrays, halos, terms, ops, flags, shape, queue=None, strict=False,
                      variant=0):
    """Contract of psb_reduce (include/pystella_b200.h): out[r] = op_r over sites of
    sum_t coef_t prod_j arr_j^power_tj, |.| first if flags bit 0."""
    views = [a.numpy()[tuple(slice(h, h + n) for n in shape)] if h else a.numpy()
             for a, h in zip(arrays, halos)]
    out = []
    for r, (op, fl) in enumerate(zip(ops, flags)):
        val = np.zeros(shape)
        for (o, coef, powers) in terms:
            if o == r:
                t = np.full(shape, float(coef))
                for v, p in zip(views, powers):
                    t = t * v.astype(np.float64)**int(p)
                val = val + t
        if fl & 1:
            val = np.abs(val)
        out.append({0: val.sum(), 1: val.max(), 2: val.min()}[op])
    return torch.tensor(out, dtype=torch.float64)


@contextlib.contextmanager
def emulated_device(monkeypatch):
    from pystella_b200 import _kernels, _lib, codegen
    jit, main = EmuJitLibrary(), EmuMainLibrary()
    monkeypatch.setattr(_lib, "on_device", lambda t: True)
    monkeypatch.setattr(_lib, "stream_ptr", lambda queue=None: None)
    monkeypatch.setattr(_lib, "load", lambda strict=False: main)
    monkeypatch.setattr(_kernels, "reduce_poly", numpy_reduce_poly)
    monkeypatch.setattr(codegen.JitLibrary, "compile", classmethod(lambda cls, ksrc: ksrc))
    monkeypatch.setattr(codegen.JitLibrary, "get", classmethod(lambda cls: jit))
    monkeypatch.setattr(torch.cuda, "get_device_properties",
                        lambda dev=None: types.SimpleNamespace(multi_processor_count=NUM_SMS))
    yield types.SimpleNamespace(jit=jit, main=main)
