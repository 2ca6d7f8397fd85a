"""
Test infrastructure: an "emulated device" for the CPU test-suite.  The product's launch code
(argument marshalling, layout inference, grid sizes, the ctypes structs) runs unchanged on
HOST tensors; where it would call into the CUDA libraries, stand-ins execute the very same
kernel source on the CPU execution model of tests/cuda_emu/cuda_emu.h:

* generated kernels (`libpystella_b200_jit.so`: psb_jit_launch / psb_jit_launch_params)
  -> the generated source compiled with g++ and run block by block;
* `libpystella_b200.so` -> a HOST build of its stand-alone translation units (halo, element-
  wise stages, finite differences (generic tile kernel), reductions, background,
  diagnostics) with the same C ABI (tests/emu_library.py).  The TMA pipelines (fused
  stages) are hardware-specific and have no emulated counterpart.

Nothing here is reachable from the product.
"""
import contextlib
import ctypes as C
import os
import types

import numpy as np
import torch

from test_cuda_emu import build_emu

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


def emulated_main_library():
    """Stands in for libpystella_b200.so: the HOST build of its stand-alone translation
    units (tests/emu_library.py), with the ctypes prototypes of pystella_b200._lib."""
    import emu_library
    from pystella_b200 import _lib
    lib = emu_library.build()
    for name, (restype, argtypes) in _lib.SYMBOLS.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype, fn.argtypes = restype, argtypes
    return lib


@contextlib.contextmanager
def emulated_device(monkeypatch):
    from pystella_b200 import _kernels, _lib, codegen
    jit, main = EmuJitLibrary(), emulated_main_library()
    monkeypatch.setattr(_lib, "on_device", lambda t: True)
    monkeypatch.setattr(_lib, "stream_ptr", lambda queue=None: None)
    monkeypatch.setattr(_lib, "load", lambda strict=False: main)
    monkeypatch.setattr(codegen.JitLibrary, "compile", classmethod(lambda cls, ksrc: ksrc))
    monkeypatch.setattr(codegen.JitLibrary, "get", classmethod(lambda cls: jit))
    monkeypatch.setattr(torch.cuda, "get_device_properties",
                        lambda dev=None: types.SimpleNamespace(multi_processor_count=NUM_SMS))
    yield types.SimpleNamespace(jit=jit, main=main)
