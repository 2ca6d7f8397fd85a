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


class EmuFusedHandle:
    """A generated instance of the fused pipeline (pystella_b200/fused_jit.py): the functor
    source that NVRTC would get, and the kernel instance to call."""

    def __init__(self, source, name_expr):
        self.source, self.name_expr = source, name_expr


class EmuJitLibrary:
    """Stands in for libpystella_b200_jit.so: handles are KernelSource objects (generated
    maps / histograms / stencils) or EmuFusedHandle objects (generated fused stages)."""

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

    def psb_jit_launch_params_v(self, handle, params, gx, gy, gz, bx, by, bz, smem, flags,
                                stream):
        """The generated instance of the warp-specialised pipeline: same parameter list as on
        the GPU (FusedParams, four tensor maps, user scalars)."""
        import re
        import emu_library
        m = re.match(r"psb::fused_stage_ws_kernel<(\w+), (\d+), (\d+),", handle.name_expr)
        T, NF = m.group(1), m.group(3)
        driver = (
            "#define __CUDACC_RTC__ 1\n" + handle.source +
            'extern "C" void run(void** params, int gx, int gy, int gz, int threads, '
            "long long smem) {\n"
            f"  const psb::FusedParams<{T}, {NF}>& p = "
            f"*static_cast<const psb::FusedParams<{T}, {NF}>*>(params[0]);\n"
            "  const CUtensorMap* m[4];\n"
            "  for (int i = 0; i < 4; ++i) m[i] = static_cast<const CUtensorMap*>(params[1 + i]);\n"
            "  const psb::PsbJitScalars& js = *static_cast<const psb::PsbJitScalars*>(params[5]);\n"
            "  emu::launch_smem(dim3(gx, gy, gz), dim3(threads, 1, 1), (size_t)smem, [&] {\n"
            f"    {handle.name_expr}(p, *m[0], *m[1], *m[2], *m[3], js); }});\n}}\n")
        here = os.path.dirname(os.path.abspath(__file__))
        lib = build_emu(driver, "emu_fused_jit",
                        first=["-include", os.path.join(here, "cuda_emu", "cuda_emu.h"),
                               "-I" + emu_library.rtc_include_dir(),
                               "-I" + os.path.join(here, "cuda_emu", "stubs"),
                               "-ffp-contract=off"])
        self.launches.append((handle.name_expr, (gx, gy, gz), (bx, by, bz), smem))
        lib.run(params, C.c_int(int(gx)), C.c_int(int(gy)), C.c_int(int(gz)), C.c_int(int(bx)),
                C.c_longlong(int(smem)))
        return 0

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
    from pystella_b200 import fused_jit
    monkeypatch.setattr(fused_jit.GeneratedFusedStage, "compile",
                        lambda self, plan, strict=False: EmuFusedHandle(
                            self.source, self.name_expression(plan)))
    monkeypatch.setattr(codegen.JitLibrary, "compile", classmethod(lambda cls, ksrc: ksrc))
    monkeypatch.setattr(codegen.JitLibrary, "get", classmethod(lambda cls: jit))
    monkeypatch.setattr(torch.cuda, "get_device_properties",
                        lambda dev=None: types.SimpleNamespace(multi_processor_count=NUM_SMS))
    yield types.SimpleNamespace(jit=jit, main=main)
