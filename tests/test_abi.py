"""The C-ABI library loads and exports every symbol include/pystella_b200.h declares."""
import ctypes
import os
import re

import pytest

from pystella_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "pystella_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psb_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("strict", [False, True])
def test_library_exports_header_symbols(strict):
    names = header_functions()
    assert len(names) >= 12
    lib = ctypes.CDLL(_lib.lib_path(strict))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SYMBOLS) == names


def test_loader_and_error_channel():
    lib = _lib.load()
    assert lib.psb_abi_version() == 1
    assert lib.psb_strict_fp() == 0
    assert _lib.load(strict=True).psb_strict_fp() == 1
    # argument validation happens before any CUDA call: safe without a GPU
    rc = lib.psb_scalar_stage_fused(None, None)
    assert rc != 0 and b"null args" in lib.psb_last_error()
    args = _lib.ScalarStageArgs()
    rc = lib.psb_scalar_stage_fused(ctypes.byref(args), None)
    assert rc != 0 and b"null array" in lib.psb_last_error()
    rc = lib.psb_halo_wrap(None, 0, 4, 4, 4, 1, 1, 1, 7, None)
    assert rc != 0
    assert lib.psb_reduce_workspace_doubles(4) > 0


def test_struct_layout_matches_header():
    """sizeof of the ctypes mirror must equal the C struct (checked via a tiny C probe)."""
    import subprocess
    import tempfile
    src = ('#include <stdio.h>\n#include "pystella_b200.h"\nint main(){printf("%zu %zu\\n",'
           'sizeof(psb_scalar_stage_args), sizeof(psb_reduce_args));'
           'printf("%zu %zu %zu %zu\\n", sizeof(psb_tensor_stage_args), sizeof(psb_peer_args),'
           'sizeof(psb_rk_stage_args), sizeof(psb_tensor_fused_args));'
           'return 0;}')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        a, b, c, d, e, f = subprocess.check_output([exe]).split()
    assert int(a) == ctypes.sizeof(_lib.ScalarStageArgs)
    assert int(b) == ctypes.sizeof(_lib.ReduceArgs)
    assert int(c) == ctypes.sizeof(_lib.TensorStageArgs)
    assert int(d) == ctypes.sizeof(_lib.PeerArgs)
    assert int(e) == ctypes.sizeof(_lib.RKStageArgs)
    assert int(f) == ctypes.sizeof(_lib.TensorFusedArgs)


def test_no_cpu_fallback():
    import torch
    import pystella_b200 as ps
    from pystella_b200 import _kernels
    t = torch.zeros(6, 6, 6, dtype=torch.float64)
    with pytest.raises(ps.PsbError):
        _kernels.halo_wrap(t, (1, 1, 1))
    with pytest.raises(ps.PsbError):
        _kernels.finite_difference(t, lap=torch.zeros(4, 4, 4, dtype=torch.float64), h=1,
                                   dx=(1, 1, 1))
    d = ps.DomainDecomposition((1, 1, 1), 1)
    with pytest.raises(ps.PsbError):
        d.share_halos(None, t)
