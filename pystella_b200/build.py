"""
In-tree build of the CUDA extension (sm_100a only).

    python -m pystella_b200.build [--force] [--strict]

Compiles every ``csrc/*.cu`` with nvcc and links ``_C/libpystella_b200.so``
(and, with ``--strict``/by default, ``_C/libpystella_b200_strict.so`` built with
``-fmad=false`` so that fp64 results are bit-comparable with the NumPy oracle).
The shared objects are git-ignored but travel to the GPU box with the snapshot.
"""

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_C")
OBJ = os.path.join(ROOT, "build", "obj")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
          "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]

SOURCES = ["common.cu", "fused_stage.cu", "fused_stage_f64.cu", "fused_stage_f32.cu",
           "stage_unfused.cu", "tensor_stage.cu", "tensor_stage_fused.cu", "rk_stage.cu", "fd.cu", "halo.cu", "reduce.cu", "expansion.cu",
           "diag.cu"]


def _digest(paths, extra):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(repr(extra).encode())
    return h.hexdigest()


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(cmd) + "\n" + res.stdout
                           + res.stderr)
    return res.stderr


def build_variant(name, extra_flags, force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    objdir = os.path.join(OBJ, name)
    os.makedirs(objdir, exist_ok=True)
    lib = os.path.join(OUT, name + ".so")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(ROOT, "include", "pystella_b200.h"))
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    stamp = os.path.join(objdir, "digest")
    digest = _digest(srcs + headers, (ARCH, COMMON, extra_flags))
    if (not force and os.path.exists(lib) and os.path.exists(stamp)
            and open(stamp).read() == digest):
        return lib

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + ARCH + COMMON + extra_flags + ["-Xptxas", "-v", "-c", src,
                                                     "-o", obj]
        log = _run(cmd)
        with open(obj + ".ptxas.log", "w") as fh:
            fh.write(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    _run([NVCC] + ARCH + ["-shared", "-o", lib] + objs + ["-lcudart_static", "-ldl",
                                                           "-lrt", "-lpthread"])
    with open(stamp, "w") as fh:
        fh.write(digest)
    if verbose:
        print("built", lib)
    return lib


def build_jit(force=False, verbose=False):
    """NVRTC + driver-API runtime of the experimental generated kernels (host code only;
    separate so that the main library carries no NVRTC dependency)."""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(CSRC, "jit.cpp")
    hdr = os.path.join(ROOT, "include", "pystella_b200_jit.h")
    lib = os.path.join(OUT, "libpystella_b200_jit.so")
    stamp = lib + ".digest"
    digest = _digest([src, hdr], "jit-v1")
    if (not force and os.path.exists(lib) and os.path.exists(stamp)
            and open(stamp).read() == digest):
        return lib
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "lib64")
    _run([NVCC, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"),
          "-shared", "-o", lib, src, "-lnvrtc", "-lcudart_static", "-ldl", "-lrt", "-lpthread",
          "-Xlinker", "-rpath", "-Xlinker", cuda_lib])
    with open(stamp, "w") as fh:
        fh.write(digest)
    if verbose:
        print("built", lib)
    return lib


def build(force=False, strict=True, verbose=False):
    libs = [build_variant("libpystella_b200", [], force, verbose)]
    if strict:
        libs.append(build_variant("libpystella_b200_strict",
                                  ["-fmad=false", "-DPSB_STRICT_FP=1"], force, verbose))
    libs.append(build_jit(force, verbose))
    return libs


if __name__ == "__main__":
    build(force="--force" in sys.argv, strict="--no-strict" not in sys.argv,
          verbose=True)
