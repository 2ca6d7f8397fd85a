"""
Test infrastructure: a HOST build of the stand-alone translation units of
pystella_b200/csrc (halo, element-wise stages, reductions, background, diagnostics) on the
CPU execution model of tests/cuda_emu/cuda_emu.h, exporting the same C ABI
(include/pystella_b200.h).  The sources are compiled as they are; the only transformation is
the kernel-launch syntax, `k<<<grid, block, smem, stream>>>(args)` ->
`emu::launch_smem(grid, block, smem, [&] { k(args); })`.  The TMA pipelines (fused stages,
fd_ws) are hardware-specific and are not part of this build.

Nothing here is reachable from the product.
"""
import ctypes as C
import hashlib
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pystella_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "cuda_emu")

SOURCES = ["halo.cu", "rk_stage.cu", "stage_unfused.cu", "tensor_stage.cu", "reduce.cu",
           "expansion.cu", "diag.cu", "fd.cu"]

#: headers replaced as a whole: the TMA pipeline of the finite-difference kernels reports
#: "layout not eligible", so that psb_finite_difference takes its generic tile kernel (the
#: path odd radii and unaligned layouts take on the GPU)
SHADOWED = {
    "fd_ws.cuh": """#pragma once
namespace psb {
template <typename T, int H, int MODE>
static int launch_fd_ws(const FDParams<T>&, cudaStream_t) { return -1; }
}  // namespace psb
""",
}

def _split_top_level(text):
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "(<[{":
            depth += 1
        elif ch in ")>]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def _kernel_start(src, at):
    """Start index of the kernel name (identifier + optional template arguments) that ends
    just before position `at` (the `<<<`)."""
    i = at
    while src[i - 1].isspace():
        i -= 1
    if src[i - 1] == ">":                       # template argument list: balance backwards
        depth = 0
        while True:
            i -= 1
            depth += {">": 1, "<": -1}.get(src[i], 0)
            if depth == 0:
                break
        while src[i - 1].isspace():
            i -= 1
    while src[i - 1].isalnum() or src[i - 1] in "_:":
        i -= 1
    return i


def transform_launches(src):
    """`k<<<g, b, s, st>>>(args);` -> emu::launch_smem(dim3(g), dim3(b), s, [&] { k(args); });"""
    out, pos = "", 0
    while True:
        at = src.find("<<<", pos)
        if at < 0:
            break
        close = src.index(">>>", at)
        start = _kernel_start(src, at)
        kernel = src[start:at].strip()
        cfg = _split_top_level(src[at + 3:close])
        i = close + 3
        while src[i].isspace():
            i += 1
        assert src[i] == "(", src[start:i + 20]
        depth, j = 1, i + 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        args = src[i + 1:j - 1]
        assert src[j:].lstrip().startswith(";"), src[start:j + 20]
        end = src.index(";", j) + 1
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += src[pos:start]
        out += (f"emu::launch_smem(dim3({cfg[0]}), dim3({cfg[1]}), (size_t)({smem}), "
                f"[&] {{ {kernel}({args}); }});")
        pos = end
    return out + src[pos:]


_ASM = [
    # programmatic dependent launch: no-ops without the launch attribute
    (re.compile(r'asm volatile\("griddepcontrol\.(?:launch_dependents|wait);"\s*:::\s*"memory"\);'),
     ";"),
    # system-scope release store / acquire load of the inter-GPU flags
    (re.compile(r'asm volatile\("st\.release\.sys\.global\.u64 \[%0\], %1;"\s*::\s*"l"\((\w+)\),'
                r'\s*"l"\((\w+)\)\s*:\s*"memory"\);'),
     r"__atomic_store_n(\1, \2, __ATOMIC_RELEASE);"),
    (re.compile(r'asm volatile\("ld\.acquire\.sys\.global\.u64 %0, \[%1\];"\s*:\s*"=l"\((\w+)\)'
                r'\s*:\s*"l"\((\w+)\)\s*:\s*"memory"\);'),
     r"\1 = __atomic_load_n(\2, __ATOMIC_ACQUIRE);"),
]


def transform_asm(src):
    """The handful of PTX statements on the emulated paths -> their C++ meaning."""
    for rx, repl in _ASM:
        src = rx.sub(repl, src)
    return src


_COMMON = r'''
// stands in for common.cu under emulation: error channel and ABI constants
#include <stdarg.h>
#include "common.cuh"
namespace psb {
static thread_local char g_error[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_error, sizeof(g_error), fmt, ap); va_end(ap);
}
int encode_field_tensor_map(CUtensorMap*, int, const void*, int64_t, int64_t, int64_t, int64_t,
                            int64_t, int, int, int) { return 1; }   // no TMA here
}
extern "C" {
const char* psb_last_error(void) { return psb::g_error; }
int psb_abi_version(void) { return PSB_ABI_VERSION; }
int psb_strict_fp(void) { return 1; }
}
'''

_lib = None


def build(verbose=False):
    """-> ctypes library with the emulated entry points (built once per source state)."""
    global _lib
    if _lib is not None:
        return _lib
    texts = {}
    for name in SOURCES:
        with open(os.path.join(CSRC, name)) as fh:
            texts[name] = transform_asm(transform_launches(fh.read()))
    texts["emu_common.cu"] = _COMMON
    headers = {}
    for hdr in sorted(os.listdir(CSRC)):
        if hdr.endswith(".cuh"):
            with open(os.path.join(CSRC, hdr)) as fh:
                headers[hdr] = SHADOWED.get(hdr) or transform_asm(fh.read())
    h = hashlib.sha256()
    for name in sorted(texts):
        h.update(texts[name].encode())
    for hdr in sorted(headers):
        h.update(headers[hdr].encode())
    for hdr in ("cuda_emu.h", "stubs/cuda.h", "stubs/cuda_runtime.h"):
        h.update(open(os.path.join(EMU, hdr), "rb").read())
    d = os.path.join(tempfile.gettempdir(), "psb_emu_lib_" + h.hexdigest()[:16])
    so = os.path.join(d, "libpystella_b200_emu.so")
    if not os.path.exists(so):
        os.makedirs(d, exist_ok=True)
        for hdr, text in headers.items():       # transformed copies shadow csrc/*.cuh
            with open(os.path.join(d, hdr), "w") as fh:
                fh.write(text)
        objs = []
        for name, text in texts.items():
            cpp = os.path.join(d, name[:-3] + ".cpp")
            with open(cpp, "w") as fh:
                fh.write(text)
            obj = cpp[:-4] + ".o"
            cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-pthread", "-fPIC", "-c",
                   "-Wno-unknown-pragmas", "-Wno-attributes", "-DPSB_STRICT_FP=1",
                   "-include", os.path.join(EMU, "cuda_emu.h"),
                   "-I" + d, "-I" + os.path.join(EMU, "stubs"), "-I" + EMU, "-I" + CSRC,
                   "-I" + os.path.join(ROOT, "include"), cpp, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"emulated build of {name} failed:\n" + res.stderr[-4000:])
            objs.append(obj)
        tmp = so + f".{os.getpid()}.tmp"
        subprocess.check_call(["g++", "-shared", "-pthread", "-o", tmp] + objs)
        os.replace(tmp, so)
    _lib = C.CDLL(so)
    return _lib
