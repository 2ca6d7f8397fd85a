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

SOURCES = ["common.cu", "halo.cu", "rk_stage.cu", "stage_unfused.cu", "tensor_stage.cu",
           "reduce.cu", "expansion.cu", "diag.cu", "fd.cu",
           "fused_stage.cu", "fused_stage_f64.cu", "fused_stage_f32.cu",
           "tensor_stage_fused.cu"]
#: the fused-stage units instantiate every kernel variant (incl. the TMA pipelines, which
#: cannot run here: their PTX traps): compiled without optimisation to keep the build short
UNOPTIMISED = {"fused_stage_f64.cu", "fused_stage_f32.cu", "tensor_stage_fused.cu"}

SHADOWED = {}


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


_OPERAND = re.compile(r'"[=+]?[a-z]"\s*\(((?:[^()]|\([^()]*\))*)\)')


def _asm_replacement(stmt):
    """C++ meaning of one `asm volatile(...)` statement under the execution model."""
    ops = [re.sub(r"^smem_u32\((.*)\)$", r"\1", m.group(1).strip())
           for m in _OPERAND.finditer(stmt)]
    if "griddepcontrol" in stmt:
        return ";"                               # no-ops without the launch attribute
    if "st.release.sys.global.u64" in stmt:
        return f"__atomic_store_n({ops[0]}, {ops[1]}, __ATOMIC_RELEASE);"
    if "ld.acquire.sys.global.u64" in stmt:
        return f"{ops[0]} = __atomic_load_n({ops[1]}, __ATOMIC_ACQUIRE);"
    if "st.volatile.global.f64" in stmt:
        return f"*(volatile double*)({ops[0]}) = ({ops[1]});"
    if "ld.volatile.global.b64" in stmt:
        return f"{ops[0]} = *(volatile unsigned long long*)({ops[1]});"
    m = re.search(r"bar\.sync (\w+|%0), (%\d)", stmt)
    if m:
        ident = ops[0] if m.group(1) == "%0" else m.group(1)
        count = ops[1] if m.group(1) == "%0" else ops[0]
        return f"emu::named_barrier({ident}, {count});"
    # mbarrier / TMA (tests/cuda_emu/cuda_emu.h: the model)
    if "mbarrier.init" in stmt:
        return f"emu::mbar_init({ops[0]}, {ops[1]});"
    if "mbarrier.arrive.expect_tx" in stmt:
        return f"emu::mbar_expect_tx({ops[0]}, {ops[1]});"
    if "mbarrier.arrive" in stmt:
        return f"emu::mbar_arrive({ops[0]});"
    if "mbarrier.try_wait.parity" in stmt:
        return f"emu::mbar_wait({ops[0]}, {ops[1]});"
    if "cp.async.bulk.tensor.4d.shared" in stmt:
        return ("emu::tma_load_4d(" + ", ".join(ops[:7]) + ");")
    if ("fence.mbarrier_init" in stmt or "fence.proxy.async" in stmt
            or "cp.async.bulk.prefetch" in stmt or "prefetch.tensormap" in stmt):
        return ";"                               # ordering / cache hints
    return "emu::unsupported_ptx();"


_DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];")


def transform_asm(src):
    # dynamic shared memory: `extern __shared__ T name[];` -> the model's buffer
    src = _DYN_SMEM.sub(r"\1* \2 = reinterpret_cast<\1*>(emu::dynamic_smem());", src)
    out, pos = "", 0
    while True:
        at = src.find("asm volatile", pos)
        if at < 0:
            break
        i = src.index("(", at)
        depth, j, in_str = 1, i + 1, False
        while depth:
            ch = src[j]
            if ch == '"' and src[j - 1] != "\\":
                in_str = not in_str
            elif not in_str:
                depth += {"(": 1, ")": -1}.get(ch, 0)
            j += 1
        end = src.index(";", j) + 1
        out += src[pos:at] + _asm_replacement(src[at:end])
        pos = end
    return out + src[pos:]


_CONTROL = '''
// knobs of the execution model, exported next to the C ABI
extern "C" void emu_set_reverse_blocks(int v) { emu::reverse_blocks() = v != 0; }
extern "C" void emu_set_tma(int v) { emu::tma_enabled() = v != 0; }
'''

_lib = None
_dir = None


def build(verbose=False):
    """-> ctypes library with the emulated entry points (built once per source state)."""
    global _lib, _dir
    if _lib is not None:
        return _lib
    texts = {}
    for name in SOURCES:
        with open(os.path.join(CSRC, name)) as fh:
            texts[name] = transform_asm(transform_launches(fh.read()))
    texts["emu_control.cu"] = _CONTROL
    headers = {}
    for hdr in sorted(os.listdir(CSRC)):
        if hdr.endswith(".cuh"):
            with open(os.path.join(CSRC, hdr)) as fh:
                headers[hdr] = SHADOWED.get(hdr) or transform_asm(
                    transform_launches(fh.read()))
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
        def compile_one(item):
            name, text = item
            cpp = os.path.join(d, name[:-3] + ".cpp")
            with open(cpp, "w") as fh:
                fh.write(text)
            obj = cpp[:-4] + ".o"
            cmd = ["g++", "-std=c++17", "-O0" if name in UNOPTIMISED else "-O1",
                   "-ffp-contract=off", "-pthread", "-fPIC", "-c",
                   "-Wno-unknown-pragmas", "-Wno-attributes", "-DPSB_STRICT_FP=1",
                   "-include", os.path.join(EMU, "cuda_emu.h"),
                   "-I" + d, "-I" + os.path.join(EMU, "stubs"), "-I" + EMU, "-I" + CSRC,
                   "-I" + os.path.join(ROOT, "include"), cpp, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"emulated build of {name} failed:\n" + res.stderr[-4000:])
            return obj

        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            objs = list(ex.map(compile_one, texts.items()))
        tmp = so + f".{os.getpid()}.tmp"
        subprocess.check_call(["g++", "-shared", "-pthread", "-o", tmp] + objs)
        os.replace(tmp, so)
    # second copy of the headers for sources written for NVRTC (the generated instances of
    # the pipeline): they are compiled with __CUDACC_RTC__ defined, which drops the host code;
    # only the first guard of common.cuh (its own stand-ins for the system headers) must go
    rtc = os.path.join(d, "rtc")
    if not os.path.isdir(rtc):
        os.makedirs(rtc + ".tmp", exist_ok=True)
        for hdr, text in headers.items():
            if hdr == "common.cuh":
                text = text.replace("#ifndef __CUDACC_RTC__", "#if 1  /* emulated */", 1)
            if hdr == "fused_stage_impl.cuh":
                # the persistent variant has no generated instances, and g++ (unlike NVRTC)
                # looks names of its templates up before fused_stage_ws.cuh has defined them
                text = text.replace('#include "fused_stage_pws.cuh"', "")
            with open(os.path.join(rtc + ".tmp", hdr), "w") as fh:
                fh.write(text)
        try:
            os.replace(rtc + ".tmp", rtc)
        except OSError:
            pass                      # another process was faster
    _dir = d
    _lib = C.CDLL(so)
    return _lib


def rtc_include_dir():
    """Directory with the transformed headers for NVRTC-style sources (after build())."""
    build()
    return os.path.join(_dir, "rtc")
