"""
Block-level kernels EXECUTED without a GPU: the diagnostics kernels (row f-4 of the scope
table) use shared memory, barriers, shared/global atomics and warp shuffles, which the
per-site source transformation of test_codegen.py cannot run.  tests/cuda_emu/cuda_emu.h is a
small CUDA-on-CPU execution model (one OS thread per CUDA thread of a block, blocks in
sequence); the device code is compiled unchanged with g++ on top of it and checked against
NumPy.  The same kernels are checked on the GPU in tests/test_zz_gpu_diagnostics.py.
"""

import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

from pystella_b200 import codegen
from pystella_b200.field import Field, var

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "cuda_emu")
CSRC = os.path.join(ROOT, "pystella_b200", "csrc")
_cache = {}

# a kernel that hangs on the execution model (a protocol bug) must fail the run, not stall it
pytestmark = pytest.mark.timeout(600, method="thread")


def build_emu(source_text, tag, extra=(), first=()):
    """g++-compile `source_text` on top of cuda_emu.h into a shared object (cached)."""
    key = hashlib.sha256(source_text.encode()).hexdigest()[:16]
    if key not in _cache:
        d = tempfile.mkdtemp(prefix="psb_emu_")
        cpp = os.path.join(d, f"{tag}.cpp")
        so = os.path.join(d, f"{tag}.so")
        with open(cpp, "w") as fh:
            fh.write(source_text)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-shared", "-fPIC",
                               "-Wno-unknown-pragmas", "-Wno-attributes"] + list(first)
                              + ["-I" + EMU, "-I" + CSRC,
                               "-I" + os.path.join(ROOT, "include")] + list(extra)
                              + [cpp, "-o", so])
        _cache[key] = C.CDLL(so)
    return _cache[key]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------
# psb_bin_power (csrc/diag_kernels.cuh)
# ---------------------------------------------------------------------------

def _spectrum_case(N, is_real, dtype, k_power, seed=1, bw_scale=1.0):
    rng = np.random.default_rng(seed)
    rd = np.float64 if dtype == "f64" else np.float32
    cd = np.complex128 if dtype == "f64" else np.complex64
    kx = np.fft.fftfreq(N[0], 1 / N[0]).astype(rd)
    ky = np.fft.fftfreq(N[1], 1 / N[1]).astype(rd)
    kz = (np.fft.rfftfreq(N[2], 1 / N[2]) if is_real
          else np.fft.fftfreq(N[2], 1 / N[2])).astype(rd)
    K = (len(kx), len(ky), len(kz))
    fk = (rng.standard_normal(K) + 1j * rng.standard_normal(K)).astype(cd)
    dk = np.array([0.3, 0.41, 0.37])
    bw = dk.min() * bw_scale
    kv = np.meshgrid(kx, ky, kz, indexing="ij")
    kmag = np.sqrt(sum((d * k.astype(np.float64))**2 for d, k in zip(dk, kv)))
    nb = int(kmag.max() / bw + .5) + 1
    b = np.floor(kmag / bw + .5).astype(int)
    mult = (np.where((kv[2] > 0) & (kv[2] < N[2] / 2), 2., 1.) if is_real
            else np.ones_like(kmag))
    w = mult * kmag**k_power * (fk.real.astype(np.float64)**2 + fk.imag.astype(np.float64)**2)
    ref = np.bincount(b.ravel(), weights=w.ravel(), minlength=nb)[:nb]
    return dict(kx=kx, ky=ky, kz=kz, K=K, fk=fk, dk=dk, bw=bw, nb=nb, ref=ref,
                nyq=N[2] / 2)


@pytest.mark.parametrize("N, is_real, dtype, k_power, bw_scale, blocks", [
    ((16, 12, 20), 1, "f64", 3, 1.0, 5),         # r2c, per-warp copies
    ((16, 12, 20), 0, "f64", 3, 1.0, 7),         # c2c (|k| not monotone along z)
    ((8, 20, 70), 1, "f32", 2.5, 1.0, 3),        # complex64, pow(), Kz = 36 (one tail chunk)
    ((8, 20, 140), 0, "f64", 0, 0.001, 4),       # ~10^4 bins: global atomics only, Kz > 64
    ((4, 9, 66), 1, "f64", 1, 0.05, 100),        # 2 copies; more CTAs than row groups
    ((4, 9, 66), 1, "f64", 1, 1.0, 0),           # automatic grid (4 CTAs per "SM")
])
def test_bin_power_kernel_on_host(N, is_real, dtype, k_power, bw_scale, blocks):
    """csrc/diag_kernels.cuh (kernel + launch set-up) behind the psb_bin_power struct."""
    from pystella_b200 import _lib
    src = open(os.path.join(EMU, "emu_diag.cpp")).read()
    lib = build_emu(src, "emu_diag")
    c = _spectrum_case(N, is_real, dtype, k_power, bw_scale=bw_scale)
    out = np.zeros(c["nb"])
    a = _lib.BinPowerArgs()
    a.dtype = _lib.PSB_F64 if dtype == "f64" else _lib.PSB_F32
    a.is_real = is_real
    a.Kx, a.Ky, a.Kz = c["K"]
    a.fk, a.momenta_x, a.momenta_y, a.momenta_z = (x.ctypes.data for x in
                                                   (c["fk"], c["kx"], c["ky"], c["kz"]))
    for mu in range(3):
        a.dk[mu] = c["dk"][mu]
    a.bin_width, a.k_power, a.nyquist = c["bw"], k_power, c["nyq"]
    a.num_bins, a.tune_blocks = c["nb"], blocks
    a.spectrum = out.ctypes.data
    assert lib.emu_psb_bin_power(C.byref(a), C.c_int(2)) == 0
    assert np.max(np.abs(out - c["ref"])) <= 1e-13 * np.max(np.abs(c["ref"]))


# ---------------------------------------------------------------------------
# generated Histogrammer kernels (codegen.generate_histogram_kernel)
# ---------------------------------------------------------------------------

def run_generated_histogram_on_host(ksrc, arrays, halos, scalars, shape, blocks):
    driver = (
        '#include "cuda_emu.h"\n' + ksrc.source +
        'extern "C" void run(const psb_jit_args* a, int blocks) {\n'
        f'  emu::launch(dim3(blocks), dim3(32, {codegen.HIST_WARPS}), '
        f'[&] {{ {ksrc.name}(*a); }});\n}}\n')
    lib = build_emu(driver, "emu_hist")
    args = codegen.JitArgs()
    for n, (name, sub, _) in enumerate(ksrc.slots):
        view = arrays[name][sub] if sub else arrays[name]
        assert view.flags["C_CONTIGUOUS"]
        args.arrays[n] = view.ctypes.data
        args.halos[n] = halos[name]
    hist = np.zeros((ksrc.num_hists, ksrc.num_bins))
    args.arrays[ksrc.hist_slot] = hist.ctypes.data
    for n, s in enumerate(ksrc.scalars):
        args.scalars[n] = scalars[s]
    args.Nx, args.Ny, args.Nz = shape
    lib.run(C.byref(args), C.c_int(blocks))
    return hist


@pytest.mark.parametrize("N, h, dtype, nb, blocks", [
    ((6, 9, 40), 2, "float64", 12, 4),           # tails in y (9 % 8) and z (40 % 32)
    ((3, 16, 150), 1, "float32", 40, 50),        # two unrolled z iterations, fp32 bins
    ((5, 8, 32), 0, "float64", 3500, 3),         # 2 x 3500 bins > 48 KB: global atomics only
])
def test_generated_field_histograms_on_host(N, h, dtype, nb, blocks):
    """The two histograms of FieldHistogrammer (pystella/histogram.py:210-290): linear and
    logarithmic bins with clipping, in one generated kernel."""
    rng = np.random.default_rng(5)
    pad = tuple(n + 2 * h for n in N)
    f_arr = (rng.standard_normal(pad) * 3).astype(dtype)
    inner = f_arr[h:h + N[0], h:h + N[1], h:h + N[2]] if h else f_arr
    f = Field("f", offset=h)
    vmin, vmax, log, fabs = var("min"), var("max"), var("log"), var("fabs")

    def clip(e):
        return vmax(vmin(e, nb - 1), 0)

    lin = (f - var("min_f")) / (var("max_f") - var("min_f"))
    lg = (log(fabs(f)) - var("min_log_f")) / (var("max_log_f") - var("min_log_f"))
    ctype = "double" if dtype == "float64" else "float"
    ksrc = codegen.generate_histogram_kernel([(clip(lin * nb), 1), (clip(lg * nb), f * f)], nb,
                                             ctype, halo_shape=h)
    use_shared, copies = codegen.histogram_layout(2, nb)
    assert use_shared == (2 * nb * 8 <= 48 * 1024)
    T = np.dtype(dtype).type
    sc = dict(min_f=float(inner.min()), max_f=float(inner.max()),
              min_log_f=float(np.log(np.abs(inner)).min()),
              max_log_f=float(np.log(np.abs(inner)).max()))
    got = run_generated_histogram_on_host(ksrc, {"f": f_arr}, {"f": h}, sc, N, blocks)
    # NumPy in the kernel's precision
    x = inner.astype(dtype)
    s = {k: T(v) for k, v in sc.items()}
    b0 = np.floor(np.maximum(np.minimum((x - s["min_f"]) / (s["max_f"] - s["min_f"]) * T(nb),
                                        T(nb - 1)), T(0))).astype(int)
    b1 = np.floor(np.maximum(np.minimum(
        (np.log(np.abs(x)) - s["min_log_f"]) / (s["max_log_f"] - s["min_log_f"]) * T(nb),
        T(nb - 1)), T(0))).astype(int)
    ref0 = np.bincount(b0.ravel(), minlength=nb).astype(float)
    ref1 = np.bincount(b1.ravel(), weights=(x * x).astype(np.float64).ravel(), minlength=nb)
    assert got[0].sum() == np.prod(N)
    # a value within one ulp of a bin edge may land in the neighbouring bin (log/divide
    # rounding of libm vs NumPy): compare cumulative counts with a tolerance of 2 sites
    assert np.max(np.abs(np.cumsum(got[0]) - np.cumsum(ref0))) <= 2
    assert np.allclose(np.cumsum(got[1]), np.cumsum(ref1), rtol=1e-12,
                       atol=2 * float((x * x).max()))


def test_generated_histogram_drops_out_of_range_bins_and_nan():
    """Bins outside [0, num_bins) are dropped (the reference's kernel would write out of
    bounds, histogram.py:118-126); NaN bins as well."""
    N, nb = (2, 3, 37), 5
    rng = np.random.default_rng(0)
    x = rng.uniform(-3, 8, N)
    x[0, 1, 5] = np.nan
    f = Field("f", offset=0)
    ksrc = codegen.generate_histogram_kernel([(f, 2)], nb, "double", halo_shape=0)
    got = run_generated_histogram_on_host(ksrc, {"f": x}, {"f": 0}, {}, N, 2)
    ok = np.isfinite(x) & (np.floor(x) >= 0) & (np.floor(x) < nb)
    ref = np.bincount(np.floor(x[ok]).astype(int), minlength=nb) * 2.
    assert np.array_equal(got[0], ref)


def test_warp_aggregation_handles_repeated_runs():
    """Equal bins that are NOT adjacent in a warp (a, b, a) must both be counted: the run
    heads then collide on one shared-memory address and rely on the atomic."""
    N, nb = (1, 1, 64), 4
    x = np.tile(np.array([0., 1., 0., 0., 2., 2., 1., 3.]), 8).reshape(N)
    f = Field("f", offset=0)
    ksrc = codegen.generate_histogram_kernel([(f, f + 1)], nb, "double", halo_shape=0)
    got = run_generated_histogram_on_host(ksrc, {"f": x}, {"f": 0}, {}, N, 1)
    ref = np.bincount(x.astype(int).ravel(), weights=(x + 1).ravel(), minlength=nb)
    assert np.array_equal(got[0], ref)


@pytest.mark.parametrize("ctype, nb", [("double", 12), ("float", 432), ("double", 7000)])
def test_generated_histogram_compiles_for_sm_100a(ctype, nb, monkeypatch):
    """NVRTC accepts the generated histogram kernels (no GPU needed): shared-memory copies
    for small bin counts, global atomics beyond 48 KB."""
    monkeypatch.setenv("PSB_JIT_CACHE", "off")
    f = Field("f", offset="h")
    lin = (f - var("min_f")) / (var("max_f") - var("min_f")) * nb
    ksrc = codegen.generate_histogram_kernel(
        [(lin, 1), (var("log")(var("fabs")(f)), f * f)], nb, ctype, halo_shape=2)
    handle = codegen.JitLibrary.compile(ksrc)
    assert handle > 0 and codegen.JitLibrary.get().psb_jit_cubin_size(handle) > 0
    assert ("__shared__ double temp" in ksrc.source) == (2 * nb * 8 <= 48 * 1024)


# ---------------------------------------------------------------------------
# stencil maps with shared-memory tiles (codegen.generate_stencil_kernel)
# ---------------------------------------------------------------------------

def run_generated_stencil_on_host(ksrc, arrays, halos, scalars, shape, gz):
    driver = (
        '#include "cuda_emu.h"\n' + ksrc.source +
        'extern "C" void run(const psb_jit_args* a, int gx, int gy, int gz, long long smem) {\n'
        f'  emu::launch_smem(dim3(gx, gy, gz), dim3({codegen.TILE_TZ}, {codegen.TILE_TY}), '
        f'(size_t)smem, [&] {{ {ksrc.name}(*a); }});\n}}\n')
    lib = build_emu(driver, "emu_stencil")
    args = codegen.JitArgs()
    for n, (name, sub, _) in enumerate(ksrc.slots):
        view = arrays[name][sub] if sub else arrays[name]
        assert view.flags["C_CONTIGUOUS"]
        args.arrays[n] = view.ctypes.data
        args.halos[n] = halos[name]
    for n, s in enumerate(ksrc.scalars):
        args.scalars[n] = scalars[s]
    args.Nx, args.Ny, args.Nz = shape
    codegen.TiledStencilMap.check_reach(ksrc.prefetch, [halos[name] for name, _, _ in ksrc.slots])
    gx = (shape[2] + codegen.TILE_TZ - 1) // codegen.TILE_TZ
    tyt = codegen.TILE_TY * codegen.TILE_RY
    gy = (shape[1] + tyt - 1) // tyt
    lib.run(C.byref(args), C.c_int(gx), C.c_int(gy), C.c_int(gz),
            C.c_longlong(ksrc.smem_bytes))


@pytest.mark.parametrize("h, dtype, N, gz", [
    (1, "float64", (6, 5, 7), 1),          # the reference's test: a single short tile
    (2, "float64", (9, 11, 70), 2),        # tails in y (11 % 8) and z (70 % 64), 2 x chunks
    (2, "float32", (7, 16, 64), 3),        # exact tiles, uneven x chunks (3 + 3 + 1)
    (1, "float64", (3, 8, 130), 5),        # more x chunks than planes
])
def test_tiled_stencil_raw_subscripts_on_host(h, dtype, N, gz):
    """test/test_stencil.py:36-90: the six-point sum, `x` staged in shared memory."""
    rng = np.random.default_rng(9)
    x, y = var("x"), var("y")
    i, j, k = var("i"), var("j"), var("k")
    map_dict = {y[i, j, k]: (x[i + h + h, j + h, k + h] + x[i + h, j + h + h, k + h]
                             + x[i + h, j + h, k + h + h] + x[i - h + h, j + h, k + h]
                             + x[i + h, j - h + h, k + h] + x[i + h, j + h, k - h + h])}
    ctype = "double" if dtype == "float64" else "float"
    ksrc = codegen.generate_stencil_kernel(map_dict, None, ctype, halo_shape=h,
                                           prefetch_args=["x"])
    assert ksrc.prefetch == [dict(name="x", slot=0, absolute=True, lo=(0, 0, 0),
                                  hi=(2 * h, 2 * h, 2 * h))]
    x_h = rng.random(tuple(n + 2 * h for n in N)).astype(dtype)
    y_h = rng.random(N).astype(dtype)
    run_generated_stencil_on_host(ksrc, {"x": x_h, "y": y_h}, {"x": h, "y": 0}, {}, N, gz)
    y_true = (x_h[2*h:, h:-h, h:-h] + x_h[h:-h, 2*h:, h:-h] + x_h[h:-h, h:-h, 2*h:]
              + x_h[:-2*h, h:-h, h:-h] + x_h[h:-h, :-2*h, h:-h] + x_h[h:-h, h:-h, :-2*h])
    tol = 5e-14 if dtype == "float64" else 1e-5
    assert np.max(np.abs(y_h - y_true)) <= tol * np.max(np.abs(y_true))


def test_tiled_stencil_fields_temporaries_and_mixed_operands_on_host():
    """Field operands with asymmetric offsets, two staged arrays with different padding, a
    directly read array, a temporary and a scalar; a second statement that sees the first."""
    from pystella_b200.field import shift_fields
    rng = np.random.default_rng(3)
    N, hf, hg = (5, 10, 66), 2, 1
    f = Field("f", offset=hf)
    g = Field("g", offset=hg)
    u = Field("u", offset=0)
    out1, out2 = Field("out1", offset=0), Field("out2", offset=1)

    def sh(fld, s):
        return shift_fields(fld, s)

    tmp = {var("c"): var("alpha") * 2}
    maps = [(out1, sh(f, (2, 0, 0)) - sh(f, (-1, 0, 0)) + sh(f, (0, -2, 1)) * var("c")
             + sh(g, (0, 1, -1)) * u),
            (out2, out1 + sh(g, (1, 0, 0)) + sh(f, (0, 0, 0)))]
    ksrc = codegen.generate_stencil_kernel(maps, tmp, "double", halo_shape=2,
                                           prefetch_args=["f", "g"])
    byname = {p["name"]: p for p in ksrc.prefetch}
    assert byname["f"]["lo"] == (-1, -2, 0) and byname["f"]["hi"] == (2, 0, 1)
    assert byname["g"]["lo"] == (0, 0, -1) and byname["g"]["hi"] == (1, 1, 0)
    arr = {"f": rng.random(tuple(n + 2 * hf for n in N)),
           "g": rng.random(tuple(n + 2 * hg for n in N)),
           "u": rng.random(N), "out1": np.zeros(N),
           "out2": np.zeros(tuple(n + 2 for n in N))}
    halos = {"f": hf, "g": hg, "u": 0, "out1": 0, "out2": 1}
    alpha = 0.7
    run_generated_stencil_on_host(ksrc, arr, halos, {"alpha": alpha}, N, 2)

    def view(a, h, s):
        return a[tuple(slice(h + si, h + si + n) for si, n in zip(s, N))]
    fa, ga = arr["f"], arr["g"]
    o1 = (view(fa, hf, (2, 0, 0)) - view(fa, hf, (-1, 0, 0))
          + view(fa, hf, (0, -2, 1)) * (alpha * 2) + view(ga, hg, (0, 1, -1)) * arr["u"])
    o2 = o1 + view(ga, hg, (1, 0, 0)) + view(fa, hf, (0, 0, 0))
    assert np.max(np.abs(arr["out1"] - o1)) < 1e-14
    assert np.max(np.abs(view(arr["out2"], 1, (0, 0, 0)) - o2)) < 1e-14
    assert arr["out2"][0].max() == 0 and arr["out2"][:, :, 0].max() == 0   # halo untouched


def test_tiled_stencil_rejects_offsets_outside_the_padding():
    x, y = var("x"), var("y")
    i, j, k = var("i"), var("j"), var("k")
    ksrc = codegen.generate_stencil_kernel({y[i, j, k]: x[i + 4, j + 2, k + 2]}, None,
                                           "double", halo_shape=2, prefetch_args=["x"])
    with pytest.raises(ValueError, match="outside its padding"):
        codegen.TiledStencilMap.check_reach(ksrc.prefetch, [1, 0])
    codegen.TiledStencilMap.check_reach(ksrc.prefetch, [2, 0])


@pytest.mark.parametrize("ctype", ["double", "float"])
def test_tiled_stencil_compiles_for_sm_100a(ctype, monkeypatch):
    monkeypatch.setenv("PSB_JIT_CACHE", "off")
    h = 2
    x, y = var("x"), var("y")
    i, j, k = var("i"), var("j"), var("k")
    map_dict = {y[i, j, k]: (x[i + h + h, j + h, k + h] + x[i + h, j + h + h, k + h]
                             + x[i + h, j + h, k + h + h] + x[i - h + h, j + h, k + h]
                             + x[i + h, j - h + h, k + h] + x[i + h, j + h, k - h + h])}
    ksrc = codegen.generate_stencil_kernel(map_dict, None, ctype, halo_shape=h,
                                           prefetch_args=["x"])
    assert codegen.JitLibrary.compile(ksrc) > 0
    assert codegen.TiledStencilMap.launch_grid((512, 512, 512), 148) == (8, 64, 8)
    assert codegen.TiledStencilMap.launch_grid((16, 12, 40), 148) == (1, 2, 2)


@pytest.mark.parametrize("seed", [0, 3, 4, 7])
def test_tiled_stencil_random_cases_on_host(seed):
    """Randomised sweep: 1-3 input arrays with their own paddings, 2-5 random offsets each
    (anywhere inside the padding, asymmetric bounding boxes), a random subset staged in shared
    memory, random lattice extents (tails in y and z, short x), random number of x chunks;
    two outputs, the second reading the first."""
    from pystella_b200.field import shift_fields
    rng = np.random.default_rng(1000 + seed)
    dtype = "float64" if seed % 3 else "float32"
    ctype = "double" if dtype == "float64" else "float"
    N = (int(rng.integers(1, 9)), int(rng.integers(1, 20)), int(rng.integers(1, 140)))
    narr = int(rng.integers(1, 4))
    names = ["p", "q", "r"][:narr]
    halos, fields, arrays, shifts = {}, {}, {}, {}
    for nm in names:
        H = int(rng.integers(1, 4))
        halos[nm] = H
        fields[nm] = Field(nm, offset=H)
        arrays[nm] = rng.random(tuple(n + 2 * H for n in N)).astype(dtype)
        shifts[nm] = [tuple(int(v) for v in rng.integers(-H, H + 1, 3))
                      for _ in range(int(rng.integers(2, 6)))]
    coef = {nm: [float(c) for c in rng.integers(1, 5, len(shifts[nm]))] for nm in names}
    out1, out2 = Field("out1", offset=0), Field("out2", offset=0)
    expr = 0
    for nm in names:
        for c, s in zip(coef[nm], shifts[nm]):
            expr = expr + c * shift_fields(fields[nm], s)
    maps = [(out1, expr * var("scale")), (out2, out1 - fields[names[0]])]
    staged = [nm for nm in names if rng.random() < 0.7] or [names[0]]
    ksrc = codegen.generate_stencil_kernel(maps, None, ctype, halo_shape=max(halos.values()),
                                           prefetch_args=staged)
    arrays.update(out1=np.zeros(N, dtype), out2=np.zeros(N, dtype))
    halos.update(out1=0, out2=0)
    scale = 0.5
    gz = int(rng.integers(1, 4))
    run_generated_stencil_on_host(ksrc, arrays, halos, {"scale": scale}, N, gz)

    def view(a, H, s):
        return a[tuple(slice(H + si, H + si + n) for si, n in zip(s, N))].astype(np.float64)
    ref1 = sum(c * view(arrays[nm], halos[nm], s)
               for nm in names for c, s in zip(coef[nm], shifts[nm])) * scale
    ref2 = ref1 - view(arrays[names[0]], halos[names[0]], (0, 0, 0))
    tol = 1e-13 if dtype == "float64" else 2e-5
    scale1 = max(1.0, np.max(np.abs(ref1)))
    assert np.max(np.abs(arrays["out1"] - ref1)) <= tol * scale1, (N, staged, shifts, gz)
    assert np.max(np.abs(arrays["out2"] - ref2)) <= tol * scale1, (N, staged, shifts, gz)


@pytest.mark.parametrize("seed", [0, 1, 3])
def test_generated_histogram_random_cases_on_host(seed):
    """Randomised sweep: 1-3 histograms over 1-2 arrays with different paddings, random
    lattice extents, bin counts on both sides of the shared-memory limit, weights that are
    constants or expressions; bins partly out of range."""
    rng = np.random.default_rng(2000 + seed)
    dtype = "float64" if seed % 2 else "float32"
    ctype = "double" if dtype == "float64" else "float"
    T = np.dtype(dtype).type
    N = (int(rng.integers(1, 7)), int(rng.integers(1, 20)), int(rng.integers(1, 150)))
    nb = int(rng.choice([3, 17, 100, 2500]))
    Hf, Hg = int(rng.integers(0, 3)), int(rng.integers(0, 3))
    f, g = Field("f", offset=Hf), Field("g", offset=Hg)
    fa = rng.random(tuple(n + 2 * Hf for n in N)).astype(dtype)
    ga = rng.random(tuple(n + 2 * Hg for n in N)).astype(dtype)
    nh = int(rng.integers(1, 4))
    s = T(nb * 1.3)
    hists = [(f * var("s") - 2, 1), (g * var("s"), f + g), (f * g * var("s"), 2.5)][:nh]
    ksrc = codegen.generate_histogram_kernel(hists, nb, ctype, halo_shape=None)
    arrays = {"f": fa, "g": ga}
    halos = {"f": Hf, "g": Hg}
    got = run_generated_histogram_on_host(ksrc, arrays, halos, {"s": float(s)}, N,
                                          int(rng.integers(1, 6)))

    def inner(a, H):
        return a[tuple(slice(H, H + n) for n in N)]
    fi, gi = inner(fa, Hf), inner(ga, Hg)
    bins = [np.floor(fi * s - T(2)), np.floor(gi * s), np.floor(fi * gi * s)][:nh]
    weights = [np.ones(N), (fi + gi).astype(np.float64), np.full(N, 2.5)][:nh]
    for j in range(nh):
        ok = (bins[j] >= 0) & (bins[j] < nb)
        ref = np.bincount(bins[j][ok].astype(int), weights=weights[j][ok], minlength=nb)
        assert np.allclose(got[j], ref, rtol=1e-12, atol=1e-12), (N, nb, j)
