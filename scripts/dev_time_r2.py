"""Round-2 development timings (run on the GPU box): fused scalar stage with the in-kernel
tail, store-policy and x-chunk variants, the fused tensor stage (config 4)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pystella_b200 as ps

PEAK = 6485.8


def timeit(fn, name, nbytes, reps=6, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = nbytes / ms / 1e6
    print(f"{name:66s} {ms:8.3f} ms {gbs:7.0f} GB/s {gbs / PEAK:6.3f}", flush=True)
    return ms


def scalar_case(shape, dtype, stage, energy, nscalars=1, tune=None, attach=True, label=""):
    h = 2
    S = torch.finfo(dtype).bits // 8
    dx = (5 / shape[1],) * 3
    pad = (nscalars,) + tuple(n + 2 * h for n in shape)
    if nscalars == 1:
        pot = lambda g: g[0]**2 / 2  # noqa: E731
    else:
        pot = lambda g: g[0]**2 / 2 + 0.3 * g[0]**2 * g[1]**2 / 2  # noqa: E731
    sector = ps.ScalarSector(nscalars, potential=pot)
    st = ps.LowStorageRK54(sector, halo_shape=h, dt=1e-3, dx=dx,
                           MapKernel=ps.FusedStencilStage, energy=energy, tune=tune)
    f = torch.rand(pad, dtype=dtype, device="cuda")
    d = torch.rand_like(f)
    sites = float(np.prod(shape))
    if energy:
        ex = ps.DeviceExpansion(1.0, ps.LowStorageRK54, grid_size=sites, nscalars=nscalars)
        if attach:
            ex.attach(st)
        kw = dict(a=ex.a, hubble=ex.hubble)
    else:
        kw = dict(a=1.0, hubble=0.1)

    def go():
        st(stage, f=f, dfdt=d, **kw)
        if energy:
            ex.step(stage, st, 1e-3)
    nb = (6 if stage == 0 else 8) * S * sites * nscalars
    name = (f"scalar {'x'.join(map(str, shape))} {str(dtype)[6:]} n={nscalars} stage={stage} "
            f"en={int(energy)} att={int(attach)} {tune or ''} {label}")
    return timeit(go, name, nb)


def gw_case(N, dtype, tune=None, stage=1):
    h = 2
    S = torch.finfo(dtype).bits // 8
    grid = (N,) * 3
    dx = (5 / N,) * 3
    pad = tuple(n + 2 * h for n in grid)
    pot = lambda g: (1.2e-6**2 / 2 * g[0]**2 + 2.5e-7 / 2 * g[0]**2 * g[1]**2) / 1.2e-6**2  # noqa
    scalar = ps.ScalarSector(2, potential=pot)
    gw = ps.TensorPerturbationSector([scalar])
    mk = lambda *lead: torch.rand(lead + pad, dtype=dtype, device="cuda") * 1e-3  # noqa
    arrays = dict(f=mk(2), dfdt=mk(2), hij=mk(6), dhijdt=mk(6))
    sites = float(N**3)
    st0 = ps.LowStorageRK54(scalar, halo_shape=h, dt=1e-3, dx=dx, MapKernel=ps.FusedStencilStage,
                            tune=tune)
    t_s = timeit(lambda: st0(stage, f=arrays["f"], dfdt=arrays["dfdt"], a=1.0, hubble=0.1),
                 f"c4 scalar part {N}^3 {str(dtype)[6:]} stage={stage} {tune or ''}",
                 (12 if stage == 0 else 16) * S * sites)
    st = ps.LowStorageRK54([scalar, gw], halo_shape=h, dt=1e-3, dx=dx,
                           MapKernel=ps.FusedStencilStage, tune=tune)
    t_all = timeit(lambda: st(stage, a=1.0, hubble=0.1, **arrays),
                   f"c4 scalar+GW stage {N}^3 {str(dtype)[6:]} stage={stage} {tune or ''}",
                   (48 if stage == 0 else 64) * S * sites)
    nb = (36 if stage == 0 else 48) * S * sites
    t = t_all - t_s
    print(f"   -> tensor kernel alone ~ {t:.3f} ms  {nb / t / 1e6:.0f} GB/s "
          f"{nb / t / 1e6 / PEAK:.3f} (48 S/site, sources from L2)", flush=True)
    del st, st0, arrays
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["scalar", "chunk", "store", "gw"]
    F64, F32 = torch.float64, torch.float32
    N5 = (512,) * 3
    if "gwvar" in which:
        for var in (0, 9):
            gw_case(512, F32, tune={"gw_variant": var})
            gw_case(256, F64, tune={"gw_variant": var})
        gw_case(512, F32, tune={"gw_variant": 9, "gw_xchunk": 32})
        for var in (0, 22):
            scalar_case(N5, F32, 1, True, nscalars=2, tune={"variant": var})
            scalar_case((256,) * 3, F64, 1, True, nscalars=2, tune={"variant": var})
    if "pws" in which:
        for var in (0, 30):
            for shape in (N5, (64, 512, 512), (256,) * 3):
                scalar_case(shape, F64, 1, True, tune={"variant": var})
            scalar_case(N5, F64, 0, True, tune={"variant": var})
            scalar_case((128, 1024, 1024), F32, 1, True, tune={"variant": var})
            scalar_case(N5, F32, 1, True, nscalars=2, tune={"variant": var})
    if "jit2" in which:
        # two-field scalar stage (config 4's): polynomial tables vs the generated functor
        from pystella_b200.fused_jit import recognise_general_system
        h = 2
        for dtype in (F32, F64):
            N = 512 if dtype == F32 else 320
            dx = (5 / N,) * 3
            pad = (2,) + (N + 2 * h,) * 3
            pot = lambda g: g[0]**2 / 2 + 86805.5 * g[0]**2 * g[1]**2  # noqa: E731
            for forced in (False, True):
                sector = ps.ScalarSector(2, potential=pot)
                st = ps.LowStorageRK54(sector, halo_shape=h, dt=1e-3, dx=dx,
                                       MapKernel=ps.FusedStencilStage, energy=True)
                if forced:
                    st._systems = [recognise_general_system(sector.rhs_dict,
                                                            sector.potential(sector.f))]
                f = torch.rand(pad, dtype=dtype, device="cuda") * 1e-2
                d = torch.rand_like(f)
                ex = ps.DeviceExpansion(1.0, ps.LowStorageRK54, grid_size=float(N**3), nscalars=2)
                ex.attach(st)
                S = torch.finfo(dtype).bits // 8
                for stage in (1, 0):
                    def go():
                        st(stage, f=f, dfdt=d, a=ex.a, hubble=ex.hubble)
                        ex.step(stage, st, 1e-3)
                    timeit(go, f"2 fields {N}^3 {str(dtype)[6:]} stage={stage} "
                               f"{'generated functor' if forced else 'polynomial tables'}",
                           (6 if stage == 0 else 8) * S * 2 * float(N**3))
    if "gwprof" in which:
        gw_case(384, F32)
    if "f32prof" in which:
        scalar_case(N5, F32, 1, True)
        scalar_case(N5, F64, 0, True)
    if "ab" in which:      # A/B of library builds (PSB_DEV_LIB): the cases that moved
        for dt_ in (F64, F32):
            for stage in (1, 0):
                scalar_case(N5, dt_, stage, False)
                scalar_case(N5, dt_, stage, True)
        scalar_case(N5, F32, 1, True, nscalars=2)
        scalar_case((256,) * 3, F64, 1, True)
    if "scalar" in which:
        for dt_ in (F64, F32):
            for stage in (1, 0):
                scalar_case(N5, dt_, stage, False)
                scalar_case(N5, dt_, stage, True, attach=False)
                scalar_case(N5, dt_, stage, True, attach=True)
        scalar_case((128, 1024, 1024), F32, 1, True)
        scalar_case((256,) * 3, F64, 1, True)
        scalar_case((256,) * 3, F64, 1, True, nscalars=2)
        scalar_case(N5, F32, 1, True, nscalars=2)
    if "chunk" in which:
        for xc in (-16, 16, 32, -32, 8):
            scalar_case((64, 512, 512), F64, 1, True, tune={"xchunk": xc})
        for xc in (-16, 16, 32):
            scalar_case((256,) * 3, F64, 1, True, tune={"xchunk": xc})
        for xc in (-16, 16, 32):
            scalar_case((128, 1024, 1024), F32, 1, True, tune={"xchunk": xc})
        for xc in (-16, 16, 32):
            scalar_case(N5, F64, 1, True, tune={"xchunk": xc})
    if "store" in which:
        for var in (10, 11, 12, 13, 14):
            scalar_case(N5, F64, 0, True, tune={"variant": var})
    if "gw" in which:
        gw_case(256, F32)
        gw_case(512, F32)
        gw_case(512, F32, tune={"gw_variant": 1})
        gw_case(512, F32, tune={"gw_xchunk": 32})
        gw_case(512, F32, tune={"gw_xchunk": 8})
        gw_case(512, F32, stage=0)
        gw_case(256, F64)
