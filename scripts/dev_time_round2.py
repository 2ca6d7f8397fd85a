"""Timings that decide round 2's tuning order (run on the GPU box): the experimental fused
tensor stage against the unfused schedule, and the run-time generated stage against the
hand-written element-wise stage."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pystella_b200 as ps


def timeit(fn, name, nbytes, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:44s} {ms:8.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s", flush=True)
    return ms


def scalar_stage(N=512, dtype=torch.float64):
    h, S = 2, torch.finfo(dtype).bits // 8
    grid = (N,) * 3
    dx = (5 / N,) * 3
    pad = (1,) + tuple(n + 2 * h for n in grid)
    f = torch.rand(pad, dtype=dtype, device="cuda")
    d = torch.rand_like(f)
    lap = torch.zeros((1,) + grid, dtype=dtype, device="cuda")
    sector = ps.ScalarSector(1, potential=lambda g: g[0]**2 / 2)
    sites = N**3
    hand = ps.LowStorageRK54(sector, halo_shape=h, dt=1e-3)
    timeit(lambda: hand(1, f=f, dfdt=d, lap_f=lap, a=1.0, hubble=0.1),
           f"hand-written element-wise stage {N}^3", 9 * S * sites)
    gen = ps.LowStorageRK54(sector, halo_shape=h, dt=1e-3, jit=True)
    gen._systems = None
    timeit(lambda: gen._jit_stage(1, None, dict(f=f, dfdt=d, lap_f=lap, a=1.0, hubble=0.1)),
           f"generated element-wise stage {N}^3", 9 * S * sites)


def fused_generated(N=512, dtype=torch.float64):
    """The fused pipeline with the built-in polynomial RHS vs the same potential through the
    generated functor (forced), and a non-polynomial potential."""
    from pystella_b200.field import var
    from pystella_b200.fused_jit import recognise_general_system
    h, S = 2, torch.finfo(dtype).bits // 8
    grid = (N,) * 3
    dx = (5 / N,) * 3
    pad = (1,) + tuple(n + 2 * h for n in grid)
    sites = N**3
    cases = [("built-in phi^2/2", lambda g: g[0]**2 / 2, False),
             ("generated phi^2/2", lambda g: g[0]**2 / 2, True),
             ("generated 1-cos(phi)", lambda g: 1 - var("cos")(g[0]), True)]
    for name, pot, force in cases:
        sector = ps.ScalarSector(1, potential=pot)
        st = ps.LowStorageRK54(sector, halo_shape=h, dt=1e-3, dx=dx,
                               MapKernel=ps.FusedStencilStage, jit=True)
        if force:
            st._systems = [recognise_general_system(sector.rhs_dict, sector.potential(sector.f))]
        f = torch.rand(pad, dtype=dtype, device="cuda")
        d = torch.rand_like(f)
        try:
            timeit(lambda: st(1, f=f, dfdt=d, a=1.0, hubble=0.1),
                   f"fused stage {N}^3 {dtype}: {name}", 8 * S * sites)
        except Exception as exc:  # noqa: BLE001
            print(f"{name}: {type(exc).__name__}: {str(exc)[:300]}", flush=True)


def aligned_variant(N=512, dtype=torch.float32, h=2):
    """Default pipeline vs the aligned-mapping variant (tune_variant 16)."""
    S = torch.finfo(dtype).bits // 8
    grid = (N,) * 3
    dx = (5 / N,) * 3
    pad = (1,) + tuple(n + 2 * h for n in grid)
    sector = ps.ScalarSector(1, potential=lambda g: g[0]**2 / 2)
    for variant in (0, 16):
        st = ps.LowStorageRK54(sector, halo_shape=h, dt=1e-3, dx=dx,
                               MapKernel=ps.FusedStencilStage, tune={"variant": variant})
        f = torch.rand(pad, dtype=dtype, device="cuda")
        d = torch.rand_like(f)
        try:
            timeit(lambda: st(1, f=f, dfdt=d, a=1.0, hubble=0.1),
                   f"fused stage {N}^3 {dtype} h={h} variant {variant}", 8 * S * N**3)
        except Exception as exc:  # noqa: BLE001
            print(f"variant {variant}: {type(exc).__name__}: {str(exc)[:300]}", flush=True)


def tensor_stage(N=256, dtype=torch.float32):
    h, S = 2, torch.finfo(dtype).bits // 8
    grid = (N,) * 3
    dx = (5 / N,) * 3
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid)
    derivs = ps.FiniteDifferencer(decomp, h, dx, rank_shape=grid)
    scalar = ps.ScalarSector(2, potential=lambda g: g[0]**2 / 2 + g[0]**2 * g[1]**2 / 2)
    gw = ps.TensorPerturbationSector([scalar])
    pad = tuple(n + 2 * h for n in grid)
    mk = lambda *lead: torch.rand(lead + pad, dtype=dtype, device="cuda") * 1e-3  # noqa
    arrays = dict(f=mk(2), dfdt=mk(2), hij=mk(6), dhijdt=mk(6),
                  lap_f=torch.zeros((2,) + grid, dtype=dtype, device="cuda"),
                  dfdx=torch.zeros((2, 3) + grid, dtype=dtype, device="cuda"),
                  lap_hij=torch.zeros((6,) + grid, dtype=dtype, device="cuda"))
    sites = N**3
    for fused in ("0", "1"):
        os.environ["PSB_FUSED_GW"] = fused
        st = ps.LowStorageRK54([scalar, gw], halo_shape=h, dt=1e-3, dx=dx, decomp=decomp,
                               MapKernel=ps.FusedStencilStage)

        def stage():
            if fused == "0":
                derivs(None, fx=arrays["f"], lap=arrays["lap_f"], grd=arrays["dfdx"])
                derivs(None, fx=arrays["hij"], lap=arrays["lap_hij"])
            st(1, a=1.0, hubble=0.1, filter_args=True, **arrays)
        try:
            timeit(stage, f"scalar+GW stage {N}^3 {dtype} PSB_FUSED_GW={fused}",
                   64 * S * sites)
        except Exception as exc:  # noqa: BLE001
            print(f"PSB_FUSED_GW={fused}: {type(exc).__name__}: {exc}", flush=True)


if __name__ == "__main__":
    aligned_variant(512, torch.float32, 2)
    aligned_variant(512, torch.float64, 2)
    aligned_variant(510, torch.float64, 1)      # odd radius: default = cooperative loader
    fused_generated(512, torch.float64)
    fused_generated(512, torch.float32)
    scalar_stage(512, torch.float64)
    scalar_stage(512, torch.float32)
    tensor_stage(256, torch.float32)
    tensor_stage(256, torch.float64)
