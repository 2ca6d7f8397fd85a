"""Development timing of fused-stage kernel variants (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pystella_b200 import _kernels as K
from pystella_b200._lib import ENERGY_STRIDE


def time_case(N, h, dtype, nscalars=1, xchunk=0, variant=0, reps=10, energy=False, A=-0.4):
    dev = "cuda"
    shape = (nscalars,) + tuple(n + 2 * h for n in N)
    td = torch.float64 if dtype == np.float64 else torch.float32
    tf = torch.rand(shape, dtype=td, device=dev); tfo = torch.zeros_like(tf)
    tdd = torch.rand(shape, dtype=td, device=dev); tkf = torch.zeros_like(tf); tkd = torch.zeros_like(tf)
    dx = (0.01,) * 3
    kw = {}
    if energy:
        kw = dict(energy_partials=torch.zeros((65536, ENERGY_STRIDE), dtype=torch.float64, device=dev),
                  V_terms=[(0.5, (2,) + (0,) * (nscalars - 1))])
    def go():
        K.scalar_stage(fused=True, f_in=tf, f_out=tfo, dfdt=tdd, kf=tkf, kdfdt=tkd, h=h,
                       dx=dx, dt=1e-3, A=A, B=0.3, xchunk=xchunk, variant=variant,
                       dV_terms=[[(1.0, (1,) + (0,) * (nscalars - 1))]] + [[]] * (nscalars - 1), **kw)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sites = np.prod(N)
    S = np.dtype(dtype).itemsize
    gb = sites * (8 if A != 0 else 6) * S * nscalars / 1e9
    print(f"N={N[0]} h={h} {np.dtype(dtype).name} n={nscalars} xc={xchunk} var={variant} en={int(energy)} A={A}: "
          f"{ms:.3f} ms  {gb/ms*1e3:.0f} GB/s  {gb/ms*1e3/6485.8*100:.1f}%  {sites/ms/1e3:.0f} MLUPS", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    N5 = (512,) * 3
    if which in ("all", "ws"):
        for var in (1, 10, 11, 12, 13, 14):
            for en in (False, True):
                time_case(N5, 2, np.float64, variant=var, energy=en)
        for xc in (16, 32, 128, 256):
            time_case(N5, 2, np.float64, variant=10, xchunk=xc, energy=True)
        time_case(N5, 2, np.float64, variant=10, energy=True, A=0.0)
        time_case((256,) * 3, 2, np.float64, variant=10, energy=True)
        time_case(N5, 2, np.float32, variant=0, energy=True)
        time_case((256,) * 3, 2, np.float64, nscalars=2, variant=0)
        time_case((256,) * 3, 2, np.float64, nscalars=2, variant=1)
