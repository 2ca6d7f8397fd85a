"""Development timing of the stand-alone API kernels at 512^3 (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pystella_b200 as ps
N = (512,) * 3; h = 2
dx = (5 / 512,) * 3
decomp = ps.DomainDecomposition((1, 1, 1), h, N)
derivs = ps.FiniteDifferencer(decomp, h, dx, rank_shape=N)
sector = ps.ScalarSector(1, potential=lambda f: f[0]**2 / 2)
red = ps.Reduction(decomp, sector, halo_shape=h, callback=ps.get_rho_and_p, grid_size=512**3)
pad = (1,) + (516,) * 3
f = torch.rand(pad, dtype=torch.float64, device="cuda"); d = torch.rand_like(f)
lap = torch.zeros((1,) + N, dtype=torch.float64, device="cuda")
grd = torch.zeros((1, 3) + N, dtype=torch.float64, device="cuda")
def timeit(fn, name, bytes_):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:28s} {ms:7.3f} ms  {bytes_/ms/1e6:7.0f} GB/s", flush=True)
S = 512**3 * 8
timeit(lambda: decomp.share_halos(None, f[0]), "share_halos", 0.03 * S)
timeit(lambda: derivs(None, fx=f, lap=lap), "derivs lap (incl. halos)", 2 * S)
timeit(lambda: derivs(None, fx=f, lap=lap, grd=grd), "derivs lap+grd", 5 * S)
timeit(lambda: red(None, f=f, dfdt=d, lap_f=lap, a=np.array(1.0)), "Reduction energy (+sync)", 3 * S)
st = ps.LowStorageRK54(sector, halo_shape=h, dt=1e-3)
timeit(lambda: st(1, f=f, dfdt=d, lap_f=lap, a=1.0, hubble=0.1), "unfused stage", 10 * S)
