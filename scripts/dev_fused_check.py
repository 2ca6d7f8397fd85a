"""Development check: fused stage vs oracle + timing (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pystella_b200 import _kernels as K
import oracle as O


def run_case(N, h, dtype, nscalars=1, strict=False, dV=None, pot=None, variant=0):
    rng = np.random.default_rng(1234)
    shape = (nscalars,) + tuple(n + 2 * h for n in N)
    f = rng.random(shape).astype(dtype)
    d = rng.random(shape).astype(dtype)
    kf = rng.random(shape).astype(dtype)
    kd = rng.random(shape).astype(dtype)
    dx = (0.1, 0.13, 0.07)
    dt, A, B = 0.01, -0.4178904745, 0.3792103129999
    a, hub = 1.3, 0.2
    for i in range(nscalars):
        O.share_halos(f[i], h)
    fo, do, kfo, kdo = f.copy(), d.copy(), kf.copy(), kd.copy()
    lap = np.zeros((nscalars,) + tuple(N), dtype)
    O.reference_schedule_stage(1, fo, do, kfo, kdo, lap, h, dx, dt, a=a, hubble=hub,
                               potential_terms=pot, tables="LowStorageRK54")
    # override A,B by calling lsrk with table coefficient: stage 1 of RK54
    A, B = O.LSRK_TABLES["LowStorageRK54"][0][1], O.LSRK_TABLES["LowStorageRK54"][1][1]
    dev = "cuda"
    tf = torch.from_numpy(f).to(dev); tfo = torch.zeros_like(tf)
    td = torch.from_numpy(d).to(dev); tkf = torch.from_numpy(kf).to(dev)
    tkd = torch.from_numpy(kd).to(dev)
    K.scalar_stage(fused=True, f_in=tf, f_out=tfo, dfdt=td, kf=tkf, kdfdt=tkd, h=h, dx=dx,
                   dt=dt, A=A, B=B, a=a, hubble=hub, dV_terms=dV, strict=strict,
                   variant=variant)
    torch.cuda.synchronize()
    res = {}
    for name, got, ref in (("f", tfo, fo), ("dfdt", td, do), ("kf", tkf, kfo), ("kdfdt", tkd, kdo)):
        g = got.cpu().numpy()
        gi = np.stack([O.interior(g[i], h) for i in range(nscalars)])
        ri = np.stack([O.interior(ref[i], h) for i in range(nscalars)])
        res[name] = (np.max(np.abs(gi - ri)) / np.max(np.abs(ri)), O.get_errs(gi, ri)[0],
                     bool((gi == ri).all()))
    # halo check: f_out faces
    g = tfo.cpu().numpy()
    ref_h = fo.copy()
    for i in range(nscalars):
        O.share_halos(ref_h[i], h)
    # faces only
    ok = True
    for i in range(nscalars):
        sl = slice(h, -h)
        ok &= np.array_equal(g[i][:h, sl, sl], ref_h[i][:h, sl, sl]) or strict is False and np.allclose(g[i][:h, sl, sl], ref_h[i][:h, sl, sl], rtol=1e-12)
        ok &= np.allclose(g[i][-h:, sl, sl], ref_h[i][-h:, sl, sl], rtol=1e-12)
        ok &= np.allclose(g[i][sl, :h, sl], ref_h[i][sl, :h, sl], rtol=1e-12)
        ok &= np.allclose(g[i][sl, -h:, sl], ref_h[i][sl, -h:, sl], rtol=1e-12)
        ok &= np.allclose(g[i][sl, sl, :h], ref_h[i][sl, sl, :h], rtol=1e-12)
        ok &= np.allclose(g[i][sl, sl, -h:], ref_h[i][sl, sl, -h:], rtol=1e-12)
    print(f"N={N} h={h} {np.dtype(dtype).name} n={nscalars} strict={strict} var={variant}: "
          + " ".join(f"{k}:{v[0]:.2e}/{v[1]:.2e}/{'EXACT' if v[2] else 'close'}" for k, v in res.items())
          + f" halos_ok={ok}")


def time_case(N, h, dtype, nscalars=1, xchunk=0, variant=0, reps=20):
    dev = "cuda"
    shape = (nscalars,) + tuple(n + 2 * h for n in N)
    td = torch.float64 if dtype == np.float64 else torch.float32
    tf = torch.rand(shape, dtype=td, device=dev); tfo = torch.zeros_like(tf)
    tdd = torch.rand(shape, dtype=td, device=dev); tkf = torch.zeros_like(tf); tkd = torch.zeros_like(tf)
    dx = (0.01,) * 3
    def go(A):
        K.scalar_stage(fused=True, f_in=tf, f_out=tfo, dfdt=tdd, kf=tkf, kdfdt=tkd, h=h,
                       dx=dx, dt=1e-3, A=A, B=0.3, xchunk=xchunk, variant=variant)
    for _ in range(3):
        go(-0.4)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go(-0.4)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sites = np.prod(N)
    S = np.dtype(dtype).itemsize
    gb = sites * 8 * S * nscalars / 1e9
    print(f"time N={N} h={h} {np.dtype(dtype).name} n={nscalars} xchunk={xchunk} var={variant}: "
          f"{ms:.3f} ms/stage  {gb/ms*1e3:.0f} GB/s  {sites/ms/1e3:.0f} MLUPS")


if __name__ == "__main__":
    pot = [(0.5, (2,))]          # V = f^2/2
    dV = [[(1.0, (1,))]]
    for strict in (False, True):
        run_case((32, 32, 32), 2, np.float64, strict=strict, dV=dV, pot=pot)
        run_case((20, 24, 36), 2, np.float64, strict=strict, dV=dV, pot=pot)
        run_case((20, 24, 36), 2, np.float64, strict=strict, dV=dV, pot=pot, variant=2)
        run_case((17, 9, 33), 1, np.float64, strict=strict)
        run_case((16, 16, 32), 1, np.float64, strict=strict)
        run_case((16, 12, 70), 3, np.float64, strict=strict)
        run_case((16, 12, 70), 4, np.float64, strict=strict)
        run_case((32, 32, 32), 2, np.float32, strict=strict, dV=dV, pot=pot)
        run_case((17, 9, 33), 1, np.float32, strict=strict)
        pot2 = [(0.5, (2, 0)), (0.5, (0, 2)), (0.5, (2, 2))]
        dV2 = [O.poly_diff(pot2, 0), O.poly_diff(pot2, 1)]
        run_case((24, 20, 64), 2, np.float64, nscalars=2, strict=strict, dV=dV2, pot=pot2)
    for N in ((256,) * 3, (512,) * 3):
        for xc in (32, 64, 128):
            time_case(N, 2, np.float64, xchunk=xc)
    time_case((512,) * 3, 2, np.float64, variant=2)
    time_case((512,) * 3, 2, np.float64, variant=3)
    time_case((512,) * 3, 2, np.float32)
    time_case((256,) * 3, 2, np.float64, nscalars=2)
