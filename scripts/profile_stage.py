"""Short driver for ncu: a few LowStorageRK54 steps of the bench workload (512^3 fp64).
Environment: PSB_N, PSB_STEPS, PSB_ENERGY, PSB_XCHUNK, PSB_VARIANT (16 = aligned mapping),
PSB_DTYPE (float64 / float32), PSB_POTENTIAL (poly / cos: the latter runs the pipeline with a
generated right-hand side)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pystella_b200 as ps

N = int(os.environ.get("PSB_N", "512"))
nsteps = int(os.environ.get("PSB_STEPS", "1"))
energy = os.environ.get("PSB_ENERGY", "1") == "1"
h = 2
dx = (5 / N,) * 3
dt = 0.1 * dx[0]
dtype = getattr(torch, os.environ.get("PSB_DTYPE", "float64"))
generated = os.environ.get("PSB_POTENTIAL", "poly") == "cos"
if generated:
    from pystella_b200.field import var
    sector = ps.ScalarSector(1, potential=lambda f: 1 - var("cos")(f[0]))
else:
    sector = ps.ScalarSector(1, potential=lambda f: f[0]**2 / 2)
tune = {"xchunk": int(os.environ.get("PSB_XCHUNK", "0")),
        "variant": int(os.environ.get("PSB_VARIANT", "0"))}
stepper = ps.LowStorageRK54(sector, halo_shape=h, dt=dt, dx=dx, MapKernel=ps.FusedStencilStage,
                            energy=energy, tune=tune, jit=generated or None)
pad = (1,) + (N + 2 * h,) * 3
torch.manual_seed(0)
f = 0.193 + 1e-6 * torch.randn(pad, dtype=dtype, device="cuda")
dfdt = -0.142 + 1e-6 * torch.randn(pad, dtype=dtype, device="cuda")
expand = ps.DeviceExpansion(0.5 * 0.142**2 + 0.5 * 0.193**2, ps.LowStorageRK54, grid_size=N**3)
for _ in range(nsteps):
    for s in range(5):
        stepper(s, f=f, dfdt=dfdt, a=expand.a, hubble=expand.hubble)
        if energy:
            expand.step(s, stepper, dt)
torch.cuda.synchronize()
print("done", float(expand.a[0]))
