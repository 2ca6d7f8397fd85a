"""
The wave equation d^2 f / dt^2 = lap f on a periodic box: the pystella_b200 port of
the reference's examples/wave_equation.py, line by line (pyopencl arrays ->
torch tensors, CommandQueue -> stream, same classes otherwise).

    python examples/wave_equation.py [--fused] [--grid 64] [--rk4]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pystella_b200 as ps  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--grid", type=int, default=32)
parser.add_argument("--fused", action="store_true",
                    help="MapKernel=FusedStencilStage (no separate derivs call)")
parser.add_argument("--rk4", action="store_true", help="classical RK4 (three copies)")
parser.add_argument("--end-time", type=float, default=1.)
args = parser.parse_args()

# set parameters
grid_shape = (args.grid,) * 3
proc_shape = (1, 1, 1)
rank_shape = tuple(Ni // pi for Ni, pi in zip(grid_shape, proc_shape))
halo_shape = 1
dtype = torch.float64
dx = tuple(10 / Ni for Ni in grid_shape)
dt = min(dx) / 10

# device and halo-sharer
device = ps.choose_device()
decomp = ps.DomainDecomposition(proc_shape, halo_shape, rank_shape)

# initialize arrays with random data
pad = tuple(ni + 2 * halo_shape for ni in rank_shape)
torch.manual_seed(12321)
copies = (3,) if args.rk4 else ()
f = torch.zeros(copies + pad, dtype=dtype, device=device)
dfdt = torch.zeros(copies + pad, dtype=dtype, device=device)
(f[0] if args.rk4 else f).copy_(torch.rand(pad, dtype=dtype, device=device))
(dfdt[0] if args.rk4 else dfdt).copy_(torch.rand(pad, dtype=dtype, device=device))
lap_f = torch.zeros(rank_shape, dtype=dtype, device=device)

# define system of equations
f_ = ps.DynamicField("f", offset="h")  # don't overwrite f
rhs_dict = {
    f_: f_.dot,  # df/dt = \dot{f}
    f_.dot: f_.lap  # d\dot{f}/dt = \nabla^2 f
}

# create time-stepping and derivative-computing kernels
derivs = ps.FiniteDifferencer(decomp, halo_shape, dx)
if args.rk4:
    stepper = ps.RungeKutta4(rhs_dict, dt=dt, halo_shape=halo_shape)
elif args.fused:
    stepper = ps.LowStorageRK54(rhs_dict, dt=dt, halo_shape=halo_shape,
                                MapKernel=ps.FusedStencilStage, dx=dx, decomp=decomp)
else:
    stepper = ps.LowStorageRK54(rhs_dict, dt=dt, halo_shape=halo_shape)


def energy():
    cur_f, cur_d = (f[0], dfdt[0]) if args.rk4 else (f, dfdt)
    derivs(None, fx=cur_f, lap=lap_f)
    s = (slice(halo_shape, -halo_shape),) * 3
    return float((cur_d[s]**2 / 2 - cur_f[s] * lap_f / 2).mean())


e0 = energy()
t = 0.
# loop over time
while t < args.end_time:
    for s in range(stepper.num_stages):
        if args.rk4:
            derivs(None, fx=f[0 if s == 0 else 1], lap=lap_f)
            stepper(s, queue=None, f=f, dfdt=dfdt, lap_f=lap_f)
        elif args.fused:
            stepper(s, queue=None, f=f, dfdt=dfdt)
        else:
            derivs(None, fx=f, lap=lap_f)
            stepper(s, queue=None, f=f, dfdt=dfdt, lap_f=lap_f)
    t += dt

torch.cuda.synchronize()
e1 = energy()
print(f"t = {t:.3f}: energy {e0:.12f} -> {e1:.12f} (relative drift {abs(e1 / e0 - 1):.2e})")
# white-noise initial data excite every mode up to Nyquist; the 4th-order scheme damps
# the fastest ones slightly, so "conserved" means a small drift, not round-off
assert abs(e1 / e0 - 1) < 2e-2, "energy of the wave equation should be (nearly) conserved"
