"""Small fused-stage launches for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pystella_b200 import _kernels as K
from pystella_b200._lib import ENERGY_STRIDE
for dt_, N, h in ((torch.float64, (20, 12, 128), 2), (torch.float32, (20, 12, 128), 2),
                  (torch.float64, (10, 9, 33), 1)):
    shape = (1,) + tuple(n + 2 * h for n in N)
    f = torch.rand(shape, dtype=dt_, device="cuda"); fo = torch.zeros_like(f)
    d = torch.rand_like(f); kf = torch.zeros_like(f); kd = torch.zeros_like(f)
    ep = torch.zeros((4096, ENERGY_STRIDE), dtype=torch.float64, device="cuda")
    K.halo_wrap(f[0], (h,) * 3)
    K.scalar_stage(fused=True, f_in=f, f_out=fo, dfdt=d, kf=kf, kdfdt=kd, h=h, dx=(0.1,) * 3,
                   dt=1e-3, A=-0.4, B=0.3, dV_terms=[[(1.0, (1,))]], V_terms=[(0.5, (2,))],
                   energy_partials=ep)
    torch.cuda.synchronize()
    print(dt_, N, h, "ok", float(fo.sum()))
