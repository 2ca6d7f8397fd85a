import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pystella_b200 import _kernels as K
N=(16,16,128); h=2
shape=(1,)+tuple(n+2*h for n in N)
for dt_ in (torch.float64, torch.float32):
    tf=torch.rand(shape,dtype=dt_,device="cuda"); tfo=torch.zeros_like(tf); d=torch.rand_like(tf); kf=torch.zeros_like(tf); kd=torch.zeros_like(tf)
    K.scalar_stage(fused=True,f_in=tf,f_out=tfo,dfdt=d,kf=kf,kdfdt=kd,h=h,dx=(0.1,)*3,dt=1e-3,A=-0.4,B=0.3, variant=int(os.environ.get("VAR","0")))
    torch.cuda.synchronize(); print(dt_, "ok", float(tfo.sum()))
