import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, torch
from successiveconvexification_b200 import dynamics, sample_problems as sp, workloads
prob = sp.base_prob_aero_scaled('tests/golden/aero_lift_drag.npz')
cache = dynamics.make_cache(prob); ctx = cache.sim_prob
print('fp64 peak TF', ctx.measure_fp64_peak())
B,K=16384,50
X,U,s,P = workloads.monte_carlo_batch(prob,K,B,1003)
dX,dU,dS = (torch.from_numpy(a).cuda() for a in (X,U,s))
out = torch.empty((B,K,23,14),dtype=torch.float64,device='cuda')
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for kern in (2,):
    ctx.set_kernel(kern)
    for it in range(3):
        ctx.linearize_ptr(dX.data_ptr(),dU.data_ptr(),dS.data_ptr(),1/51,10,0,K+1,B,out.data_ptr())
        torch.cuda.synchronize()
        ms = ctx.last_kernel_ms()
        print('kernel',kern,'ms',ms,'intervals/s',B*K/ms*1e3)
