"""Quick device-resident timing of the hot path on one GPU (C5 shard at half width); prints per-call kernel ms.
    python profiles/quick_gpu.py [B] [mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from successiveconvexification_b200 import dynamics, sample_problems as sp, workloads
prob = sp.base_prob_aero_scaled('tests/golden/aero_lift_drag.npz')
cache = dynamics.make_cache(prob); ctx = cache.sim_prob
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
K = 50
X, U, s, P = workloads.monte_carlo_batch(prob, K, B, 1003)
dX, dU, dS = (torch.from_numpy(a).cuda() for a in (X, U, s))
out = torch.empty((B, K, 23, 14), dtype=torch.float64, device='cuda')
err = torch.empty((B, K, 14), dtype=torch.float64, device='cuda')
tlb = torch.empty((B, K + 1, 4), dtype=torch.float64, device='cuda')
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for it in range(6):
    ctx.linearize_ptr(dX.data_ptr(), dU.data_ptr(), dS.data_ptr(), 1 / 51, 10, mode, K + 1, B, out.data_ptr(), err.data_ptr(), tlb.data_ptr())
    torch.cuda.synchronize()
    ms = ctx.last_kernel_ms()
    print('ms', round(ms, 4), 'intervals/s %.4e' % (B * K / ms * 1e3))
print('checksum %.17g' % float(out.abs().sum().item()), 'finite', bool(torch.isfinite(out).all().item()))
