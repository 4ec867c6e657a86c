"""e2e sweep over the host-path chunk size (SCVX_HOST_CHUNK_MB): pinned host buffers through scvx_linearize_batch."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    from successiveconvexification_b200 import dynamics, sample_problems as sp, workloads
    prob = sp.base_prob_aero_scaled('tests/golden/aero_lift_drag.npz')
    cache = dynamics.make_cache(prob); ctx = cache.sim_prob
    B, K = 32768, 50
    X, U, s, P = workloads.monte_carlo_batch(prob, K, B, 1003)
    hX, hU, hS = (torch.from_numpy(a).pin_memory() for a in (X, U, s))
    hO = torch.empty((B, K, 23, 14), dtype=torch.float64).pin_memory()
    hE = torch.empty((B, K, 14), dtype=torch.float64).pin_memory()
    hT = torch.empty((B, K + 1, 4), dtype=torch.float64).pin_memory()
    def step():
        ctx.linearize_ptr(hX.data_ptr(), hU.data_ptr(), hS.data_ptr(), 1 / 51, 10, 0, K + 1, B, hO.data_ptr(), hE.data_ptr(), hT.data_ptr())
    step(); step()
    t0 = time.perf_counter()
    for _ in range(4): step()
    dt = (time.perf_counter() - t0) / 4
    print(os.environ.get("SCVX_HOST_CHUNK_MB"), "MB chunks:", round(B * K / dt / 1e6, 2), "M intervals/s", round((hO.nbytes + hE.nbytes + hT.nbytes) / dt / 1e9, 1), "GB/s D2H")
else:
    for mb in (16, 32, 64, 128, 256, 512):
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, SCVX_HOST_CHUNK_MB=str(mb)))
