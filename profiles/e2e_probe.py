"""Where does the end-to-end host-pointer path lose D2H bandwidth against a raw pinned copy?  (one GPU)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
dO = torch.empty((B, K, 23, 14), dtype=torch.float64, device='cuda')

def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

# (a) raw D2H of the block array in 256 MiB pieces on two alternating streams
st = [torch.cuda.Stream(), torch.cuda.Stream()]
flatd, flath = dO.view(-1), hO.view(-1)
piece = (256 << 20) // 8
def raw():
    for i, off in enumerate(range(0, flatd.numel(), piece)):
        with torch.cuda.stream(st[i & 1]):
            flath[off:off + piece].copy_(flatd[off:off + piece], non_blocking=True)
t = timeit(raw); print(f"raw D2H, 256 MiB pieces, 2 streams: {hO.nbytes / t / 1e9:.1f} GB/s")
def raw1():
    hO.copy_(dO, non_blocking=True)
t = timeit(raw1); print(f"raw D2H, one copy: {hO.nbytes / t / 1e9:.1f} GB/s")
# (b) the library path
def e2e(err=True, tlb=True):
    ctx.linearize_ptr(hX.data_ptr(), hU.data_ptr(), hS.data_ptr(), 1 / 51, 10, 0, K + 1, B, hO.data_ptr(),
                      hE.data_ptr() if err else 0, hT.data_ptr() if tlb else 0)
t = timeit(e2e); print(f"library host path: {B * K / t / 1e6:.2f} M intervals/s, {(hO.nbytes + hE.nbytes + hT.nbytes) / t / 1e9:.1f} GB/s D2H")
t = timeit(lambda: e2e(False, False)); print(f"library host path, blocks only: {B * K / t / 1e6:.2f} M intervals/s, {hO.nbytes / t / 1e9:.1f} GB/s D2H")
# (c) raw D2H while the kernels run on another stream
dX, dU, dS = (torch.from_numpy(a).cuda() for a in (X, U, s))
dO2 = torch.empty((B, K, 23, 14), dtype=torch.float64, device='cuda')
ks = torch.cuda.Stream()
ctx.set_stream(ks.cuda_stream)
def both():
    ctx.linearize_ptr(dX.data_ptr(), dU.data_ptr(), dS.data_ptr(), 1 / 51, 10, 0, K + 1, B, dO2.data_ptr())
    ctx.linearize_ptr(dX.data_ptr(), dU.data_ptr(), dS.data_ptr(), 1 / 51, 10, 0, K + 1, B, dO2.data_ptr())
    ctx.linearize_ptr(dX.data_ptr(), dU.data_ptr(), dS.data_ptr(), 1 / 51, 10, 0, K + 1, B, dO2.data_ptr())
    ctx.linearize_ptr(dX.data_ptr(), dU.data_ptr(), dS.data_ptr(), 1 / 51, 10, 0, K + 1, B, dO2.data_ptr())
    ctx.linearize_ptr(dX.data_ptr(), dU.data_ptr(), dS.data_ptr(), 1 / 51, 10, 0, K + 1, B, dO2.data_ptr())
    raw1()
t = timeit(both); print(f"raw D2H (one copy) with the kernels running beside it: {hO.nbytes / t / 1e9:.1f} GB/s (5 kernel steps = {5 * 17:.0f} ms)")
