import numpy as np, sys
from successiveconvexification_b200 import dynamics, sample_problems as sp, workloads
prob = sp.base_prob_aero_scaled('tests/golden/aero_lift_drag.npz')
cache = dynamics.make_cache(prob); ctx = cache.sim_prob
ctx.set_kernel(int(sys.argv[1]))
X,U,s,dt = workloads.sample_trajectory(prob)
b,e,t = dynamics.linearize_batch(cache, X,U,s,dt)
print('ok', np.isfinite(b).all(), b[0,0,:2,:3])
