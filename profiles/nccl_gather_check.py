"""NCCL check of the result collection: every rank linearises its contiguous trajectory shard on its own
GPU, the 14x23 block slabs are all-gathered over NCCL (sharding.gather_shards) and rank 0 compares the concatenation with
the whole batch computed on one GPU — bit for bit.  Run under torchrun with one rank per GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from successiveconvexification_b200 import dynamics, sample_problems as sp, sharding, workloads

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
prob = sp.base_prob_aero_scaled("tests/golden/aero_lift_drag.npz")
cache = dynamics.make_cache(prob, device_ids=[lr])
B, K = 1001, 20                                                  # ragged shards
X, U, sigma, _ = workloads.monte_carlo_batch(prob, K, B, 7, sigma_range=(0.8, 1.5))
b0, b1 = sharding.shard_range(B, rank, world)
blocks, err, tlb = dynamics.linearize_batch(cache, X[b0:b1], U[b0:b1], sigma[b0:b1], 1 / (K + 1))
full = sharding.gather_shards(torch.from_numpy(blocks).to(dev))
rooted = sharding.gather_to_root(torch.from_numpy(blocks).to(dev), dst=0)        # send/recv into one result on rank 0
ok, chk = sharding.reduce_status(bool(np.isfinite(blocks).all()), float(blocks[:, :, 0, :].sum()), dev)
if rank == 0:
    ref, _, _ = dynamics.linearize_batch(cache, X, U, sigma, 1 / (K + 1))
    same = np.array_equal(full.cpu().numpy(), ref) and np.array_equal(rooted.cpu().numpy(), ref)
    print(f"ranks {world}: gathered {tuple(full.shape)} identical to the single-GPU result: {same}; status ok {ok}; "
          f"checksum {chk:.12e} vs {ref[:, :, 0, :].sum():.12e}")
    assert same and ok
dist.destroy_process_group()
