"""N>1 path on CPU: world_size-2 gloo run of the trajectory-sharded pipeline.  The per-rank compute is
stood in by the CPU oracle (tests may use it as the checker); what is under test is the host logic —
shard ranges, independent per-shard input streams, rank-ordered gather, status/checksum reduce."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import AERO_NPZ, ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, B, K, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from successiveconvexification_b200 import sample_problems as sp, sharding, workloads
    prob = sp.base_prob_aero_scaled(AERO_NPZ)
    tb = oracle.OracleTables.from_aero(prob.aero)
    X, U, sigma, params = workloads.monte_carlo_batch(prob, K, B, 1003, sigma_range=(0.8, 1.5))
    b0, b1 = sharding.shard_range(B, rank, world)
    blocks, _, _, _ = oracle.linearize_batch(params, tb, X[b0:b1], U[b0:b1], sigma[b0:b1], 1.0 / (K + 1),
                                             want_lin_err=False, want_tlb=False, nthreads=1)
    full = sharding.gather_shards(torch.from_numpy(blocks))
    rooted = sharding.gather_to_root(torch.from_numpy(blocks), dst=0)
    assert (rooted is None) == (rank != 0)
    if rank == 0:
        assert torch.equal(rooted, full)
    ok, csum = sharding.reduce_status(bool(np.isfinite(blocks).all()), float(blocks.sum()), torch.device("cpu"))
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
        np.save(os.path.join(out_dir, "status.npy"), np.array([float(ok), csum]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo_shard_and_gather(tmp_path, prob_aero, oracle_tables):
    from oracle import oracle
    from successiveconvexification_b200 import workloads
    B, K, world = 5, 3, 2                       # uneven shards (2 + 3)
    mp.spawn(_worker, args=(world, _free_port(), B, K, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    status = np.load(tmp_path / "status.npy")
    X, U, sigma, params = workloads.monte_carlo_batch(prob_aero, K, B, 1003, sigma_range=(0.8, 1.5))
    ref, _, _, _ = oracle.linearize_batch(params, oracle_tables, X, U, sigma, 1.0 / (K + 1), want_lin_err=False,
                                          want_tlb=False, nthreads=1)
    assert got.shape == ref.shape == (B, K, 23, 14)
    assert np.array_equal(got, ref)
    assert status[0] == 1.0 and status[1] == pytest.approx(ref.sum(), rel=1e-12)
