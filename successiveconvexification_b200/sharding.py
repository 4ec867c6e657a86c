"""Trajectory sharding across GPUs (one process per GPU) — SURVEY.md §8e.

Intervals and trajectories are independent (reference dynamics.jl:324-332 reads nodes i, i+1 only), so the
path shards by contiguous blocks of trajectories with NO data-path collective.  `torch.distributed`
(NCCL on GPUs, gloo in the CPU tests) is used only to gather results / status afterwards.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous trajectory block [b0, b1) of `rank` (same rule as the in-library multi-device split)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return B * rank // world, B * (rank + 1) // world


def gather_shards(local, group=None):
    """all_gather of equally- or unequally-sized leading-axis shards; returns the concatenation in rank
    order (every rank).  `local` is a torch tensor on the backend's device."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))])
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def gather_to_root(local, dst: int = 0, group=None):
    """Result collection for a host-side consumer: every rank sends its leading-axis shard to `dst`,
    which receives each one straight into its slice of ONE result tensor — no padding and no copy on the other ranks
    (`gather_shards` leaves the whole result on every rank: world x the memory, for device-side consumers only).
    Returns the concatenation in rank order on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    local = local.contiguous()
    if rank != dst:
        if sizes[rank]:
            dist.send(local, dst, group=group)
        return None
    out = local.new_empty((sum(sizes),) + tuple(local.shape[1:]))
    offs = [sum(sizes[:r]) for r in range(world)]
    out[offs[dst]:offs[dst] + sizes[dst]] = local
    for r in range(world):
        if r != dst and sizes[r]:
            dist.recv(out[offs[r]:offs[r] + sizes[r]], r, group=group)
    return out


def reduce_status(ok: bool, checksum: float, device, group=None):
    """Tiny status / checksum all-reduce: (all ranks ok?, sum of per-shard checksums)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([0.0 if ok else 1.0, float(checksum)], dtype=torch.float64, device=device)
    dist.all_reduce(t, group=group)
    return bool(t[0].item() == 0.0), float(t[1].item())
