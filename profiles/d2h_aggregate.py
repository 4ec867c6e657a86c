"""Aggregate pinned-host D2H bandwidth of a node: every rank copies at once vs rank 0 alone (torchrun, one rank per GPU).
Explains why the end-to-end figure of bench.py does not scale with the number of GPUs while the device-resident one does."""
import glob, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import bind_to_gpu_numa

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
note = bind_to_gpu_numa(lr) if os.environ.get("NO_BIND") is None else "unbound"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")

def bw(reps=8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9

bw(2)
if world > 1: dist.barrier()
alone = bw() if rank == 0 else 0.0
if world > 1: dist.barrier()
together = bw()
t = torch.tensor([together], device="cuda")
if world > 1:
    dist.all_reduce(t)
if rank == 0:
    nodes = len(glob.glob("/sys/devices/system/node/node[0-9]*"))
    print(f"ranks {world}  numa nodes {nodes}  cpus {os.cpu_count()}  {note}")
    print(f"D2H GB/s: rank 0 alone {alone:.1f}; all ranks at once: sum {t.item():.1f} (rank 0: {together:.1f})")
if world > 1:
    dist.destroy_process_group()
