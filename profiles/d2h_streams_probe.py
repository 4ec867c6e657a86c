"""Pinned D2H rate of one GPU as a function of the number of concurrent copy streams and the chunk size (is one DMA engine
enough to saturate this pool's PCIe?).  python profiles/d2h_streams_probe.py"""
import time, torch
dev = torch.device("cuda", 0)
tot = 2 << 30
d = torch.empty(tot, dtype=torch.uint8, device=dev)
h = torch.empty(tot, dtype=torch.uint8).pin_memory()
for chunk_mb in (64, 256):
    for ns in (1, 2, 4):
        streams = [torch.cuda.Stream() for _ in range(ns)]
        chunk = chunk_mb << 20
        n = tot // chunk
        def run():
            for k in range(n):
                with torch.cuda.stream(streams[k % ns]):
                    h[k * chunk:(k + 1) * chunk].copy_(d[k * chunk:(k + 1) * chunk], non_blocking=True)
            torch.cuda.synchronize()
        run()
        t0 = time.perf_counter(); run(); run(); dt = (time.perf_counter() - t0) / 2
        print(f"chunk {chunk_mb:4d} MiB  streams {ns}: {tot / dt / 1e9:6.1f} GB/s")
