"""Measures pinned D2H bandwidth for one copy vs the same bytes split over 2 / 4 streams (how bench.py's e2e copy should be issued)."""
import time
import torch

n = 1310965120 // 4
dev = torch.empty(n, dtype=torch.float32, device="cuda")
host = torch.empty(n, dtype=torch.float32).pin_memory()
for parts in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    chunk = (n + parts - 1) // parts
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                host[i * chunk:(i + 1) * chunk].copy_(dev[i * chunk:(i + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"D2H {n * 4 / 1e9:.2f} GB in {parts} stream(s): {best * 1e3:.2f} ms = {n * 4 / best / 1e9:.1f} GB/s")
