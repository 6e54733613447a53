"""torchrun --nproc-per-node N tools/pcie_probe_ranks.py: pinned device-to-host bandwidth of every rank's GPU measured alone
(the other ranks idle) and with all ranks copying at the same time -- is the end-to-end gfx export of an N-GPU run limited by
each GPU's PCIe link (aggregate scales with N) or by something the ranks share on the host side (aggregate flat)?"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.pop("NCCL_DEBUG", None)
from fluid_simulator_b200 import dist as fdist
numa = fdist.bind_to_gpu_numa_node(local) if os.environ.get("BIND", "1") == "1" else {"bound": False}
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = (1310965120 // world) // 4
dev = torch.empty(n, dtype=torch.float32, device="cuda")
host = torch.empty(n, dtype=torch.float32).pin_memory()


def once():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


once()
alone = None
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        alone = min(once() for _ in range(3))
if world > 1:
    dist.barrier()
together = min(once() for _ in range(3))  # (ranks start within a barrier's skew of each other)
res = {"rank": rank, "bytes": n * 4, "alone_GBps": n * 4 / alone / 1e9, "together_GBps": n * 4 / together / 1e9, "numa": numa}
allr = [None] * world
if world > 1:
    dist.all_gather_object(allr, res)
else:
    allr = [res]
if rank == 0:
    print(json.dumps({"ranks": world, "aggregate_alone_GBps": sum(r["alone_GBps"] for r in allr) / world, "aggregate_together_GBps": sum(r["together_GBps"] for r in allr),
                      "per_rank": allr}))
if world > 1:
    dist.destroy_process_group()
