"""Measurement protocol of the multi-rank benchmark (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

`bench.py --gpus N` either cuts ONE domain into z-slabs (fluid_simulator_b200.slab, csrc/dist.cu: ghost planes, particle
migration and the PCG's reductions travel through NVLink peer memory inside the library) or, with `--shard replicas`, runs
an independent domain per GPU.  In both modes torch.distributed only carries the protocol below: barrier, max-over-ranks
of the device time, sum of the units processed.
"""
import os


def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def replica_seed(base_seed, rank):
    """Every replica gets its own jitter stream so that ranks do not simulate bit-identical scenes."""
    return int(base_seed) + 7919 * int(rank)


def aggregate(dist, device, local_units, local_ms, world):
    """Whole-job throughput: sum of the units all ranks processed / max over ranks of the device time.
    Returns (total_units, max_ms, units_per_second)."""
    import torch
    t = torch.tensor([float(local_units), float(local_ms)], dtype=torch.float64, device=device)
    if world > 1:
        units = t[0:1].clone()
        ms = t[1:2].clone()
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        total_units, max_ms = float(units.item()), float(ms.item())
    else:
        total_units, max_ms = float(local_units), float(local_ms)
    return total_units, max_ms, total_units / (max_ms * 1e-3)
