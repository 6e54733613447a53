"""N>1 path on CPU: two gloo ranks run the benchmark's measurement protocol (replica seeds, barrier, max-over-ranks time,
summed units) -- the cross-rank logic that lives in Python (the slab exchanges themselves are CUDA kernels,
tests/test_gpu_slab.py, and their host-side arithmetic is covered by tests/test_slab_host.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fluid_simulator_b200 import dist as fdist
from fluid_simulator_b200 import scenes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, w, lr = fdist.rank_info()
    assert (r, w, lr) == (rank, world, rank)
    # every replica builds its own scene: different jitter, same particle count
    pos = scenes.block_positions_f32(1, 4, 1, 7, 1, 7, seed=fdist.replica_seed(scenes.SEED, rank))
    units = pos.shape[0] * 5                    # 5 steps
    ms = 10.0 + 3.0 * rank                      # rank 1 is slower: the job time is the max
    dist.barrier()
    total, max_ms, value = fdist.aggregate(dist, "cpu", units, ms, world)
    checksum = torch.tensor([float(pos.sum())], dtype=torch.float64)
    gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, checksum)
    if rank == 0:
        out.put((total, max_ms, value, [float(g.item()) for g in gathered], pos.shape[0]))
    dist.destroy_process_group()


def test_two_rank_replica_protocol():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    total, max_ms, value, sums, n = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == 2 * n * 5                   # whole-job units over both ranks
    assert max_ms == 13.0                       # max over ranks, not the mean
    assert abs(value - total / 13.0e-3) < 1e-6 * value
    assert sums[0] != sums[1]                   # replicas are not bit-identical scenes


def test_single_rank_aggregate_is_identity():
    total, ms, value = fdist.aggregate(None, "cpu", 1000, 2.0, 1)
    assert (total, ms) == (1000.0, 2.0) and abs(value - 5.0e5) < 1e-9
