"""SURVEY.md §8d cfg 5: projection-only sweep.  For N in the list: dam block WATER, v = 0, postP2GUpdate(g*dt) (hydrostatic
start), one pressure solve to ||r||_inf < 1e-6.  Prints one JSON line per N with iterations, ms/solve, ms/iteration and
the residual after the solve.  One GPU:  python tools/bench_projection.py 128 192 256 384 512
Several GPUs (the grid cut into z-slabs):  torchrun --nproc-per-node G tools/bench_projection.py 256 512"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fluid_simulator_b200 import abi, scenes
from fluid_simulator_b200.sim import FluidSim

os.environ["FSIM_NO_WARM_START"] = "1"   # every repetition is a cold solve (p = 0 start), like the reference's

rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    from fluid_simulator_b200 import slab
    os.environ.pop("NCCL_DEBUG", None)
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

for n in [int(x) for x in sys.argv[1:]] or [128, 256]:
    sim = FluidSim((float(n),) * 3, 1.0, False, 0.25, device=local_rank, rank=rank, nranks=world)
    if world > 1:
        slab.connect_torch(sim, dist)
    sim.set_params(scenes.default_params(abi.FLIP, pressure_enabled=False, max_iterations=2000, tol=1e-6))
    z0, gzl = int(sim.slab.z_offset), int(sim.slab.gz_local)
    types = scenes.hydrostatic_types(n).reshape(n, n, n)[:, :, z0:z0 + gzl].reshape(-1)   # the block this handle stores
    res = []
    for rep in range(3):
        sim.upload_grid(abi.FIELD_TYPE, types)
        sim.upload_grid(abi.FIELD_V, np.zeros((n * n * gzl, 3)))
        sim.post_p2g_update(-39.24 * 0.005)
        sim.synchronize()
        if dist:
            dist.barrier()
        sim.timer_record(0)
        its = sim.stage_project(0.005)
        sim.timer_record(1)
        sim.synchronize()
        res.append((its, sim.timer_elapsed_ms(0, 1)))
    info = sim.solve_info()
    its, ms = res[-1]   # repetition 0 pays the one-time hierarchy allocation and graph capture
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    nf = int(info.fluid_cells)
    if rank == 0:
        print(json.dumps({"cfg": 5, "grid": n, "gpus": world, "slab_solver": os.environ.get("FSIM_SLAB_SOLVER", "hybrid") if world > 1 else None,
                          "fluid_cells": nf, "iterations": its, "ms_per_solve": ms,
                          "ms_per_iteration": ms / max(its, 1), "first_solve_ms_incl_setup": res[0][1],
                          "algorithmic_GBps": (its * nf * 174 + n ** 3 * 60) / (ms * 1e-3) / 1e9,
                          "residual_max": info.residual_max, "reference_mic0_iterations": {64: 41, 128: 80, 256: 155}.get(n)}), flush=True)
    if dist:
        dist.barrier()
    sim.close()
if dist:
    dist.destroy_process_group()
