"""Worker of tests/test_gpu_slab.py::test_two_processes_ipc: one process per slab rank (launched by torchrun), the
ranks' exports exchanged through torch.distributed (gloo), neighbours mapped over CUDA IPC.  Rank 0 compares the
stitched result with a single-handle run of the same scene."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from fluid_simulator_b200 import abi, scenes, slab
    from fluid_simulator_b200.sim import FluidSim
    from util import rel_l2
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    dev = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
    grid = int(os.environ.get("FSIM_WORKER_GRID", "16"))
    if rank == 0:
        print(f"slab worker: {world} ranks, devices {[r % torch.cuda.device_count() for r in range(world)]} of {torch.cuda.device_count()}, grid {grid}")

    def compare_scene(tag, sc, steps=3):
        """`steps` steps of the scene on the slab group and, on rank 0, on a single handle: flags equal, fields and particles 1e-5."""
        s = FluidSim(sc.dims, sc.resolution, sc.two_d, sc.particle_radius, device=dev, rank=rank, nranks=world, capacity=sc.n_particles)
        slab.connect_torch(s, dist)
        s.set_params(sc.params)
        s.set_obstacles([])
        own = slab.owner_of(sc.particles[:, 2], s.info.cell_d_inv[2], s.grid_size[2], world)
        idx = np.nonzero(own == rank)[0]
        s.upload_particles(sc.particles[idx])
        s.upload_particle_ids(idx.astype(np.uint32))
        dist.barrier()
        its = [s.step(sc.dt) for _ in range(steps)]
        s.synchronize()
        dist.barrier()
        mine = {"its": its, "p": s.download_particles(by_id=False), "ids": s.download_particle_ids(),
                "v2": s.download_grid(abi.FIELD_V2), "type": s.download_grid(abi.FIELD_TYPE)}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        dist.barrier()
        s.close()
        if rank == 0:
            one = FluidSim(sc.dims, sc.resolution, sc.two_d, sc.particle_radius, device=dev)
            one.set_params(sc.params)
            one.upload_particles(sc.particles)
            its1 = [one.step(sc.dt) for _ in range(steps)]
            parts = slab.partition(one.grid_size[2], world)
            v2 = slab.stitch([g["v2"] for g in gathered], parts, one.grid_size, (3,)).reshape(-1, 3)
            ty = slab.stitch([g["type"] for g in gathered], parts, one.grid_size).reshape(-1)
            p = np.concatenate([g["p"] for g in gathered])[np.argsort(np.concatenate([g["ids"] for g in gathered]), kind="stable")]
            assert all(g["its"] == gathered[0]["its"] for g in gathered), "ranks disagree on iteration counts"
            assert np.array_equal(ty, one.download_grid(abi.FIELD_TYPE)), f"{tag}: cell flags differ"
            e_v2 = rel_l2(v2, one.download_grid(abi.FIELD_V2))
            e_p = rel_l2(p[:, 0:6], one.download_particles()[:, 0:6])
            print(f"{tag}its slab={gathered[0]['its']} single={its1} v2 rel L2 {e_v2:.3e} particles rel L2 {e_p:.3e}")
            assert e_v2 <= 1e-5 and e_p <= 1e-5, tag
            one.close()
        dist.barrier()

    compare_scene("", scenes.dam_break_3d(grid, abi.FLIP, tol=1e-9))
    # push-apart (boundary-plane positions pulled from the neighbours, second migration) and the BasicMacGrid solver (colour-masked
    # hand-over of the shared z faces) over the same fabric
    compare_scene("push-apart: ", scenes.dam_break_3d(grid, abi.FLIP, tol=1e-9, push_apart_enabled=True))
    compare_scene("basic solver: ", scenes.dam_break_3d(grid, abi.PIC, solver_type=abi.SOLVER_BASIC, max_iterations=20))

    # ---- second scene: PIC with a particle source, a sink and a moving box (SURVEY cfg 4 style): spawning draws from libc
    # rand(), so every rank (= process) seeds it identically and keeps the spawned particles that start in its planes
    import cases
    sc, obs_fn, steps = cases.CASES["3d_pic_source_sink"]()

    def drive(sim):
        obs = list(sc.obstacles)
        for st in range(steps):
            new = obs_fn(st)
            if st > 0:
                old = sim.get_obstacles(obs)
                for i, x in enumerate(new):
                    x.last_spawn_fraction = old[i].last_spawn_fraction
            obs = new
            sim.set_obstacles(obs)
            sim.srand(1000 + st)
            sim.step(sc.dt)
        sim.synchronize()

    s = FluidSim(sc.dims, sc.resolution, sc.two_d, sc.particle_radius, device=dev, rank=rank, nranks=world, capacity=2 * sc.n_particles)
    slab.connect_torch(s, dist)
    s.set_params(sc.params)
    own = slab.owner_of(sc.particles[:, 2], s.info.cell_d_inv[2], s.grid_size[2], world)
    idx = np.nonzero(own == rank)[0]
    s.upload_particles(sc.particles[idx])
    s.upload_particle_ids(idx.astype(np.uint32))
    dist.barrier()
    drive(s)
    dist.barrier()
    mine = {"p": s.download_particles(by_id=False), "ids": s.download_particle_ids(), "v2": s.download_grid(abi.FIELD_V2),
            "type": s.download_grid(abi.FIELD_TYPE)}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    dist.barrier()
    s.close()
    if rank == 0:
        one = FluidSim(sc.dims, sc.resolution, sc.two_d, sc.particle_radius, device=dev)
        one.set_params(sc.params)
        one.upload_particles(sc.particles)
        drive(one)
        parts = slab.partition(one.grid_size[2], world)
        v2 = slab.stitch([g["v2"] for g in gathered], parts, one.grid_size, (3,)).reshape(-1, 3)
        ty = slab.stitch([g["type"] for g in gathered], parts, one.grid_size).reshape(-1)
        p = np.concatenate([g["p"] for g in gathered])[np.argsort(np.concatenate([g["ids"] for g in gathered]), kind="stable")]
        p1 = one.download_particles()
        print(f"source/sink scene: {sc.n_particles} initial particles -> {p.shape[0]} (slabs) vs {p1.shape[0]} (single)")
        assert p.shape == p1.shape and p.shape[0] != sc.n_particles, "spawn / despawn did not change the particle count identically"
        assert np.array_equal(ty, one.download_grid(abi.FIELD_TYPE)), "cell flags differ (source/sink scene)"
        e_v2 = rel_l2(v2, one.download_grid(abi.FIELD_V2))
        e_p = rel_l2(p[:, 0:6], p1[:, 0:6])
        print(f"source/sink scene: v2 rel L2 {e_v2:.3e} particles rel L2 {e_p:.3e}")
        assert e_v2 <= 1e-5 and e_p <= 1e-5
        one.close()
        print("SLAB_WORKER_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
