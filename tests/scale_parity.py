"""One-step parity of the CUDA path against the COMPILED REFERENCE (oracle/_ref, OpenMP build) at the sizes BASELINE.json
benchmarks -- shared by tests/test_gpu_parity_scale.py (128^3 / 96^3, every `pytest -m gpu` run) and tools/parity_256.py
(the 256^3 headline scene, run once per round; numbers committed under profiles/).

What is compared, on identical fp32-representable initial conditions (simulator.cpp:51-100 on both sides):
  * cell flags after the step and particle->cell indices before it: bit-exact;
  * particle->cell indices after the step: the device's index array must equal ivec3(pos * cellDInv) evaluated in fp64 on
    the positions the device stores (simulator.cpp:358-359) -- bit-exact; against the reference's own post-step indices
    only the particles whose fp32-rounded position crosses a cell boundary may differ (counted and bounded);
  * v, v2, wsum, avgPNum, pressure, particle positions / velocities (and APIC matrices): relative L2 against the reference.
"""
import os
import time

import numpy as np

from fluid_simulator_b200 import abi
from util import rel_l2

FIELDS = [(abi.FIELD_V, "v"), (abi.FIELD_V2, "v2"), (abi.FIELD_WSUM, "wsum"), (abi.FIELD_AVGPNUM, "avgp"),
          (abi.FIELD_PRESSURE, "pressure")]


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cells_of(pos, info):
    """ivec3(pos * cellDInv) flattened in the reference order (simulator.cpp:358-359), fp64 like the reference."""
    gs = info.grid_size
    ix = (pos[:, 0] * info.cell_d_inv[0]).astype(np.int64)
    iy = (pos[:, 1] * info.cell_d_inv[1]).astype(np.int64)
    iz = (pos[:, 2] * info.cell_d_inv[2]).astype(np.int64)
    return (ix * gs[1] * gs[2] + iy * gs[2] + iz).astype(np.int32)


def one_step(gpu_cls, sc, obstacles=None, tol=None, srand=None, apic=False, by_position=False, gpu_kw=None):
    """Runs ONE simulate(dt) of scene `sc` on the device and on the compiled reference; returns the comparison dict."""
    from oracle import refsim
    params = sc.params
    if tol is not None:
        params = abi.Params.from_buffer_copy(bytes(sc.params))
        params.residual_tolerance = tol
    obstacles = list(sc.obstacles if obstacles is None else obstacles)
    res = {"grid": list(sc.dims), "particles": sc.n_particles, "tol": float(params.residual_tolerance)}

    g = gpu_cls(sc.dims, sc.resolution, sc.two_d, sc.particle_radius, **(gpu_kw or {}))
    g.set_params(params)
    g.set_obstacles(obstacles)
    g.upload_particles(sc.particles)
    cells0_g = g.download_particle_cells()
    if srand is not None:
        g.srand(srand)
    t0 = time.perf_counter()
    its_g = g.step(sc.dt)
    g.synchronize()
    res["gpu_step_s"] = time.perf_counter() - t0
    info = g.solve_info()
    res.update(its_gpu=its_g, residual_max_gpu=float(info.residual_max), fluid_cells=int(info.fluid_cells))
    flags_g = g.download_grid(abi.FIELD_TYPE)
    grid_g = {nm: g.download_grid(f) for f, nm in FIELDS}
    pg = g.download_particles(by_id=not by_position)
    cells1_g = g.download_particle_cells(by_id=not by_position)
    gfx_g = g.export_gfx(by_id=not by_position)
    ginfo = g.info
    g.close()
    del g

    r = refsim.RefSim(sc.dims, sc.resolution, sc.two_d, sc.particle_radius)
    r.set_omp_threads(host_threads())
    res["ref_threads"] = r.omp_threads()
    r.set_params(params)
    r.set_obstacles(obstacles)
    r.upload_particles(sc.particles)
    cells0_r = r.download_particle_cells()
    # the manager exports gfx BEFORE simulate() (simulationManager.cpp:218-231); compare the post-step export of both
    if srand is not None:
        r.srand(srand)
    t0 = time.perf_counter()
    its_r = r.step(sc.dt)
    res["ref_step_s"] = time.perf_counter() - t0
    res["its_ref"] = its_r
    flags_r = r.download_grid(abi.FIELD_TYPE)
    pr = r.download_particles()
    cells1_r = r.download_particle_cells()
    gfx_r = r.export_gfx()

    res["cells_before_mismatch"] = int((cells0_g != cells0_r).sum())
    res["type_mismatch"] = int((flags_g != flags_r).sum())
    for f, nm in FIELDS:
        res[nm] = rel_l2(grid_g[nm], r.download_grid(f))
    r.close()
    del r, grid_g

    res["count_gpu"], res["count_ref"] = int(pg.shape[0]), int(pr.shape[0])
    if pg.shape == pr.shape:
        if by_position:  # after removals the device order is cell-binned and the reference's is compacted: compare as sets
            kg = np.lexsort((pg[:, 2].astype(np.float32), pg[:, 1].astype(np.float32), pg[:, 0].astype(np.float32)))
            kr = np.lexsort((pr[:, 2].astype(np.float32), pr[:, 1].astype(np.float32), pr[:, 0].astype(np.float32)))
            pg, cells1_g, gfx_g = pg[kg], cells1_g[kg], gfx_g[kg]
            pr, cells1_r, gfx_r = pr[kr], cells1_r[kr], gfx_r[kr]
        res["pos"] = rel_l2(pg[:, 0:3], pr[:, 0:3])
        res["vel"] = rel_l2(pg[:, 3:6], pr[:, 3:6])
        if apic:
            res["c"] = rel_l2(pg[:, 6:15], pr[:, 6:15])
        # the device's index is the reference expression on the device's stored position: bit-exact
        res["cells_after_vs_own_positions"] = int((cells1_g != cells_of(pg, ginfo)).sum())
        # against the reference's own indices only boundary-crossing fp32 roundings differ
        res["cells_after_vs_reference"] = int((cells1_g != cells1_r).sum())
        # gfx export (manager/simulationManager.cpp:218-231): float pos, |v|, density
        res["gfx_pos"] = rel_l2(gfx_g[:, 0:3], gfx_r[:, 0:3])
        res["gfx_speed"] = rel_l2(gfx_g[:, 3], gfx_r[:, 3])
        res["gfx_density"] = rel_l2(gfx_g[:, 4], gfx_r[:, 4])
    return res
