#!/usr/bin/env python
"""bench.py -- headline benchmark of the Simulator hot path (BASELINE.json: 3D FLIP particle-updates/s and ms/step at
256^3, fraction of the HBM roofline), plus the reference arm (`--impl reference`: the unmodified reference Simulator
compiled into oracle/_ref, timed on the box's host cores).

One "step" = one Simulator::simulate(dt) over the synthetic dam-break scene: advect -> bin/sort ->
P2G -> classify -> pressure projection (PCG) -> extrapolate -> G2P.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "particle_updates_per_s"
UNIT = "particle-updates/s"
DT = 0.005

# ALGORITHMIC bytes: compulsory fp32-SoA traffic only
B_PARTICLE = {0: 72, 1: 84, 2: 144}   # PIC / FLIP / APIC bytes per particle-update (two compulsory particle passes)
B_CELL = 205                          # bytes per cell per step for the grid stages
B_IT_SURVEY = 136                     # CG core 53 B + multigrid V(2,2) 83 B per fluid cell and iteration (all-fp32 vectors)
B_IT = 174                            # bytes per fluid cell per PCG iteration: fp64 CG core 86 + multigrid cycle 88


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first(self, timeout=3.0):
        """nvidia-smi needs a moment to start: wait for its first answer so that the timed region is covered."""
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def scene_params(n, transfer, tol=1e-6):
    from fluid_simulator_b200 import scenes
    return scenes.default_params(transfer, max_iterations=2000, tol=tol)


def run_reference(args, n, steps, warmup):
    """Reference arm / cpu_baseline: the unmodified reference Simulator (oracle/_ref) on the host cores, on a bounded
    sample of the workload: the same dam-break recipe on an n^3 grid for `steps` steps after `warmup` (the --impl
    reference arm runs the metric's own 256^3 scene for one step; the cpu_baseline leg of the B200 arm a 96^3 one)."""
    from fluid_simulator_b200 import abi, scenes
    from oracle import refsim
    kind = "reference"
    if refsim.available():
        sc = scenes.dam_break_3d(n, abi.FLIP)
        sim = refsim.RefSim(sc.dims, sc.resolution, sc.two_d, sc.particle_radius)
        # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1: override it for the reference arm)
        sim.set_omp_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
        cores = sim.omp_threads()
    else:  # the compiled reference always travels with the repo; the plain-C restatement is the documented stand-in
        from oracle.oracle import OracleSim
        kind = "port"
        n = min(n, 64)
        sc = scenes.dam_break_3d(n, abi.FLIP)
        sim = OracleSim(sc.dims, sc.resolution, sc.two_d, sc.particle_radius)
        cores = 1
    sim.set_params(sc.params)
    sim.upload_particles(sc.particles)
    for _ in range(warmup):
        sim.step(DT)
    t0 = time.perf_counter()
    its = [sim.step(DT) for _ in range(steps)]
    el = time.perf_counter() - t0
    value = sc.n_particles * steps / el
    same = "the GPU workload itself" if n == args.grid else "same recipe as the GPU workload, smaller grid"
    return {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "ms_per_step": 1e3 * el / steps, "grid": n, "steps": steps,
            "warmup": warmup, "particles": sc.n_particles,
            "sample": f"3D FLIP dam break {n}^3, {sc.n_particles} particles, {steps} step(s) after {warmup} warm-up "
                      f"({same}), PCG its {its}"}


def reference_plan(args):
    """Grid / steps / warm-up of the --impl reference arm.  Default: the metric's own scene (--grid, 256^3) so that the
    driver's ratio is a same-config ratio; one reference step there costs minutes (MIC(0)-PCG: ~155 iterations with strictly
    sequential triangular solves), so the K / W of the command line are cut down to what fits ~4 minutes of host time."""
    n = args.cpu_grid if args.cpu_grid else args.grid
    est = 1.5 * (n / 96.0) ** 4.7  # s per step: 1.4 s measured at 96^3 on 16 threads, ~150 s at 256^3 (the triangular solves are serial)
    steps = max(1, min(args.steps, int(240.0 / est)))
    warmup = min(args.warmup, 1 if est * (steps + 1) < 300.0 else 0)
    return n, steps, warmup


def verify_slabs(args, dist, rank, world, local_rank, red_dev, n=64, steps=3):
    """Pre-flight of the N-GPU path ON the N GPUs it is about to be timed on: a 64^3 dam break advanced `steps` steps by
    the slab group (one rank per GPU, exchanges over NVLink peer memory) and by a single handle on rank 0; cell flags and
    particle->cell indices must agree bit for bit, v2 / pressure / particle state to 1e-5 relative L2 (the tolerances of
    tests/test_gpu_slab.py, which can only run all ranks on one device)."""
    import numpy as np
    from fluid_simulator_b200 import abi, scenes
    from fluid_simulator_b200 import slab as fslab
    from fluid_simulator_b200.sim import FluidSim
    sc = scenes.dam_break_3d(n, abi.FLIP, tol=1e-9)
    own = fslab.owner_of(sc.particles[:, 2], 1.0, n, world)
    idx = np.nonzero(own == rank)[0]
    out = {"grid": n, "steps": steps, "ranks": world}
    try:
        s = FluidSim(sc.dims, sc.resolution, sc.two_d, sc.particle_radius, capacity=int(idx.size * 1.5) + 4096, device=local_rank, rank=rank, nranks=world)
        fslab.connect_torch(s, dist)
        s.set_params(sc.params)
        s.upload_particles(sc.particles[idx])
        s.upload_particle_ids(idx.astype(np.uint32))
        dist.barrier()
        its = [s.step(sc.dt) for _ in range(steps)]
        s.synchronize()
        mine = {"its": its, "ids": s.download_particle_ids(), "particles": s.download_particles(by_id=False),
                "cells": s.download_particle_cells(by_id=False),
                "type": s.download_grid(abi.FIELD_TYPE), "v2": s.download_grid(abi.FIELD_V2), "pressure": s.download_grid(abi.FIELD_PRESSURE)}
        dist.barrier()
        s.close()
        err = None
    except Exception as ex:  # noqa: BLE001
        mine, err = None, f"rank {rank}: {ex}"
    allr = [None] * world
    dist.all_gather_object(allr, (mine, err))
    errs = [e for _, e in allr if e]
    if errs:
        out.update(ok=False, error="; ".join(errs))
        return out
    if rank == 0:
        from numpy.linalg import norm
        g = FluidSim(sc.dims, sc.resolution, sc.two_d, sc.particle_radius, device=local_rank)
        g.set_params(sc.params)
        g.upload_particles(sc.particles)
        its1 = [g.step(sc.dt) for _ in range(steps)]
        parts = fslab.partition(n, world)
        rel = lambda a, b: float(norm((a - b).ravel()) / max(norm(b.ravel()), 1e-300))
        ty = fslab.stitch([m["type"] for m, _ in allr], parts, (n, n, n)).reshape(-1)
        v2 = fslab.stitch([m["v2"] for m, _ in allr], parts, (n, n, n), (3,)).reshape(-1, 3)
        pr = fslab.stitch([m["pressure"] for m, _ in allr], parts, (n, n, n)).reshape(-1)
        ids = np.concatenate([m["ids"] for m, _ in allr])
        order = np.argsort(ids, kind="stable")
        pa = np.concatenate([m["particles"] for m, _ in allr])[order]
        ce = np.concatenate([m["cells"] for m, _ in allr])[order]
        p1, c1 = g.download_particles(), g.download_particle_cells()
        out.update(its_slab=allr[0][0]["its"], its_single=its1,
                   type_mismatch=int((ty != g.download_grid(abi.FIELD_TYPE)).sum()),
                   particles_lost_or_duplicated=int(pa.shape[0] != p1.shape[0] or not np.array_equal(np.sort(ids), np.arange(p1.shape[0], dtype=ids.dtype))),
                   v2=rel(v2, g.download_grid(abi.FIELD_V2)), pressure=rel(pr, g.download_grid(abi.FIELD_PRESSURE)))
        if pa.shape == p1.shape:
            out.update(cells_mismatch=int((ce != c1).sum()), pos=rel(pa[:, 0:3], p1[:, 0:3]), vel=rel(pa[:, 3:6], p1[:, 3:6]))
        g.close()
        out["ok"] = bool(out["type_mismatch"] == 0 and out["particles_lost_or_duplicated"] == 0 and out.get("cells_mismatch", 1) == 0
                         and max(out["v2"], out["pressure"], out.get("pos", 1.0), out.get("vel", 1.0)) <= 1e-5)
    res = [out]
    dist.broadcast_object_list(res, src=0)
    return res[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=256, help="N of the N^3 dam-break grid (256 = BASELINE.json metric config)")
    ap.add_argument("--transfer", default="FLIP", choices=["PIC", "FLIP", "APIC"])
    ap.add_argument("--cpu-grid", type=int, default=0, help="grid of the reference run: --impl reference defaults to --grid (the metric's "
                    "own scene), the cpu_baseline leg of the B200 arm to 96")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--cpu-warmup", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the pre-flight that compares a small slab run with the single handle")
    ap.add_argument("--obstacle-box", action="store_true", help="add the static box of SURVEY cfg 3 (size (24,96,N) at (0.7N,48,0.5N))")
    ap.add_argument("--shard", default=os.environ.get("FSIM_BENCH_SHARD", "slab"), choices=["slab", "replicas"],
                    help="N > 1: 'slab' cuts ONE N^3 domain into z-slabs (strong scaling; halos, migration and reductions over NVLink peer "
                         "memory), 'replicas' runs an independent domain per GPU (weak scaling, no data-path exchange)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries ONE JSON line and nothing else: libraries that write to fd 1 (NCCL prints its version banner there when
    # NCCL_DEBUG is set, whatever NCCL_DEBUG_FILE says) are sent to stderr, where the log stays visible
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        json_out.write(json.dumps(obj) + "\n")
        json_out.flush()

    if args.impl == "reference":
        if rank != 0:
            return
        n_ref, k_ref, w_ref = reference_plan(args)
        try:
            r = run_reference(args, n_ref, k_ref, w_ref)
        except (MemoryError, OSError) as ex:  # a host too small for the 256^3 reference state (~20 GB): cfg 2's grid instead
            print(f"bench.py: reference run at {n_ref}^3 failed ({ex}); falling back to 128^3", file=sys.stderr)
            r = run_reference(args, 128, min(args.steps, 3), min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
                "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": r["sample"], "grid": [r["grid"]] * 3, "particles_total": r["particles"],
                           "parallelism": f"host cores x{r['cores']} (OpenMP)"},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from fluid_simulator_b200 import abi, scenes
    from fluid_simulator_b200.sim import FluidSim

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    # FSIM_BENCH_SAME_DEVICE=1 (debugging on a one-GPU box): all ranks share cuda:0 and the plumbing runs over gloo
    same_dev = os.environ.get("FSIM_BENCH_SAME_DEVICE") == "1"
    if same_dev:
        local_rank = 0
    red_dev = "cpu" if same_dev else "cuda"
    torch.cuda.set_device(local_rank)
    from fluid_simulator_b200 import dist as fdist
    numa = fdist.bind_to_gpu_numa_node(local_rank) if world > 1 else {"bound": False, "why": "single rank"}
    if world > 1:
        # NCCL_DEBUG at VERSION / WARN or above prints on stdout, in front of the JSON line: keep the setting, but send the log to
        # a file per rank (gpurun_out/nccl_<host>_<pid>.log); rank 0 echoes the communicator lines to stderr at the end
        if "FSIM_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["FSIM_NCCL_DEBUG"]
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            os.environ["NCCL_DEBUG_FILE"] = os.path.join(ROOT, "gpurun_out", "nccl_%h_%p.log")
        if same_dev:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    transfer = {"PIC": abi.PIC, "FLIP": abi.FLIP, "APIC": abi.APIC}[args.transfer]
    n = args.grid
    t_gen = time.perf_counter()
    from fluid_simulator_b200 import dist as fdist
    from fluid_simulator_b200 import slab as fslab
    slab_mode = world > 1 and args.shard == "slab"
    slab_fallback = None
    own_lo, own_hi, zoff, gzl = 0, n, 0, n
    if slab_mode:
        # strong scaling: ONE N^3 dam break cut into z-slabs; every rank builds the part of the dam block that
        # lies in the planes it owns (the jitter keeps a particle inside its cell, so ownership holds by construction)
        own_lo, own_hi, zoff, gzl = fslab.partition(n, world)[rank]
        pos = scenes.block_positions_f32(1, n // 2, 1, n - 1, max(own_lo, 1), min(own_hi, n - 1), seed=fdist.replica_seed(scenes.SEED, rank))
    else:
        # one GPU, or (--shard replicas) every rank advances its own N^3 dam break
        pos = scenes.block_positions_f32(1, n // 2, 1, n - 1, 1, n - 1, seed=fdist.replica_seed(scenes.SEED, rank))
    np_local = pos.shape[0]

    def make_sim():
        """The handle of this rank with the scene uploaded (slab mode: connected to its neighbours); collective."""
        nonlocal slab_mode, slab_fallback
        sim = None
        if slab_mode:
            try:
                sim = FluidSim((float(n),) * 3, 1.0, False, 0.25, capacity=int(np_local * 1.25) + 1024, device=local_rank, rank=rank, nranks=world)
                ok = None
            except Exception as ex:  # noqa: BLE001
                ok = f"rank {rank}: {ex}"
            oks = [None] * world
            dist.all_gather_object(oks, ok)
            if not any(oks):
                try:
                    fslab.connect_torch(sim, dist)  # raises on every rank if any rank could not map its neighbours (e.g. no CUDA IPC)
                except RuntimeError as ex:
                    oks = [str(ex)]
            if any(oks):  # consistent on all ranks: nothing to fall back to with a slab-local particle set -- fail loudly
                raise SystemExit("bench.py: z-slab set-up failed: " + "; ".join(o for o in oks if o))
        else:
            sim = FluidSim((float(n),) * 3, 1.0, False, 0.25, capacity=np_local, device=local_rank)
        sim.set_id_tracking(False)  # ids are test support; the hot path does not carry them
        sim.set_params(scene_params(n, transfer))
        if args.obstacle_box:
            sim.set_obstacles([scenes.cfg3_box(n)])
        sim.upload_particles_f32(pos)
        return sim

    sim = make_sim()
    nc_local = n * n * gzl if slab_mode else n ** 3
    t_gen = time.perf_counter() - t_gen

    def barrier():
        sim.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def allsum(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=red_dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=red_dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    verify = None
    if slab_mode and not args.no_verify:
        verify = verify_slabs(args, dist, rank, world, local_rank, red_dev)  # pre-flight: slabs vs the single handle, small scene

    if world > 1:
        dist.barrier()  # slab steps are collective: start them together (a rank still setting up would eat into the exchange time-out)
    # the clock sampler (nvidia-smi every 25 ms) runs from the warm-up on: at N = 8 the timed region is ~50 ms long
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    for _ in range(args.warmup):
        sim.step(DT)
    sim.profile_read(reset=True)  # zero the launch counters
    if slab_mode:
        sim.dist_wait_stats(reset=True)

    barrier()
    n_before = len(sampler.lines)
    sim.timer_record(0)
    t0 = time.perf_counter()
    step_its, step_rmax = [], []
    for _ in range(args.steps):
        step_its.append(sim.step(DT))
        step_rmax.append(float(sim.solve_info().residual_max))  # host copy of the solve's scalars: no extra device work
    sim.timer_record(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = sim.timer_elapsed_ms(0, 1)
    n_timed = len(sampler.lines) - n_before
    counts = sim.profile_read(reset=True)
    launches = sum(v[2] for v in counts.values())
    # time this rank's exchange kernels spent spinning on a peer's flag (device clock, inside the timed region): the part of the
    # exchange cost that is skew between ranks + NVLink flight, as opposed to this rank's own launches and copies
    exchange_wait = None
    if slab_mode:
        ws = sim.dist_wait_stats(reset=True)
        exchange_wait = {k: {"ms_per_step_mean_over_ranks": round(allsum(1e3 * v[0]) / world / args.steps, 4),
                             "ms_per_step_max_over_ranks": round(allmax(1e3 * v[0]) / args.steps, 4),
                             "kernels_per_step": round(allsum(v[1]) / world / args.steps, 2),
                             "in_kernel_ms_per_step_mean_over_ranks": round(allsum(1e3 * v[2]) / world / args.steps, 4)} for k, v in ws.items()}
    info = sim.solve_info()
    nf = int(info.fluid_cells)
    nf_local = nf // world if slab_mode else nf  # the solve reports the all-rank count
    solver_mode = (os.environ.get("FSIM_SLAB_SOLVER", "hybrid") if slab_mode else "single")
    timings = sim.timings()

    # ---- correctness gate on the state the timed region left behind (VERDICT r1 #1c): every rank contributes its owned planes
    tol = float(scene_params(n, transfer).residual_tolerance)
    wsum = sim.download_grid(abi.FIELD_WSUM).reshape(n, n, gzl if slab_mode else n, 3)[:, :, own_lo - zoff:own_hi - zoff]
    dens = sim.download_grid(abi.FIELD_AVGPNUM).reshape(n, n, gzl if slab_mode else n)[:, :, own_lo - zoff:own_hi - zoff]
    np_after_total = allsum(sim.particle_count())
    np_before_total = allsum(np_local)
    w_tot = [allsum(wsum[..., a].sum(dtype=np.float64)) for a in range(3)]
    d_tot = allsum(dens.sum(dtype=np.float64))
    del wsum, dens
    pou = max(abs(w / np_after_total - 1.0) for w in w_tot + [d_tot])
    rmax = allmax(max(step_rmax))
    checks = {"residual_max": rmax, "residual_tolerance": tol, "residual_ok": bool(rmax < tol),
              "particles_before": int(np_before_total), "particles_after": int(np_after_total),
              "particles_conserved": bool(int(np_before_total) == int(np_after_total)),
              "partition_of_unity_err": pou, "partition_of_unity_ok": bool(pou < 1e-5),
              "what": "max ||r||_inf over the timed steps < residualTolerance; particle count unchanged (no source / sink in the scene); "
                      "sum over owned cells of particleWeightSum (x, y, z faces) and of avgPNum == particle count to 1e-5 relative "
                      "(trilinear weights sum to 1: simulator.cpp:317-333, 362-366)"}
    checks["ok"] = bool(checks["residual_ok"] and checks["particles_conserved"] and checks["partition_of_unity_ok"])
    if verify is not None:
        checks["slab_verify"] = verify
        checks["ok"] = bool(checks["ok"] and verify.get("ok", False))

    # short run: keep the same load up until nvidia-smi has answered a few times (untimed; the decision and the number of
    # extra steps are the same on every rank -- slab steps are collective)
    if allmax(1.0 if len(sampler.lines) < 3 else 0.0) > 0:
        for _ in range(40):
            sim.step(DT)
        sim.synchronize()
        sim.profile_read(reset=True)
    clocks = sampler.stop()
    clocks["samples_in_timed_region"] = n_timed

    total_units, dev_ms_max, value = fdist.aggregate(dist, red_dev, np_local * args.steps, dev_ms, world)
    total_particles = int(round(total_units / args.steps))

    # ---- end-to-end through the C ABI with host buffers: per step set_params + set_obstacles (H2D) + simulate + the
    # manager's gfx export into pinned host memory (D2H 20 B/particle), what simulationThreadWorker does per iteration
    e2e = None
    if not args.no_e2e:
        # two pinned host buffers: the D2H copy of step k overlaps the simulation of step k+1 (the consumer of this data,
        # the render thread, is asynchronous in the reference as well); every buffer is complete before it is reused
        gfx_cap = int(np_local * 1.25) + 1024 if slab_mode else np_local  # slab ranks gain and lose particles
        gfx = [torch.empty((gfx_cap, 5), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        params = scene_params(n, transfer)
        k = max(2, args.steps)  # same step count as the device-resident measurement (pipeline fill and drain are inside the timed region)
        sim.set_params(params); sim.set_obstacles([]); sim.step(DT); sim.export_gfx_async_ptr(gfx[0].data_ptr(), gfx_cap)
        sim.export_gfx_wait()
        barrier()
        t0 = time.perf_counter()
        for i in range(k):
            sim.set_params(params)
            sim.set_obstacles([scenes.cfg3_box(n)] if args.obstacle_box else [])
            sim.step(DT)
            sim.export_gfx_async_ptr(gfx[i % 2].data_ptr(), gfx_cap)
            sim.export_gfx_wait_previous()            # buffer (i-1)%2 has landed and may be consumed
        sim.export_gfx_wait()
        barrier()
        el = allmax(time.perf_counter() - t0)
        e2e = {"value": total_particles * k / el, "unit": UNIT, "steps": k,
               "h2d_bytes_per_step": int(__import__("ctypes").sizeof(abi.Params)),
               "d2h_bytes_per_step": int(np_local * 20), "numa": numa,
               "what": "per step: fsim_set_params + fsim_set_obstacles + fsim_step + fsim_export_gfx_async into pinned host memory "
                       "(double-buffered: the copy of step k overlaps step k+1; all copies complete inside the timed region)"}
        # the same loop with the subsampled export (every 8th particle; an option of the ABI, not the manager's contract)
        sub = 8
        barrier()
        t0 = time.perf_counter()
        lap, laps, sub_its = t0, [], []
        for i in range(k):
            sim.set_params(params)
            sim.set_obstacles([scenes.cfg3_box(n)] if args.obstacle_box else [])
            sub_its.append(sim.step(DT))
            sim.export_gfx_strided_async_ptr(gfx[i % 2].data_ptr(), gfx_cap, sub)
            sim.export_gfx_wait_previous()
            laps.append(round(1e3 * (time.perf_counter() - lap), 2)); lap = time.perf_counter()
        sim.export_gfx_wait()
        barrier()
        el = allmax(time.perf_counter() - t0)
        e2e["subsampled_export"] = {"value": total_particles * k / el, "unit": UNIT, "stride": sub, "d2h_bytes_per_step": int((np_local + sub - 1) // sub * 20),
                                    "host_ms_per_iteration": laps, "pcg_iterations": sub_its,
                                    "what": "same loop, fsim_export_gfx_strided_async with every 8th particle (not the headline: the manager's contract is the full export)"}
        del gfx

    # ---- per-kernel-class durations: CUDA events around every launch on the launching stream.  The timed steps replay the
    # solver iteration as a CUDA graph whose nodes cannot be bracketed one by one, so the scene is REPLAYED on a fresh handle:
    # W warm-up steps, then the first `prof_steps` steps of the timed region again with event brackets (same particles, same
    # step indices => the same PCG iteration counts as the timed steps they repeat)
    barrier()
    sim.close()
    sim = make_sim()
    if world > 1:
        dist.barrier()
    for _ in range(args.warmup):
        sim.step(DT)
    sim.profile_read(reset=True)
    sim.profile_enable(sim.kernel_classes())
    prof_steps = min(2, args.steps)
    prof_its = [sim.step(DT) for _ in range(prof_steps)]
    prof = sim.profile_read(reset=True)
    sim.profile_enable([])
    kernel_ms = {k: {"ms_per_step": round(v[0] / prof_steps, 4), "launches_per_step": v[2] / prof_steps} for k, v in prof.items() if v[2]}
    barrier()
    sim.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    its_mean = float(np.mean(step_its))
    # per-launch algorithmic bytes
    pb = {abi.PIC: 24, abi.FLIP: 24, abi.APIC: 60}[transfer]
    nc = nc_local
    if slab_mode and solver_mode.startswith("r"):  # replicated projection: the solver kernels sweep the whole grid on every rank
        nf_local = nf
    alg = {"advect": np_local * 2 * pb,                      # pass A: read + write pos, vel (+C)
           "p2g": np_local * pb + nc * (8 + 28),             # particles in, bin table, 7 accumulator channels out
           "g2p": np_local * (36 if transfer == abi.FLIP else pb + (12 if transfer == abi.PIC else 48)) + nc * (24 if transfer == abi.FLIP else 12),
           "reorder": np_local * (2 * pb + 8), "bin": np_local * 20,   # sort passes: implementation overhead, listed with their own minimal traffic
           "spmv": nf_local * 16 + nc * 2, "pcg_update": nf_local * 56 + nc * 2, "pcg_direction": nf_local * 20 + nc * 2,
           "mg": nf_local * 15 + nc * 2,                          # average level-0 multigrid kernel (jacobi 14, restrict 10, prolong 14, jacobi+dot 22 B)
           "mg_level1": (nc // 8) * 28, "finalize": nc * 53, "extrapolate": nc * 25, "classify": nc * 5, "rhs": nc * 17 + nf_local * 28}
    if not prof.get("g2p", (0, 0, 0))[2]:  # the G2P ran inside the fused G2P + advect + bin kernel: its grid reads and key/rank writes join that pass
        alg["advect"] += nc * (24 if transfer == abi.FLIP else 12) + np_local * ((36 if transfer == abi.APIC else 0) + 8)
    # roofline of ONE kernel: p2g_kernel, the single longest launch of the step (one launch per step; the kernel VERDICT r1 names)
    dominant = "p2g"
    traffic, traffic_src = None, None
    traffic_path = os.path.join(ROOT, "profiles", "r2_dram_traffic_256_flip.json")
    if n == 256 and transfer == abi.FLIP and world == 1 and os.path.exists(traffic_path):
        tj = json.load(open(traffic_path))
        traffic, traffic_src = tj.get("p2g_kernel"), tj.get("source")
    dom_ms, dom_n, _ = prof[dominant]
    roofline = None
    class_total = sum(v[0] for v in prof.values())
    if dom_n > 0:
        ach = alg[dominant] / (dom_ms / dom_n * 1e-3) / 1e9
        roofline = {"kernel": "p2g_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "avg_launch_ms": dom_ms / dom_n,
                    "launches_timed": dom_n, "algorithmic_bytes_per_launch": alg[dominant],
                    "share_of_step": dom_ms / class_total if class_total > 0 else None,
                    "timed_in": f"{prof_steps} event-bracketed steps replaying the first steps of the timed region on a fresh handle "
                                f"(PCG iterations {prof_its} vs {step_its[:prof_steps]} in the timed steps they repeat)"}
    per_class = {}
    for k_, v in prof.items():  # every class with a traffic model: achieved algorithmic GB/s and fraction of the roof
        if v[1] > 0 and k_ in alg and v[0] > 0:
            a_ = alg[k_] / (v[0] / v[1] * 1e-3) / 1e9
            per_class[k_] = {"avg_launch_ms": round(v[0] / v[1], 5), "algorithmic_GBps": round(a_, 1), "frac": round(a_ / peak, 4)}
    # whole-step fraction of the HBM roof: with the survey's b_it (136 B per fluid cell and iteration) and with
    # the traffic this implementation's fp64 CG vectors add (174 B)
    gpus = world if slab_mode else 1
    def step_frac_of(b_it):
        b = (total_particles if slab_mode else np_local) * B_PARTICLE[transfer] + n ** 3 * B_CELL + its_mean * nf * b_it
        return b, b / (dev_ms_max / args.steps * 1e-3) / 1e9 / (peak * gpus)
    b_step, step_frac = step_frac_of(B_IT_SURVEY)
    b_step_impl, step_frac_impl = step_frac_of(B_IT)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak" if (world > 1 and not slab_mode) else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"3D {args.transfer} dam break {n}^3 (SURVEY §8d cfg 3 headline variant), "
                                   + (f"{total_particles} particles in ONE domain cut into {world} z-slabs, " if slab_mode else f"{np_local} particles per GPU, ")
                                   + f"8 per fluid cell, dt {DT}, PCG tol 1e-6", "grid": [n, n, n], "particles_per_gpu": np_local,
                       "particles_total": total_particles, "fluid_cells": nf, "pcg_iterations_mean": its_mean,
                       "slab_solver": solver_mode,
                       "parallelism": (f"z-slabs x{world} (ghost planes, particle migration and PCG reductions over NVLink peer memory; projection: {solver_mode})" if slab_mode
                                       else (f"replicas x{world}" if world > 1 else "single GPU")),
                       "l2": "inputs larger than L2 (particle + grid state >> 126 MB), no flush needed"},
            "wall_ms_per_step": 1e3 * wall / args.steps, "clocks": clocks, "gpu_launches": int(launches),
            "checks": checks,
            "step_hbm_frac": step_frac, "step_algorithmic_bytes": b_step, "b_it": B_IT_SURVEY,
            "step_hbm_frac_impl_b_it": step_frac_impl, "b_it_impl": B_IT,
            "stage_us": {k: v for k, v in zip(["advect", "", "", "p2g", "classify", "project", "extrapolate", "g2p"], list(timings.last_raw_us)) if k},
            "sort_us": timings.last_sort_us, "exchange_wait": exchange_wait, "kernel_ms": kernel_ms, "kernel_roofline": per_class, "roofline": roofline, "e2e": e2e, "setup_s": t_gen}
    if not args.no_cpu_baseline and world == 1:  # (the contract asks for it on rank 0 at N = 1 only)
        try:
            r = run_reference(args, args.cpu_grid if args.cpu_grid else 96, args.cpu_steps, args.cpu_warmup)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"error": str(ex)}
    emit(line)
    if world > 1:
        if os.environ.get("NCCL_DEBUG_FILE"):  # the communicator lines of NCCL's log, for whoever reads this run's stderr
            import glob
            for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "nccl_*.log")))[:world]:
                for ln in open(path, errors="replace"):
                    if "Init COMPLETE" in ln or "nranks" in ln:
                        print(ln.rstrip(), file=sys.stderr)
        dist.destroy_process_group()
    if not checks["ok"]:
        raise SystemExit("bench.py: correctness gate failed: " + json.dumps(checks))


if __name__ == "__main__":
    main()
