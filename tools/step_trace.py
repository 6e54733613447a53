"""Per-step trace of the bench scene (3D FLIP dam break, bench.py recipe): wall ms, PCG iterations, residual, stage times.
    python tools/step_trace.py [grid] [steps] > trace.jsonl
Used to look at how the cost of a step develops as the dam break evolves (the bench times steps W .. W+K of the scene)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from fluid_simulator_b200 import abi, scenes  # noqa: E402
from fluid_simulator_b200.sim import FluidSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 80
pos = scenes.block_positions_f32(1, n // 2, 1, n - 1, 1, n - 1, seed=scenes.SEED)
sim = FluidSim((float(n),) * 3, 1.0, False, 0.25, capacity=pos.shape[0])
sim.set_id_tracking(False)
sim.set_params(scenes.default_params(abi.FLIP, max_iterations=2000, tol=1e-6))
sim.upload_particles_f32(pos)
for st in range(steps):
    sim.synchronize()
    t0 = time.perf_counter()
    its = sim.step(0.005)
    sim.synchronize()
    ms = 1e3 * (time.perf_counter() - t0)
    info = sim.solve_info()
    d = sim.step_durations()
    print(json.dumps({"step": st, "ms": round(ms, 3), "its": its, "rmax": float(info.residual_max), "fluid_cells": int(info.fluid_cells),
                      "stages_us": {k: round(float(v), 1) for k, v in d.items()} if isinstance(d, dict) else None}), flush=True)
