"""SURVEY.md §8d cfg 5: projection-only sweep.  For N in the list: dam block WATER, v = 0, postP2GUpdate(g*dt) (hydrostatic
start), one pressure solve to ||r||_inf < 1e-6.  Prints one JSON line per N with iterations, ms/solve, ms/iteration and
the residual (divergence) after the solve.  Runs on the GPU box:  python tools/bench_projection.py 128 192 256 384 512"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fluid_simulator_b200 import abi, scenes
from fluid_simulator_b200.sim import FluidSim

os.environ["FSIM_NO_WARM_START"] = "1"   # every repetition is a cold solve (p = 0 start), like the reference's

for n in [int(x) for x in sys.argv[1:]] or [128, 256]:
    sim = FluidSim((float(n),) * 3, 1.0, False, 0.25)
    sim.set_params(scenes.default_params(abi.FLIP, pressure_enabled=False, max_iterations=2000, tol=1e-6))
    types = scenes.hydrostatic_types(n)
    res = []
    for rep in range(3):
        sim.upload_grid(abi.FIELD_TYPE, types)
        sim.upload_grid(abi.FIELD_V, np.zeros((n ** 3, 3)))
        sim.post_p2g_update(-39.24 * 0.005)
        sim.synchronize()
        sim.timer_record(0)
        its = sim.stage_project(0.005)
        sim.timer_record(1)
        sim.synchronize()
        res.append((its, sim.timer_elapsed_ms(0, 1)))
    info = sim.solve_info()
    its, ms = res[-1]   # repetition 0 pays the one-time hierarchy allocation and graph capture
    nf = int(info.fluid_cells)
    print(json.dumps({"cfg": 5, "grid": n, "fluid_cells": nf, "iterations": its, "ms_per_solve": ms,
                      "ms_per_iteration": ms / max(its, 1), "first_solve_ms_incl_setup": res[0][1],
                      "algorithmic_GBps": (its * nf * 174 + n ** 3 * 60) / (ms * 1e-3) / 1e9,
                      "residual_max": info.residual_max, "reference_mic0_iterations": {64: 41, 128: 80, 256: 155}.get(n)}))
    sim.close()
