"""Development tool: fits project-stage time = F + c * iterations by capping the PCG iteration count (runs on the GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fluid_simulator_b200 import abi, scenes
from fluid_simulator_b200.sim import FluidSim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pos = scenes.block_positions_f32(1, n // 2, 1, n - 1, 1, n - 1)
sim = FluidSim((float(n),) * 3, 1.0, False, 0.25, capacity=pos.shape[0])
sim.set_id_tracking(False)
sim.set_params(scenes.default_params(abi.FLIP))
sim.upload_particles_f32(pos)
for _ in range(4):
    sim.step(0.005)
# now time the projection alone, repeatedly, on the same grid state (warm start makes later solves trivial -> disable by resetting)
for cap in (1, 2, 4, 8, 16, 1, 2, 4, 8):
    sim.set_params(scenes.default_params(abi.FLIP, max_iterations=cap))
    sim.stage_p2g(); sim.stage_classify(0.005)
    sim.synchronize()
    t0 = time.perf_counter()
    sim.timer_record(0)
    its = sim.stage_project(0.005)
    sim.timer_record(1)
    sim.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"cap {cap:3d} its {its:3d} device {sim.timer_elapsed_ms(0, 1):7.3f} ms wall {wall:7.3f} ms rmax {sim.solve_info().residual_max:.3e}")
