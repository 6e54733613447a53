"""Per-step wall-clock trace of bench.py's end-to-end loop (step + async gfx export into pinned host memory)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fluid_simulator_b200 import abi, scenes
from fluid_simulator_b200.sim import FluidSim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
pos = scenes.block_positions_f32(1, n // 2, 1, n - 1, 1, n - 1)
npart = pos.shape[0]
sim = FluidSim((float(n),) * 3, 1.0, False, 0.25, capacity=npart)
params = scenes.default_params(abi.FLIP)
sim.set_params(params)
sim.upload_particles_f32(pos)
gfx = [torch.empty((npart, 5), dtype=torch.float32, pin_memory=True) for _ in range(2)]
for _ in range(3):
    sim.step(0.005)
sim.export_gfx_async_ptr(gfx[0].data_ptr(), npart); sim.export_gfx_wait()
t = [time.perf_counter()]
marks = []
for i in range(steps):
    a = time.perf_counter()
    sim.set_params(params); sim.set_obstacles([])
    sim.step(0.005)
    b = time.perf_counter()
    sim.export_gfx_async_ptr(gfx[i % 2].data_ptr(), npart)
    c = time.perf_counter()
    sim.export_gfx_wait_previous()
    d = time.perf_counter()
    marks.append((b - a, c - b, d - c))
    t.append(d)
sim.export_gfx_wait()
t.append(time.perf_counter())
print("per-step wall ms:", [round(1e3 * (t[i + 1] - t[i]), 2) for i in range(len(t) - 1)])
print("step / enqueue export / wait previous (ms):", [tuple(round(1e3 * x, 2) for x in m) for m in marks])
print("total ms/step:", round(1e3 * (t[-1] - t[0]) / steps, 2), " D2H bytes/step:", npart * 20)
