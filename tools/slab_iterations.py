"""PCG iteration counts of the slab decomposition vs the single handle (all ranks on cuda:0; iteration counts do not
depend on where the ranks run).  usage: slab_iterations.py N steps ranks[,ranks...]"""
import json
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fluid_simulator_b200 import abi, scenes  # noqa: E402
from fluid_simulator_b200.sim import FluidSim  # noqa: E402
from fluid_simulator_b200.slab import SlabGroup  # noqa: E402

n, steps = int(sys.argv[1]), int(sys.argv[2])
ranks = [int(x) for x in sys.argv[3].split(",")]
sc = scenes.dam_break_3d(n, abi.FLIP, tol=1e-6)
one = FluidSim(sc.dims, sc.resolution, sc.two_d, sc.particle_radius)
one.set_params(sc.params); one.upload_particles(sc.particles)
base = [one.step(sc.dt) for _ in range(steps)]
one.close()
print(json.dumps({"grid": n, "ranks": 1, "its": base}), flush=True)
for r in ranks:
    for cut in ("hybrid", "replicated"):
        os.environ["FSIM_SLAB_SOLVER"] = cut
        g = SlabGroup(r, sc.dims, sc.resolution, sc.two_d, sc.particle_radius, capacity=sc.n_particles)
        g.set_params(sc.params); g.set_obstacles([]); g.upload_particles(sc.particles)
        its = [g.step(sc.dt) for _ in range(steps)]
        ms = [s.last_step_stats()[0] for s in g.sims]
        print(json.dumps({"grid": n, "ranks": r, "cut": cut, "its": its, "last_step_ms_per_rank": ms}), flush=True)
        g.close()
