"""Runs a small slab-vs-single comparison under several library switches, each in its own process (GPU box helper)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from fluid_simulator_b200 import abi, scenes
from fluid_simulator_b200.sim import FluidSim
from fluid_simulator_b200.slab import SlabGroup
from util import rel_l2
n = int(sys.argv[1]); ranks = int(sys.argv[2]); steps = int(sys.argv[3])
sc = scenes.dam_break_3d(n, abi.FLIP, tol=1e-9)
one = FluidSim(sc.dims, sc.resolution, sc.two_d, sc.particle_radius)
grp = SlabGroup(ranks, sc.dims, sc.resolution, sc.two_d, sc.particle_radius, capacity=sc.n_particles)
for s in (one, grp):
    s.set_params(sc.params); s.set_obstacles([]); s.upload_particles(sc.particles)
for st in range(steps):
    i1 = one.step(sc.dt); ig = grp.step(sc.dt)
    grp.synchronize()
    t = int((grp.download_grid(abi.FIELD_TYPE) != one.download_grid(abi.FIELD_TYPE)).sum())
    print('step', st, 'its', i1, ig, 'type_mismatch', t, 'v', rel_l2(grp.download_grid(abi.FIELD_V), one.download_grid(abi.FIELD_V)),
          'v2', rel_l2(grp.download_grid(abi.FIELD_V2), one.download_grid(abi.FIELD_V2)),
          'p', rel_l2(grp.download_grid(abi.FIELD_PRESSURE), one.download_grid(abi.FIELD_PRESSURE)), 'counts', grp.particle_counts(), flush=True)
pa, pb = grp.download_particles(), one.download_particles()
print('particles', pa.shape, pb.shape, rel_l2(pa[:, :6], pb[:, :6]))
print('VARIANT_OK')
""" % (ROOT, ROOT)

VARIANTS = [
    ("eager_nograph_cluster1", {"CUDA_MODULE_LOADING": "EAGER", "FSIM_NO_GRAPH": "1", "FSIM_MG_TAIL_CLUSTER": "1"}),
    ("eager_default", {"CUDA_MODULE_LOADING": "EAGER"}),
    ("nograph_cluster1", {"FSIM_NO_GRAPH": "1", "FSIM_MG_TAIL_CLUSTER": "1"}),
    ("graph_cluster1", {"FSIM_MG_TAIL_CLUSTER": "1"}),
    ("default", {}),
]

if __name__ == "__main__":
    n, ranks, steps = (sys.argv[1:4] + ["16", "2", "2"][len(sys.argv) - 1:])[:3]
    for name, env in VARIANTS:
        e = dict(os.environ, FSIM_DIST_TIMEOUT_MS="5000", **env)
        try:
            r = subprocess.run([sys.executable, "-c", CHILD, n, ranks, steps], env=e, capture_output=True, text=True, timeout=120)
            out = (r.stdout + "\n" + r.stderr[-700:]).strip()
        except subprocess.TimeoutExpired as ex:
            out = "TIMEOUT " + str(ex.stdout)[-1500:]
        print(f"===== {name} =====\n{out}\n", flush=True)
