"""TEST INFRASTRUCTURE ONLY -- not part of the product.

ctypes wrapper of oracle/libfsim_oracle.so (fsim_oracle.c: the plain-C fp64 restatement of the
reference hot path).  Same method names as RefSim (oracle/refsim.py) and as the product binding
FluidSim (fluid_simulator_b200/sim.py).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from fluid_simulator_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib_path():
    return os.path.join(_HERE, "libfsim_oracle.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, os.path.join(_HERE, "libfsim_oracle.so")], stdout=subprocess.DEVNULL)


def _load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(lib_path()):
        build()
    L = C.CDLL(lib_path(), mode=os.RTLD_LOCAL | os.RTLD_NOW)
    vp, dbl, i32, i64 = C.c_void_p, C.c_double, C.c_int, C.c_int64
    sigs = {
        "oracle_create": (vp, [C.POINTER(abi.GridDesc)]),
        "oracle_destroy": (None, [vp]),
        "oracle_get_grid_info": (None, [vp, C.POINTER(abi.GridInfo)]),
        "oracle_set_params": (None, [vp, C.POINTER(abi.Params)]),
        "oracle_set_obstacles": (None, [vp, C.POINTER(abi.Obstacle), i32]),
        "oracle_get_obstacles": (None, [vp, C.POINTER(abi.Obstacle), i32]),
        "oracle_set_particles": (None, [vp, vp, i64]),
        "oracle_append_particles": (None, [vp, vp, i64]),
        "oracle_remove_particles": (None, [vp, vp, i64]),
        "oracle_particle_count": (i64, [vp]),
        "oracle_get_particles": (None, [vp, vp, i64]),
        "oracle_get_particle_cells": (None, [vp, vp, i64]),
        "oracle_step": (i32, [vp, dbl]),
        "oracle_stage_spawn": (None, [vp, dbl]),
        "oracle_stage_advect": (None, [vp, dbl]),
        "oracle_stage_push_out": (None, [vp]),
        "oracle_stage_stop": (None, [vp]),
        "oracle_stage_p2g": (None, [vp]),
        "oracle_stage_classify": (None, [vp, dbl]),
        "oracle_stage_project": (i32, [vp, dbl]),
        "oracle_stage_extrapolate": (None, [vp]),
        "oracle_stage_g2p": (None, [vp]),
        "oracle_post_p2g_update": (None, [vp, dbl]),
        "oracle_get_grid": (i32, [vp, i32, vp]),
        "oracle_set_grid": (i32, [vp, i32, vp]),
        "oracle_export_gfx": (None, [vp, vp, i64]),
        "oracle_get_solve_info": (None, [vp, C.POINTER(abi.SolveInfo)]),
        "oracle_srand": (None, [C.c_uint]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L


class OracleSim:
    """fp64 CPU restatement of BridsonSolverGrid + HashedParticles + Simulator."""

    def __init__(self, dims, resolution=1.0, two_d=False, particle_radius=0.25, **_):
        self.L = _load()
        d = abi.GridDesc()
        d.target_dims[:] = dims
        d.resolution = resolution
        d.two_d = int(two_d)
        d.particle_radius = particle_radius
        self.h = self.L.oracle_create(C.byref(d))
        self.info = abi.GridInfo()
        self.L.oracle_get_grid_info(self.h, C.byref(self.info))
        self.grid_size = tuple(self.info.grid_size)
        self.nc = int(self.info.cell_count)

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params): self.L.oracle_set_params(self.h, C.byref(params))

    def set_obstacles(self, obstacles):
        arr = (abi.Obstacle * max(1, len(obstacles)))(*obstacles)
        self.L.oracle_set_obstacles(self.h, arr, len(obstacles))

    def get_obstacles(self, obstacles):
        arr = (abi.Obstacle * max(1, len(obstacles)))(*obstacles)
        self.L.oracle_get_obstacles(self.h, arr, len(obstacles))
        return list(arr)[:len(obstacles)]

    def upload_particles(self, aos15):
        a = np.ascontiguousarray(aos15, dtype=np.float64).reshape(-1, 15)
        self.L.oracle_set_particles(self.h, a.ctypes.data, a.shape[0])

    def append_particles(self, aos15):
        a = np.ascontiguousarray(aos15, dtype=np.float64).reshape(-1, 15)
        self.L.oracle_append_particles(self.h, a.ctypes.data, a.shape[0])

    def remove_particles(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int32).copy()
        self.L.oracle_remove_particles(self.h, a.ctypes.data, a.shape[0])

    def particle_count(self): return int(self.L.oracle_particle_count(self.h))

    def download_particles(self):
        n = self.particle_count()
        out = np.empty((n, 15), dtype=np.float64)
        self.L.oracle_get_particles(self.h, out.ctypes.data, n)
        return out

    def download_particle_cells(self):
        n = self.particle_count()
        out = np.empty(n, dtype=np.int32)
        self.L.oracle_get_particle_cells(self.h, out.ctypes.data, n)
        return out

    def step(self, dt): return int(self.L.oracle_step(self.h, dt))
    def stage_spawn(self, dt): self.L.oracle_stage_spawn(self.h, dt)
    def stage_advect(self, dt): self.L.oracle_stage_advect(self.h, dt)
    def stage_push_out(self): self.L.oracle_stage_push_out(self.h)
    def stage_p2g(self, dt=0.0): self.L.oracle_stage_p2g(self.h)
    def stage_classify(self, dt): self.L.oracle_stage_classify(self.h, dt)
    def stage_project(self, dt): return int(self.L.oracle_stage_project(self.h, dt))
    def stage_extrapolate(self): self.L.oracle_stage_extrapolate(self.h)
    def stage_g2p(self): self.L.oracle_stage_g2p(self.h)
    def post_p2g_update(self, gravity_increment): self.L.oracle_post_p2g_update(self.h, gravity_increment)

    def download_grid(self, field):
        shape, dt = abi.field_shape_dtype(field, self.nc)
        out = np.zeros(shape, dtype=dt)
        self.L.oracle_get_grid(self.h, field, out.ctypes.data)
        return out

    def upload_grid(self, field, arr):
        shape, dt = abi.field_shape_dtype(field, self.nc)
        a = np.ascontiguousarray(arr, dtype=dt).reshape(shape)
        self.L.oracle_set_grid(self.h, field, a.ctypes.data)

    def export_gfx(self):
        n = self.particle_count()
        out = np.zeros((n, 5), dtype=np.float32)
        self.L.oracle_export_gfx(self.h, out.ctypes.data, n)
        return out

    def solve_info(self):
        s = abi.SolveInfo()
        self.L.oracle_get_solve_info(self.h, C.byref(s))
        return s

    def srand(self, seed): self.L.oracle_srand(seed)
