"""TEST INFRASTRUCTURE ONLY -- not part of the product.

ctypes wrapper of oracle/_ref/libfsim_ref{,_serial}.so: the UNMODIFIED reference Simulator compiled
by oracle/build_ref.sh from /root/reference/src/Simulator.  Same method names as the product
binding (fluid_simulator_b200.sim.FluidSim) so parity tests read symmetrically.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm may import this.
"""
import ctypes as C
import os

import numpy as np

from fluid_simulator_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib_path(serial=False):
    return os.path.join(_HERE, "_ref", "libfsim_ref_serial.so" if serial else "libfsim_ref.so")


def available(serial=False):
    return os.path.exists(lib_path(serial))


def _load(serial):
    if serial in _LIBS:
        return _LIBS[serial]
    L = C.CDLL(lib_path(serial), mode=os.RTLD_LOCAL | os.RTLD_NOW)
    vp, dbl, i32, i64 = C.c_void_p, C.c_double, C.c_int, C.c_int64
    sigs = {
        "ref_create": (vp, [C.POINTER(abi.GridDesc)]),
        "ref_destroy": (None, [vp]),
        "ref_get_grid_info": (None, [vp, C.POINTER(abi.GridInfo)]),
        "ref_set_params": (None, [vp, C.POINTER(abi.Params)]),
        "ref_set_obstacles": (None, [vp, C.POINTER(abi.Obstacle), i32]),
        "ref_get_obstacles": (None, [vp, C.POINTER(abi.Obstacle), i32]),
        "ref_set_particles": (None, [vp, vp, i64]),
        "ref_particle_count": (i64, [vp]),
        "ref_get_particles": (None, [vp, vp, i64]),
        "ref_get_particle_cells": (None, [vp, vp, i64]),
        "ref_step": (i32, [vp, dbl]),
        "ref_last_step_seconds": (dbl, [vp]),
        "ref_stage_spawn": (None, [vp, dbl]),
        "ref_stage_advect": (None, [vp, dbl]),
        "ref_stage_push_apart": (None, [vp]),
        "ref_stage_push_out": (None, [vp]),
        "ref_stage_p2g": (None, [vp, dbl]),
        "ref_stage_classify": (None, [vp, dbl]),
        "ref_stage_project": (i32, [vp, dbl]),
        "ref_stage_extrapolate": (None, [vp]),
        "ref_stage_g2p": (None, [vp]),
        "ref_get_grid": (i32, [vp, i32, vp]),
        "ref_set_grid": (i32, [vp, i32, vp]),
        "ref_post_p2g_update": (None, [vp, dbl]),
        "ref_export_gfx": (None, [vp, vp, i64]),
        "ref_get_step_durations": (None, [vp, C.POINTER(abi.Timings)]),
        "ref_srand": (None, [C.c_uint]),
        "ref_omp_threads": (i32, []),
        "ref_set_omp_threads": (None, [i32]),
        "ref_is_serial_build": (i32, []),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIBS[serial] = L
    return L


class RefSim:
    """The reference's BridsonSolverGrid + HashedParticles + Simulator, driven headless."""

    def __init__(self, dims, resolution=1.0, two_d=False, particle_radius=0.25, serial=False, basic_solver=False, **_):
        self.L = _load(serial)
        d = abi.GridDesc()
        d.target_dims[:] = dims
        d.resolution = resolution
        d.two_d = int(two_d)
        d.particle_radius = particle_radius
        d.reserved0 = 1 if basic_solver else 0  # BasicMacGrid instead of BridsonSolverGrid
        self.h = self.L.ref_create(C.byref(d))
        self.info = abi.GridInfo()
        self.L.ref_get_grid_info(self.h, C.byref(self.info))
        self.grid_size = tuple(self.info.grid_size)
        self.nc = int(self.info.cell_count)
        self._nobs = 0

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- configuration -------------------------------------------------------------------
    def set_params(self, params):
        self.L.ref_set_params(self.h, C.byref(params))

    def set_obstacles(self, obstacles):
        arr = (abi.Obstacle * max(1, len(obstacles)))(*obstacles)
        self._nobs = len(obstacles)
        self.L.ref_set_obstacles(self.h, arr, len(obstacles))

    def get_obstacles(self, obstacles):
        arr = (abi.Obstacle * max(1, len(obstacles)))(*obstacles)
        self.L.ref_get_obstacles(self.h, arr, len(obstacles))
        return list(arr)[:len(obstacles)]

    # --- particles -----------------------------------------------------------------------
    def upload_particles(self, aos15):
        a = np.ascontiguousarray(aos15, dtype=np.float64).reshape(-1, 15)
        self.L.ref_set_particles(self.h, a.ctypes.data, a.shape[0])

    def particle_count(self):
        return int(self.L.ref_particle_count(self.h))

    def download_particles(self):
        n = self.particle_count()
        out = np.empty((n, 15), dtype=np.float64)
        self.L.ref_get_particles(self.h, out.ctypes.data, n)
        return out

    def download_particle_cells(self):
        n = self.particle_count()
        out = np.empty(n, dtype=np.int32)
        self.L.ref_get_particle_cells(self.h, out.ctypes.data, n)
        return out

    # --- stepping ------------------------------------------------------------------------
    def step(self, dt):
        return int(self.L.ref_step(self.h, dt))

    def last_step_seconds(self):
        return float(self.L.ref_last_step_seconds(self.h))

    def stage_spawn(self, dt): self.L.ref_stage_spawn(self.h, dt)
    def stage_advect(self, dt): self.L.ref_stage_advect(self.h, dt)
    def stage_push_apart(self): self.L.ref_stage_push_apart(self.h)
    def stage_push_out(self): self.L.ref_stage_push_out(self.h)
    def stage_p2g(self, dt=0.0): self.L.ref_stage_p2g(self.h, dt)
    def stage_classify(self, dt): self.L.ref_stage_classify(self.h, dt)
    def stage_project(self, dt): return int(self.L.ref_stage_project(self.h, dt))
    def stage_extrapolate(self): self.L.ref_stage_extrapolate(self.h)
    def stage_g2p(self): self.L.ref_stage_g2p(self.h)
    def post_p2g_update(self, gravity_increment): self.L.ref_post_p2g_update(self.h, gravity_increment)

    # --- grid ----------------------------------------------------------------------------
    def download_grid(self, field):
        shape, dt = abi.field_shape_dtype(field, self.nc)
        out = np.zeros(shape, dtype=dt)
        self.L.ref_get_grid(self.h, field, out.ctypes.data)
        return out

    def upload_grid(self, field, arr):
        shape, dt = abi.field_shape_dtype(field, self.nc)
        a = np.ascontiguousarray(arr, dtype=dt).reshape(shape)
        self.L.ref_set_grid(self.h, field, a.ctypes.data)

    def export_gfx(self):
        n = self.particle_count()
        out = np.zeros((n, 5), dtype=np.float32)
        self.L.ref_export_gfx(self.h, out.ctypes.data, n)
        return out

    def step_durations(self):
        t = abi.Timings()
        self.L.ref_get_step_durations(self.h, C.byref(t))
        return t.as_map()

    def srand(self, seed): self.L.ref_srand(seed)
    def omp_threads(self): return int(self.L.ref_omp_threads())
    def set_omp_threads(self, n): self.L.ref_set_omp_threads(int(n))
