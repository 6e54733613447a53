"""Python binding of libfsim_b200.so (include/fsim.h) through ctypes.

`FluidSim` mirrors the interface of the reference's Simulator + MacGrid + HashedParticles as the
SimulationManager drives them (set config / obstacles, simulate(dt), read particles and cells), with
the same method names as the test-only wrappers of the compiled reference (oracle/refsim.py) and of the
oracle restatement (oracle/oracle.py).  There is no fallback: if the CUDA library is missing or no
sm_100 device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfsim_b200.so")
_LIB = None


class FsimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fsim error {code}: {msg}")
        self.code = code


def exported_symbols():
    """Names declared in include/fsim.h (parsed from the header)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "fsim.h")
    return re.findall(r"FSIM_API\s+[\w\s\*]+?\b(fsim_\w+)\s*\(", open(hdr).read())


def load_library():
    """Loads libfsim_b200.so; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise FsimError(abi.ERR_CUDA, f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                      "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    vp, dbl, i32, i64 = C.c_void_p, C.c_double, C.c_int, C.c_int64
    P = C.POINTER
    sigs = {
        "fsim_abi_version": (i32, []),
        "fsim_last_error": (C.c_char_p, [vp]),
        "fsim_create": (i32, [P(abi.GridDesc), P(vp)]),
        "fsim_destroy": (i32, [vp]),
        "fsim_get_grid_info": (i32, [vp, P(abi.GridInfo)]),
        "fsim_create_slab": (i32, [P(abi.GridDesc), i32, i32, P(vp)]),
        "fsim_get_slab_info": (i32, [vp, P(abi.SlabInfo)]),
        "fsim_dist_export": (i32, [vp, P(abi.DistExport)]),
        "fsim_dist_connect": (i32, [vp, P(abi.DistExport), i32]),
        "fsim_dist_wait_stats": (i32, [vp, P(abi.DistWaitStats), i32]),
        "fsim_slab_partition": (i32, [i32, i32, i32, P(abi.SlabInfo)]),
        "fsim_upload_particle_ids": (i32, [vp, vp, i64]),
        "fsim_set_params": (i32, [vp, P(abi.Params)]),
        "fsim_set_obstacles": (i32, [vp, P(abi.Obstacle), i32]),
        "fsim_get_obstacles": (i32, [vp, P(abi.Obstacle), i32, P(i32)]),
        "fsim_set_particle_radius": (i32, [vp, dbl]),
        "fsim_upload_particles": (i32, [vp, vp, i64]),
        "fsim_append_particles": (i32, [vp, vp, i64]),
        "fsim_remove_particles": (i32, [vp, vp, i64]),
        "fsim_particle_count": (i32, [vp, P(i64)]),
        "fsim_download_particles": (i32, [vp, vp, i64, P(i64)]),
        "fsim_set_id_tracking": (i32, [vp, i32]),
        "fsim_download_particle_ids": (i32, [vp, vp, i64]),
        "fsim_upload_particles_f32": (i32, [vp, vp, vp, vp, i64]),
        "fsim_download_particles_f32": (i32, [vp, vp, vp, vp, i64, P(i64)]),
        "fsim_step": (i32, [vp, dbl, P(i32)]),
        "fsim_stage_spawn": (i32, [vp, dbl]),
        "fsim_stage_advect": (i32, [vp, dbl]),
        "fsim_stage_push_apart": (i32, [vp]),
        "fsim_stage_push_out": (i32, [vp]),
        "fsim_stage_p2g": (i32, [vp]),
        "fsim_stage_classify": (i32, [vp, dbl]),
        "fsim_stage_post_p2g_update": (i32, [vp, dbl]),
        "fsim_stage_project": (i32, [vp, dbl, P(i32)]),
        "fsim_stage_extrapolate": (i32, [vp]),
        "fsim_stage_g2p": (i32, [vp]),
        "fsim_download_grid": (i32, [vp, i32, vp, i64]),
        "fsim_upload_grid": (i32, [vp, i32, vp, i64]),
        "fsim_download_particle_cells": (i32, [vp, vp, i64]),
        "fsim_export_gfx": (i32, [vp, vp, i64, P(i64)]),
        "fsim_export_gfx_async": (i32, [vp, vp, i64, P(i64)]),
        "fsim_export_gfx_strided_async": (i32, [vp, vp, i64, i64, P(i64)]),
        "fsim_export_gfx_wait": (i32, [vp]),
        "fsim_export_gfx_wait_previous": (i32, [vp]),
        "fsim_get_step_durations": (i32, [vp, P(abi.Timings)]),
        "fsim_get_solve_info": (i32, [vp, P(abi.SolveInfo)]),
        "fsim_get_last_step_stats": (i32, [vp, P(dbl), P(i64)]),
        "fsim_synchronize": (i32, [vp]),
        "fsim_kernel_class_count": (i32, []),
        "fsim_kernel_class_name": (C.c_char_p, [i32]),
        "fsim_profile_enable": (i32, [vp, C.c_uint32]),
        "fsim_profile_read": (i32, [vp, vp, vp, vp, i32]),
        "fsim_timer_record": (i32, [vp, i32]),
        "fsim_timer_elapsed_ms": (i32, [vp, i32, i32, P(dbl)]),
    }
    for name, (res, args) in sigs.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    if L.fsim_abi_version() != abi.ABI_VERSION:
        raise FsimError(abi.ERR_INVALID, "ABI version mismatch between libfsim_b200.so and fluid_simulator_b200.abi")
    _LIB = L
    return L


class FluidSim:
    """B200 Simulator: BridsonSolverGrid + HashedParticles + Simulator behind the fsim C ABI."""

    def __init__(self, dims, resolution=1.0, two_d=False, particle_radius=0.25, capacity=0, device=0, rank=0, nranks=1, **_):
        self.L = load_library()
        d = abi.GridDesc()
        d.target_dims[:] = dims
        d.resolution = resolution
        d.two_d = int(two_d)
        d.particle_radius = particle_radius
        d.particle_capacity = int(capacity)
        d.device = int(device)
        self.h = C.c_void_p()
        if nranks > 1:  # one z-slab of the grid (fluid_simulator_b200.slab drives the group)
            rc = self.L.fsim_create_slab(C.byref(d), int(rank), int(nranks), C.byref(self.h))
        else:
            rc = self.L.fsim_create(C.byref(d), C.byref(self.h))
        if rc:
            msg = self.L.fsim_last_error(None).decode()
            self.h = None
            raise FsimError(rc, msg)
        self.info = abi.GridInfo()
        self._ck(self.L.fsim_get_grid_info(self.h, C.byref(self.info)))
        self.grid_size = tuple(self.info.grid_size)
        self.nc = int(self.info.cell_count)
        self.slab = abi.SlabInfo()
        self._ck(self.L.fsim_get_slab_info(self.h, C.byref(self.slab)))
        if nranks > 1:  # grid transfers move the local block: the owned planes + one ghost plane per neighbour
            self.local_grid_size = (self.grid_size[0], self.grid_size[1], int(self.slab.gz_local))
            self.nc = self.local_grid_size[0] * self.local_grid_size[1] * self.local_grid_size[2]
        else:
            self.local_grid_size = self.grid_size

    def _ck(self, rc):
        if rc:
            raise FsimError(rc, self.L.fsim_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.fsim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- z-slab group membership (include/fsim.h "z-slab decomposition") ----------------------------
    def dist_export(self):
        ex = abi.DistExport()
        self._ck(self.L.fsim_dist_export(self.h, C.byref(ex)))
        return ex

    def dist_connect(self, exports):
        arr = (abi.DistExport * len(exports))(*exports)
        self._ck(self.L.fsim_dist_connect(self.h, arr, len(exports)))

    def dist_wait_stats(self, reset=False):
        """{class: (seconds one thread per kernel spent waiting on peers' flags, number of such kernels, seconds inside the
        signalling block of the push / all-reduce kernels)} since connect / the last reset (fsim_dist_wait_stats)"""
        st = abi.DistWaitStats()
        self._ck(self.L.fsim_dist_wait_stats(self.h, C.byref(st), 1 if reset else 0))
        return {k: (st.wait_ns[i] * 1e-9, int(st.waits[i]), st.kernel_ns[i] * 1e-9) for i, k in enumerate(abi.WAIT_CLASSES)}

    def upload_particle_ids(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.uint32)
        self._ck(self.L.fsim_upload_particle_ids(self.h, a.ctypes.data, a.shape[0]))

    # --- configuration (simulator->config = ..., simulator->obstacles = ...) ---------------------
    def set_params(self, params):
        self._ck(self.L.fsim_set_params(self.h, C.byref(params)))

    def set_obstacles(self, obstacles):
        arr = (abi.Obstacle * max(1, len(obstacles)))(*obstacles)
        self._ck(self.L.fsim_set_obstacles(self.h, arr, len(obstacles)))

    def get_obstacles(self, obstacles=None):
        arr = (abi.Obstacle * abi.MAX_OBSTACLES)()
        n = C.c_int32()
        self._ck(self.L.fsim_get_obstacles(self.h, arr, abi.MAX_OBSTACLES, C.byref(n)))
        return list(arr)[:n.value]

    # --- particles ------------------------------------------------------------------------------
    def upload_particles(self, aos15):
        a = np.ascontiguousarray(aos15, dtype=np.float64).reshape(-1, 15)
        self._ck(self.L.fsim_upload_particles(self.h, a.ctypes.data, a.shape[0]))

    def append_particles(self, aos15):
        a = np.ascontiguousarray(aos15, dtype=np.float64).reshape(-1, 15)
        self._ck(self.L.fsim_append_particles(self.h, a.ctypes.data, a.shape[0]))

    def remove_particles(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        self._ck(self.L.fsim_remove_particles(self.h, a.ctypes.data, a.shape[0]))

    def particle_count(self):
        n = C.c_int64()
        self._ck(self.L.fsim_particle_count(self.h, C.byref(n)))
        return int(n.value)

    def download_particles(self, by_id=True):
        """Particles in the reference `Particle` layout.  The device keeps them cell-binned; with by_id (default)
        they are returned in persistent-id order so that row i is the particle uploaded as row i (while no particle
        has been removed)."""
        n = self.particle_count()
        out = np.empty((n, 15), dtype=np.float64)
        self._ck(self.L.fsim_download_particles(self.h, out.ctypes.data, n, None))
        if by_id and n:
            out = out[np.argsort(self.download_particle_ids(), kind="stable")]
        return out

    def download_particle_ids(self):
        n = self.particle_count()
        ids = np.empty(n, dtype=np.uint32)
        self._ck(self.L.fsim_download_particle_ids(self.h, ids.ctypes.data, n))
        return ids

    def download_particle_cells(self, by_id=True):
        n = self.particle_count()
        out = np.empty(n, dtype=np.int32)
        self._ck(self.L.fsim_download_particle_cells(self.h, out.ctypes.data, n))
        if by_id and n:
            out = out[np.argsort(self.download_particle_ids(), kind="stable")]
        return out

    def upload_particles_f32(self, pos, vel=None, c=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float32).reshape(-1, 3)
        c = None if c is None else np.ascontiguousarray(c, dtype=np.float32).reshape(-1, 9)
        self._ck(self.L.fsim_upload_particles_f32(self.h, pos.ctypes.data, None if vel is None else vel.ctypes.data,
                                                  None if c is None else c.ctypes.data, pos.shape[0]))

    def download_particles_f32(self):
        n = self.particle_count()
        pos = np.empty((n, 3), dtype=np.float32)
        vel = np.empty((n, 3), dtype=np.float32)
        self._ck(self.L.fsim_download_particles_f32(self.h, pos.ctypes.data, vel.ctypes.data, None, n, None))
        return pos, vel

    # --- Simulator::simulate and its stages -------------------------------------------------------
    def step(self, dt):
        it = C.c_int32()
        self._ck(self.L.fsim_step(self.h, dt, C.byref(it)))
        return int(it.value)

    def stage_spawn(self, dt): self._ck(self.L.fsim_stage_spawn(self.h, dt))
    def stage_advect(self, dt): self._ck(self.L.fsim_stage_advect(self.h, dt))
    def stage_push_apart(self): self._ck(self.L.fsim_stage_push_apart(self.h))
    def stage_push_out(self): self._ck(self.L.fsim_stage_push_out(self.h))
    def stage_p2g(self, dt=0.0): self._ck(self.L.fsim_stage_p2g(self.h))
    def stage_classify(self, dt): self._ck(self.L.fsim_stage_classify(self.h, dt))
    def post_p2g_update(self, gravity_increment): self._ck(self.L.fsim_stage_post_p2g_update(self.h, gravity_increment))

    def stage_project(self, dt):
        it = C.c_int32()
        self._ck(self.L.fsim_stage_project(self.h, dt, C.byref(it)))
        return int(it.value)

    def stage_extrapolate(self): self._ck(self.L.fsim_stage_extrapolate(self.h))
    def stage_g2p(self): self._ck(self.L.fsim_stage_g2p(self.h))

    # --- grid -----------------------------------------------------------------------------------
    def download_grid(self, field):
        shape, dt = abi.field_shape_dtype(field, self.nc)
        out = np.zeros(shape, dtype=dt)
        self._ck(self.L.fsim_download_grid(self.h, field, out.ctypes.data, out.nbytes))
        return out

    def upload_grid(self, field, arr):
        shape, dt = abi.field_shape_dtype(field, self.nc)
        a = np.ascontiguousarray(arr, dtype=dt).reshape(shape)
        self._ck(self.L.fsim_upload_grid(self.h, field, a.ctypes.data, a.nbytes))

    def export_gfx(self, by_id=True, out=None):
        n = self.particle_count()
        if out is None:
            out = np.zeros((n, 5), dtype=np.float32)
        self._ck(self.L.fsim_export_gfx(self.h, out.ctypes.data, n, None))
        if by_id and n:
            out = out[np.argsort(self.download_particle_ids(), kind="stable")]
        return out

    def export_gfx_ptr(self, ptr, cap):
        """gfx export straight into caller-owned (pinned) host memory; returns the particle count."""
        n = C.c_int64()
        self._ck(self.L.fsim_export_gfx(self.h, ptr, cap, C.byref(n)))
        return int(n.value)

    def export_gfx_async_ptr(self, ptr, cap):
        """Enqueues the gfx export; the copy into `ptr` (pinned host memory) overlaps the following steps."""
        n = C.c_int64()
        self._ck(self.L.fsim_export_gfx_async(self.h, ptr, cap, C.byref(n)))
        return int(n.value)

    def export_gfx_strided_async_ptr(self, ptr, cap, stride):
        """Every `stride`-th particle only (an option for consumers that draw a subset); returns the record count."""
        n = C.c_int64()
        self._ck(self.L.fsim_export_gfx_strided_async(self.h, ptr, cap, int(stride), C.byref(n)))
        return int(n.value)

    def export_gfx_wait(self): self._ck(self.L.fsim_export_gfx_wait(self.h))
    def export_gfx_wait_previous(self): self._ck(self.L.fsim_export_gfx_wait_previous(self.h))

    # --- introspection --------------------------------------------------------------------------
    def step_durations(self):
        t = abi.Timings()
        self._ck(self.L.fsim_get_step_durations(self.h, C.byref(t)))
        return t.as_map()

    def timings(self):
        t = abi.Timings()
        self._ck(self.L.fsim_get_step_durations(self.h, C.byref(t)))
        return t

    def solve_info(self):
        s = abi.SolveInfo()
        self._ck(self.L.fsim_get_solve_info(self.h, C.byref(s)))
        return s

    def last_step_stats(self):
        ms, n = C.c_double(), C.c_int64()
        self._ck(self.L.fsim_get_last_step_stats(self.h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def synchronize(self): self._ck(self.L.fsim_synchronize(self.h))

    def synchronize_quiet(self):
        if getattr(self, "h", None):
            self.L.fsim_synchronize(self.h)

    def set_id_tracking(self, on): self._ck(self.L.fsim_set_id_tracking(self.h, int(on)))

    # --- measurement support ----------------------------------------------------------------------
    def kernel_classes(self):
        return [self.L.fsim_kernel_class_name(k).decode() for k in range(self.L.fsim_kernel_class_count())]

    def profile_enable(self, names):
        """Brackets every launch of the named kernel classes with CUDA events on the launching stream."""
        classes = self.kernel_classes()
        mask = 0
        for nm in names:
            mask |= 1 << classes.index(nm)
        self._ck(self.L.fsim_profile_enable(self.h, mask))

    def profile_read(self, reset=True):
        """{class: (event-timed ms, event-timed launches, all launches)} since the last reset."""
        classes = self.kernel_classes()
        ms = np.zeros(len(classes), dtype=np.float64)
        pn = np.zeros(len(classes), dtype=np.int64)
        ln = np.zeros(len(classes), dtype=np.int64)
        self._ck(self.L.fsim_profile_read(self.h, ms.ctypes.data, pn.ctypes.data, ln.ctypes.data, int(reset)))
        return {c: (float(ms[i]), int(pn[i]), int(ln[i])) for i, c in enumerate(classes)}

    def timer_record(self, slot): self._ck(self.L.fsim_timer_record(self.h, slot))

    def timer_elapsed_ms(self, a, b):
        ms = C.c_double()
        self._ck(self.L.fsim_timer_elapsed_ms(self.h, a, b, C.byref(ms)))
        return float(ms.value)

    def srand(self, seed):
        """Seeds libc rand(), which spawnParticles draws from (util/random.h:13-15)."""
        C.CDLL(None).srand(C.c_uint(seed))
