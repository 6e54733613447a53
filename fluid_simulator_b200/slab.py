"""z-slab groups of the B200 Simulator (include/fsim.h "z-slab decomposition", csrc/dist.cu).

One slab handle per rank; ranks exchange ghost planes, migrating particles and the PCG's reductions through NVLink
peer memory inside the library.  This module is the host-side plumbing the library leaves to its caller:

* `connect_torch(sim, dist)`   the production path: one process per GPU (torchrun), the ranks' exports travel through
                               `torch.distributed.all_gather_object` and every rank maps its neighbours over CUDA IPC;
* `SlabGroup`                  all ranks inside ONE process (one host thread per rank, plain peer pointers; the handles may
                               even share a device): what the parity tests use to compare a sharded run with the
                               single-handle run cell by cell and particle by particle;
* `partition(...)` / `owner_of(...)` / `stitch(...)`  pure numpy helpers (which planes a rank owns, which rank owns a
                               particle, how local blocks become the global grid) -- covered by CPU tests.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import abi
from .sim import FluidSim, load_library


def partition(global_gz, nranks):
    """[(own_lo, own_hi, z_offset, gz_local)] per rank: same arithmetic as the library (fsim_slab_partition)."""
    out = []
    half = global_gz // 2
    for r in range(nranks):
        # slab boundaries are even planes (the multigrid's 2x2x2 aggregates never straddle two ranks)
        lo = 0 if r == 0 else 2 * (half * r // nranks)
        hi = global_gz if r == nranks - 1 else 2 * (half * (r + 1) // nranks)
        if nranks == 1:
            out.append((0, global_gz, 0, global_gz))
            continue
        z0, z1 = max(lo - 1, 0), min(hi + 1, global_gz)
        out.append((lo, hi, z0, z1 - z0))
    return out


def owner_of(pos_z, cell_d_inv_z, global_gz, nranks):
    """Rank that owns each particle: the plane is trunc(float32(z) * cellDInv.z) evaluated in fp64, the expression the
    device bins with (simulator.cpp:358-359)."""
    z = np.asarray(pos_z, dtype=np.float32).astype(np.float64)
    iz = np.clip(np.trunc(z * cell_d_inv_z).astype(np.int64), 0, global_gz - 1)
    bounds = np.array([p[1] for p in partition(global_gz, nranks)], dtype=np.int64)  # own_hi of every rank
    return np.searchsorted(bounds, iz, side="right").astype(np.int64)


def stitch(local_blocks, parts, grid_size, trailing=()):
    """Global reference-order array (x, y, z[, ...]) from the ranks' local blocks, taking each plane from its owner."""
    gx, gy, gz = grid_size
    out = None
    for blk, (lo, hi, zoff, gzl) in zip(local_blocks, parts):
        b = np.asarray(blk).reshape((gx, gy, gzl) + tuple(trailing))
        if out is None:
            out = np.zeros((gx, gy, gz) + tuple(trailing), dtype=b.dtype)
        out[:, :, lo:hi] = b[:, :, lo - zoff:hi - zoff]
    return out


def connect_torch(sim, dist):
    """Production wiring: every rank publishes its export, gathers all of them and maps its neighbours (CUDA IPC)."""
    ex = sim.dist_export()
    blobs = [None] * dist.get_world_size()
    dist.all_gather_object(blobs, bytes(ex))
    err = None
    try:
        sim.dist_connect([abi.DistExport.from_buffer_copy(b) for b in blobs])
    except Exception as e:  # noqa: BLE001 -- e.g. CUDA IPC not permitted in this container
        err = f"rank {dist.get_rank()}: {e}"
    errs = [None] * dist.get_world_size()
    dist.all_gather_object(errs, err)   # doubles as the barrier; every rank learns about every failure and raises together
    bad = [e for e in errs if e]
    if bad:
        raise RuntimeError("slab connect failed: " + "; ".join(bad))


class SlabGroup:
    """N slab handles driven from one process, one host thread per rank (fsim_step on a slab handle is collective: every
    rank must be inside it at the same time).  Interface mirrors FluidSim where the tests need it."""

    def __init__(self, nranks, dims, resolution=1.0, two_d=False, particle_radius=0.25, capacity=0, devices=None, **_):
        if os.environ.get("CUDA_MODULE_LOADING", "").upper() != "EAGER":
            # a rank's exchange kernel spins until its neighbour's kernel runs; with lazy loading the neighbour's first launch of
            # a kernel waits for the device to go idle first -> both wait for ever (only when ranks share a process)
            raise RuntimeError("SlabGroup needs CUDA_MODULE_LOADING=EAGER in the environment before CUDA is initialised")
        load_library()
        self.n = int(nranks)
        devices = list(devices) if devices is not None else [0] * self.n
        if len(set(devices)) < len(devices):
            # ranks share a device: kernels of different ranks must be able to co-reside while they wait for each other, so
            # none of them may need a whole GPC (multigrid tail cluster) or most of the SMs (gather blocks) for itself
            os.environ.setdefault("FSIM_MG_TAIL_CLUSTER", "2")
            os.environ.setdefault("FSIM_DIST_GATHER_BLOCKS", "16")
        self.sims = [FluidSim(dims, resolution, two_d, particle_radius, capacity=capacity, device=devices[r], rank=r, nranks=self.n)
                     for r in range(self.n)]
        s0 = self.sims[0]
        self.info = s0.info
        self.grid_size = s0.grid_size
        self.parts = partition(self.grid_size[2], self.n)
        for r, s in enumerate(self.sims):
            assert (s.slab.own_lo, s.slab.own_hi, s.slab.z_offset, s.slab.gz_local) == self.parts[r], "partition mismatch"
        exports = [s.dist_export() for s in self.sims]
        for s in self.sims:
            s.dist_connect(exports)
        self.pool = ThreadPoolExecutor(max_workers=self.n)

    def close(self):
        for s in self.sims:
            s.synchronize_quiet()
        for s in self.sims:
            s.close()
        self.pool.shutdown()

    def _all(self, fn):
        """fn(rank, sim) on every rank concurrently; re-raises the first failure after all threads returned."""
        futs = [self.pool.submit(fn, r, s) for r, s in enumerate(self.sims)]
        res, err = [], None
        for f in futs:
            try:
                res.append(f.result())
            except Exception as e:  # noqa: BLE001
                err = err or e
                res.append(None)
        if err:
            raise err
        return res

    def set_params(self, params):
        for s in self.sims:
            s.set_params(params)

    def set_obstacles(self, obstacles):
        for s in self.sims:
            s.set_obstacles(obstacles)

    def upload_particles(self, aos15):
        a = np.ascontiguousarray(aos15, dtype=np.float64).reshape(-1, 15)
        own = owner_of(a[:, 2], self.info.cell_d_inv[2], self.grid_size[2], self.n)
        for r, s in enumerate(self.sims):
            idx = np.nonzero(own == r)[0]
            s.upload_particles(a[idx])
            s.upload_particle_ids(idx.astype(np.uint32))  # ids stay the global row numbers

    def step(self, dt):
        its = self._all(lambda r, s: s.step(dt))
        if len(set(its)) != 1:
            self.synchronize()  # a timed-out exchange surfaces here as FsimError(ERR_COMM) with the wait that gave up
        assert len(set(its)) == 1, f"ranks disagree on the PCG iteration count: {its}"
        return its[0]

    def synchronize(self):
        self._all(lambda r, s: s.synchronize())

    def particle_counts(self):
        return [s.particle_count() for s in self.sims]

    def download_particles(self):
        """All particles in global-id order (row i = the particle uploaded as row i while none has been removed)."""
        self.synchronize()
        parts = [s.download_particles(by_id=False) for s in self.sims]
        ids = [s.download_particle_ids() for s in self.sims]
        a, i = np.concatenate(parts), np.concatenate(ids)
        return a[np.argsort(i, kind="stable")]

    def download_particle_cells(self):
        self.synchronize()
        cells = np.concatenate([s.download_particle_cells(by_id=False) for s in self.sims])
        ids = np.concatenate([s.download_particle_ids() for s in self.sims])
        return cells[np.argsort(ids, kind="stable")]

    def download_grid(self, field):
        """Global grid in the reference layout, every plane taken from the rank that owns it."""
        self.synchronize()
        blocks = [s.download_grid(field) for s in self.sims]
        trailing = (3,) if field in (abi.FIELD_V, abi.FIELD_V2, abi.FIELD_WSUM) else ()
        g = stitch(blocks, self.parts, self.grid_size, trailing)
        return g.reshape((-1,) + trailing)

    def upload_grid(self, field, arr):
        """Global reference-order array -> every rank's local block (owned planes and ghost planes alike)."""
        gx, gy, gz = self.grid_size
        trailing = (3,) if field in (abi.FIELD_V, abi.FIELD_V2, abi.FIELD_WSUM) else ()
        g = np.asarray(arr).reshape((gx, gy, gz) + trailing)
        for s, (lo, hi, zoff, gzl) in zip(self.sims, self.parts):
            s.upload_grid(field, np.ascontiguousarray(g[:, :, zoff:zoff + gzl]))

    def post_p2g_update(self, gravity_increment):
        for s in self.sims:
            s.post_p2g_update(gravity_increment)

    def stage_project(self, dt):
        its = self._all(lambda r, s: s.stage_project(dt))
        if len(set(its)) != 1:
            self.synchronize()
        assert len(set(its)) == 1, f"ranks disagree on the PCG iteration count: {its}"
        return its[0]

    def download_grid_local(self, field, rank):
        self.synchronize()
        return self.sims[rank].download_grid(field)

    def export_gfx(self):
        self.synchronize()
        out = np.concatenate([s.export_gfx(by_id=False) for s in self.sims])
        ids = np.concatenate([s.download_particle_ids() for s in self.sims])
        return out[np.argsort(ids, kind="stable")]

    def solve_info(self):
        return self.sims[0].solve_info()
