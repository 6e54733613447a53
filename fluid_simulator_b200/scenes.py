"""Synthetic dam-break scenes: identical initial conditions for the B200 path,
the compiled reference and the oracle restatement.

resolution 1 => cellD = 1; dam block = cells x in [1, N/2), y in [1, N-1), z in [1, N-1); 8 particles
per cell on stratified 2x2x2 sub-cell sites (0.25 / 0.75) plus a +-0.1 cell jitter from a
counter-based hash RNG (seed 1234); positions are rounded to fp32 and then widened, so the fp64
reference and the fp32 device path start from bit-identical coordinates; v = 0, C = 0.
"""
from dataclasses import dataclass, field

import numpy as np

from . import abi

SEED = 1234


def hash32(seed, idx, comp):
    """Counter-based 32-bit mix of (seed, particle id, component) -> uint32 (vectorised)."""
    x = (np.asarray(idx, dtype=np.uint64) * np.uint64(3) + np.uint64(comp)) & np.uint64(0xFFFFFFFF)
    x = (x ^ np.uint64(seed * 0x9E3779B9 & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF)
    for mul in (0x85EBCA6B, 0xC2B2AE35):
        x ^= x >> np.uint64(16)
        x = (x * np.uint64(mul)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x.astype(np.uint32)


def uniform01(seed, idx, comp):
    return hash32(seed, idx, comp).astype(np.float64) * (1.0 / 4294967296.0)


def block_particles(x0, x1, y0, y1, z0, z1, seed=SEED, two_d_z=None, jitter=0.1, id_offset=0):
    """8 (3D) or 4 (2D) jittered particles per cell of the half-open cell block; returns (n, 15)
    float64 in the reference `Particle` layout with fp32-representable positions."""
    if two_d_z is None:
        cx, cy, cz = np.meshgrid(np.arange(x0, x1), np.arange(y0, y1), np.arange(z0, z1), indexing="ij")
        cells = np.stack([cx.ravel(), cy.ravel(), cz.ravel()], axis=1).astype(np.float64)
        sub = np.array([[i, j, k] for i in (0.25, 0.75) for j in (0.25, 0.75) for k in (0.25, 0.75)])
    else:
        cx, cy = np.meshgrid(np.arange(x0, x1), np.arange(y0, y1), indexing="ij")
        cells = np.stack([cx.ravel(), cy.ravel(), np.zeros(cx.size)], axis=1).astype(np.float64)
        sub = np.array([[i, j, 0.0] for i in (0.25, 0.75) for j in (0.25, 0.75)])
    pos = (cells[:, None, :] + sub[None, :, :]).reshape(-1, 3)
    n = pos.shape[0]
    ids = np.arange(n, dtype=np.uint64) + np.uint64(id_offset)
    for a in range(3):
        pos[:, a] += (uniform01(seed, ids, a) * 2.0 - 1.0) * jitter
    if two_d_z is not None:
        pos[:, 2] = two_d_z
    out = np.zeros((n, 15), dtype=np.float64)
    out[:, 0:3] = pos.astype(np.float32).astype(np.float64)
    return out


def block_positions_f32(x0, x1, y0, y1, z0, z1, seed=SEED, jitter=0.1, chunk_x=8):
    """Same particles as block_particles (3D), but only the fp32 positions, built in x-chunks so that the 65 M-particle
    256^3 scene never materialises an fp64 AoS copy."""
    out = np.empty(((x1 - x0) * (y1 - y0) * (z1 - z0) * 8, 3), dtype=np.float32)
    sub = np.array([[i, j, k] for i in (0.25, 0.75) for j in (0.25, 0.75) for k in (0.25, 0.75)])
    per_x = (y1 - y0) * (z1 - z0) * 8
    for xa in range(x0, x1, chunk_x):
        xb = min(x1, xa + chunk_x)
        cx, cy, cz = np.meshgrid(np.arange(xa, xb), np.arange(y0, y1), np.arange(z0, z1), indexing="ij")
        cells = np.stack([cx.ravel(), cy.ravel(), cz.ravel()], axis=1).astype(np.float64)
        pos = (cells[:, None, :] + sub[None, :, :]).reshape(-1, 3)
        first = (xa - x0) * per_x
        ids = np.arange(pos.shape[0], dtype=np.uint64) + np.uint64(first)
        for a in range(3):
            pos[:, a] += (uniform01(seed, ids, a) * 2.0 - 1.0) * jitter
        out[first:first + pos.shape[0]] = pos.astype(np.float32)
    return out


@dataclass
class Scene:
    name: str
    dims: tuple
    resolution: float
    two_d: bool
    particle_radius: float
    dt: float
    params: abi.Params
    particles: np.ndarray
    obstacles: list = field(default_factory=list)
    steps: int = 1

    @property
    def n_particles(self):
        return int(self.particles.shape[0])


def default_params(transfer_type=abi.FLIP, max_iterations=2000, tol=1e-6, **kw):
    d = dict(transfer_type=transfer_type, flip_ratio=0.95, gravity=-9.81 * 4, gravity_enabled=True,
             push_apart_enabled=False, top_solid=False, pressure_enabled=True, max_iterations=max_iterations,
             pressure_k=1.0, average_pressure=8.0, fluid_density=1.0, residual_tolerance=tol)
    d.update(kw)
    return abi.make_params(**d)


def dam_break_3d(n, transfer_type=abi.FLIP, obstacles=None, seed=SEED, ny=None, nz=None, **param_kw):
    """3D dam break on an n x ny x nz grid (cfg 2: n=128 FLIP; cfg 3 headline: n=256 FLIP)."""
    ny = n if ny is None else ny
    nz = n if nz is None else nz
    parts = block_particles(1, n // 2, 1, ny - 1, 1, nz - 1, seed=seed)
    return Scene(name=f"dam3d_{n}x{ny}x{nz}_{('PIC', 'FLIP', 'APIC')[transfer_type]}", dims=(float(n), float(ny), float(nz)),
                 resolution=1.0, two_d=False, particle_radius=0.25, dt=0.005,
                 params=default_params(transfer_type, **param_kw), particles=parts, obstacles=list(obstacles or []))


def dam_break_2d(n=64, transfer_type=abi.FLIP, seed=SEED, **param_kw):
    """cfg 1: 2D FLIP dam break, n x n x 3 grid, block (n/2-1) x (n-2) cells x 8 particles, z pinned at dims.z/2.

    The reference's 2D mode uses 3 z-cells of size dims.z/3 (macGrid.cpp:10-12); with dims.z = 3 that is cellD.z = 1
    and particles sit at z = 1.5.  8 particles per (x,y) cell: two jittered 2x2 layers share the pinned plane."""
    a = block_particles(1, n // 2, 1, n - 1, 0, 1, seed=seed, two_d_z=1.5)
    b = block_particles(1, n // 2, 1, n - 1, 0, 1, seed=seed + 1, two_d_z=1.5, id_offset=a.shape[0])
    parts = np.concatenate([a, b], axis=0)
    return Scene(name=f"dam2d_{n}", dims=(float(n), float(n), 3.0), resolution=1.0, two_d=True, particle_radius=0.25,
                 dt=0.005, params=default_params(transfer_type, **param_kw), particles=parts, steps=200)


def cfg3_box(n):
    """Static box of SURVEY §8(d) cfg 3: size (24, 96, N) at (0.7N, 48, 0.5N)."""
    return abi.make_obstacle(abi.BOX, pos=(0.7 * n, 48.0 * n / 256.0 if n < 256 else 48.0, 0.5 * n),
                             size=(24.0 * n / 256.0 if n < 256 else 24.0, 96.0 * n / 256.0 if n < 256 else 96.0, float(n)))


def cfg4_obstacles(n, step, dt):
    """cfg 4: source sphere, sink sphere and a sinusoidally moving box; speed = (pos - prevPos)/dt as the
    manager computes it (obstacles.hpp:19-21, simulationManager.cpp:202-207)."""
    def box_x(s):
        return 0.6 * n + 0.1 * n * np.sin(2.0 * np.pi * s / 200.0)
    src = abi.make_obstacle(abi.SOURCE, pos=(0.25 * n, 0.8 * n, 0.5 * n), r=6.0 * n / 192.0, spawn_rate=4.0e5,
                            spawn_speed=4.0)
    sink = abi.make_obstacle(abi.SINK, pos=(0.8 * n, 0.08 * n, 0.5 * n), r=12.0 * n / 192.0)
    pos = (box_x(step), 0.25 * n, 0.5 * n)
    prev = (box_x(step - 1) if step > 0 else box_x(0), 0.25 * n, 0.5 * n)
    speed = tuple((p - q) / dt for p, q in zip(pos, prev))
    box = abi.make_obstacle(abi.BOX, pos=pos, prev_pos=prev, speed=speed,
                            size=(16.0 * n / 192.0, 48.0 * n / 192.0, 64.0 * n / 192.0))
    return [src, sink, box]


def hydrostatic_types(n, ny=None, nz=None):
    """cfg 5 (projection only): cell types with the dam block WATER, the rest of the interior AIR and the border
    shell SOLID except the open top (what the MacGrid constructor leaves behind, macGrid.cpp:15-16,75-140);
    reference order (x-major, z fastest)."""
    ny = n if ny is None else ny
    nz = n if nz is None else nz
    t = np.full((n, ny, nz), abi.AIR, dtype=np.uint8)
    t[0, :, :] = t[-1, :, :] = abi.SOLID
    t[:, 0, :] = abi.SOLID
    t[:, :, 0] = t[:, :, -1] = abi.SOLID
    t[1:n // 2, 1:ny - 1, 1:nz - 1] = abi.WATER
    return t.reshape(-1)
