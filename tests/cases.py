"""Small seeded scenes shared by the golden-fixture generator, the oracle pin tests and the GPU
parity tests.  Each case is (scene, per-step obstacle function or None, number of steps)."""
import numpy as np

from fluid_simulator_b200 import abi, scenes


def _moving_sphere(n, step, dt):
    def cx(s):
        return 0.3 * n + 0.05 * n * np.sin(2.0 * np.pi * s / 20.0)
    pos = (cx(step), 0.3 * n, 0.5 * n)
    prev = (cx(step - 1) if step > 0 else cx(0), 0.3 * n, 0.5 * n)
    return abi.make_obstacle(abi.SPHERE, pos=pos, prev_pos=prev, r=0.12 * n,
                             speed=tuple((p - q) / dt for p, q in zip(pos, prev)))


def case_2d_flip(n=24, steps=10):
    return scenes.dam_break_2d(n), None, steps


def case_3d_flip(n=12, steps=3):
    return scenes.dam_break_3d(n, abi.FLIP), None, steps


def case_3d_apic_obstacles(n=12, steps=4):
    sc = scenes.dam_break_3d(n, abi.APIC)
    box = abi.make_obstacle(abi.BOX, pos=(0.7 * n, 0.25 * n, 0.5 * n), size=(0.2 * n, 0.4 * n, 0.5 * n))
    return sc, (lambda st: [box, _moving_sphere(n, st, sc.dt)]), steps


def case_3d_pic_source_sink(n=16, steps=4):
    """cfg-4 style: source + sink placed INSIDE the dam block (so despawning is exercised) + moving box."""
    sc = scenes.dam_break_3d(n, abi.PIC, spawning_enabled=True, despawning_enabled=True)

    def obs(st):
        o = scenes.cfg4_obstacles(n, st, sc.dt)
        o[0].spawn_rate = 4.0e4
        o[1].pos[:] = (0.3 * n, 0.3 * n, 0.5 * n)
        o[1].prev_pos[:] = o[1].pos[:]
        o[1].r = 0.1 * n
        return o
    return sc, obs, steps


def case_3d_flip_nonunit(n=10, steps=3):
    """non-power-of-two cell size (resolution 1.508 like the GUI's 3D preset), top of container solid."""
    res = 1.508
    dims = (n / res + 1e-9, n / res + 1e-9, n / res + 1e-9)
    sc = scenes.dam_break_3d(n, abi.FLIP, top_solid=True)
    sc.dims = dims
    sc.resolution = res
    sc.particle_radius = 0.25 / res
    sc.particles = sc.particles.copy()
    sc.particles[:, 0:3] = (sc.particles[:, 0:3] / res).astype(np.float32).astype(np.float64)
    sc.name = f"dam3d_{n}_res1508"
    return sc, None, steps


CASES = {
    "2d_flip": case_2d_flip,
    "3d_flip": case_3d_flip,
    "3d_apic_obstacles": case_3d_apic_obstacles,
    "3d_pic_source_sink": case_3d_pic_source_sink,
    "3d_flip_nonunit": case_3d_flip_nonunit,
}

GRID_FIELDS = [(abi.FIELD_TYPE, "type"), (abi.FIELD_V, "v"), (abi.FIELD_V2, "v2"), (abi.FIELD_WSUM, "wsum"),
               (abi.FIELD_AVGPNUM, "avgp"), (abi.FIELD_PRESSURE, "pressure")]


def run_case(sim_cls, name, record_every_step=False, **sim_kw):
    """Runs a case through any object with the RefSim/OracleSim/FluidSim interface; returns dict of arrays."""
    sc, obs_fn, steps = CASES[name]()
    sim = sim_cls(sc.dims, sc.resolution, sc.two_d, sc.particle_radius, **sim_kw)
    sim.set_params(sc.params)
    sim.upload_particles(sc.particles)
    its = []
    obs = list(sc.obstacles)
    for st in range(steps):
        if obs_fn is not None:
            new = obs_fn(st)
            if st > 0:  # carry SphericalParticleSource::lastSpawnFraction like the live obstacle object does
                old = sim.get_obstacles(obs)
                for i, x in enumerate(new):
                    x.last_spawn_fraction = old[i].last_spawn_fraction
            obs = new
        sim.set_obstacles(obs)
        sim.srand(1000 + st)  # spawnParticles draws from libc rand() (util/random.h:13-15)
        its.append(sim.step(sc.dt))
    out = {"its": np.array(its, dtype=np.int32), "particles": sim.download_particles(),
           "cells": sim.download_particle_cells()}
    for f, nm in GRID_FIELDS:
        out[nm] = sim.download_grid(f)
    out["gfx"] = sim.export_gfx()
    sim.close()
    return out
