"""Scene factory: the reference's only scene (src/sph.hpp:127-186) and the dam-break family the benchmarks use.

All scenes are deterministic lattices (no RNG), mass 1, zero initial velocity."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from .capi import PARTICLE, Params, lib

H = 0.1  # smoothing length every driver passes to the solver constructor (benchmark.cpp:160-163)


def base_params(iteration: int, scale: float = 500.0) -> Params:
    """SphParams of simpleConfigWith2Cubes — sph.hpp:168-177; McParams — sph.hpp:179-184 (disabled by default)."""
    p = Params()
    p.dt = np.float32(0.0083 * np.float32(1.5))  # `0.0083 * 1.5f` is a double product narrowed to float
    p.scale = scale
    p.iteration = iteration
    p.constant_force[:] = (0.0, 9.8, 0.0)
    p.min_bound[:] = (0.0, 0.0, 0.0)
    p.max_bound[:] = (1000.0, 1000.0, 1000.0)
    p.wait = 1
    p.surface_enabled = 0
    p.surface.resolution = 2.0
    p.surface.isolevel = 100.0
    p.surface.particle_size = 25.0
    p.surface.particle_influence = 0.5
    return p


def make_cube(first_id: int, spacing: float, side: int, origin, colour) -> np.ndarray:
    """makeCube — sph.hpp:127-145: side^3 particles, x slowest / z fastest, pos = (x,y,z)*spacing + origin."""
    xs = np.zeros(side ** 3, PARTICLE)
    g = np.arange(side, dtype=np.float32)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    lattice = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    xs["position"] = lattice * np.float32(spacing) + np.asarray(origin, np.float32)
    xs["id"] = first_id + np.arange(side ** 3, dtype=np.uint64)
    xs["mass"] = 1.0
    xs["colour"] = np.asarray(colour, np.float32)
    return xs


def two_cubes(count: int = 20000, iteration: int = 6, scale: float = 500.0):
    """simpleConfigWith2Cubes(count, iteration, scale) — sph.hpp:160-186.  count=20000 -> 2 x 21^3 = 18522."""
    side = int(math.cbrt(count // 2))  # static_cast<size_t>(std::cbrt(count)) — sph.hpp:134
    a = make_cube(0, 22.0, side, (100.0, 0.0, 100.0), (0.0, 0.1, 0.8, 1.0))
    b = make_cube(len(a), 22.0, side, (600.0, 0.0, 600.0), (0.1, 0.8, 0.1, 1.0))
    return base_params(iteration, scale), np.concatenate([a, b])


def apply_motion(params: Params, frame: int) -> Params:
    """applyMotionSinXCosZ — sph.hpp:147-158 (evaluated by the library's host code with the C float sin/cos)."""
    out = Params()
    lib().pbf_host_apply_motion(C.byref(params), frame, C.byref(out))
    return out


def dam_break(side: int, iteration: int = 4, scale: float = 500.0):
    """dam(side): one side^3 block, spacing 22, origin (100, Ly - 22*side - 50, 100) in a box
    Lx = 44*side + 200, Ly = Lz = 22*side + 200.  +y is the direction of gravity."""
    lx, ly = 44.0 * side + 200.0, 22.0 * side + 200.0
    p = base_params(iteration, scale)
    p.max_bound[:] = (lx, ly, ly)
    xs = make_cube(0, 22.0, side, (100.0, ly - 22.0 * side - 50.0, 100.0), (0.0, 0.1, 0.8, 1.0))
    return p, xs


def dam_break_shard(side: int, iteration: int, rank: int, world: int, scale: float = 500.0):
    """Block `rank` of `world` of dam_break(side)'s particle array (the blocks dist.shard() cuts), generated directly:
    a 64 M-particle array is 3.6 GB, which eight ranks should not each build in full."""
    lx, ly = 44.0 * side + 200.0, 22.0 * side + 200.0
    p = base_params(iteration, scale)
    p.max_bound[:] = (lx, ly, ly)
    n = side ** 3
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    i = np.arange(lo, hi, dtype=np.int64)
    xs = np.zeros(hi - lo, PARTICLE)
    lattice = np.stack([i // (side * side), (i // side) % side, i % side], axis=1).astype(np.float32)
    xs["position"] = lattice * np.float32(22.0) + np.asarray((100.0, ly - 22.0 * side - 50.0, 100.0), np.float32)
    xs["id"] = i.astype(np.uint64)
    xs["mass"] = 1.0
    xs["colour"] = np.asarray((0.0, 0.1, 0.8, 1.0), np.float32)
    return p, xs, n


WORKLOADS = {
    # name: (factory, kwargs) — BASELINE.json configs
    "ref-2cubes": lambda: two_cubes(20000, 6),
    "dam-64k": lambda: dam_break(40, 4),
    "dam-1m": lambda: dam_break(100, 4),
    "dam-8m": lambda: dam_break(200, 4),
}
