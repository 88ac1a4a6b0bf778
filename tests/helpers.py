"""Shared helpers of the parity tests."""
from __future__ import annotations

import numpy as np

DOMAIN = 1000.0  # box edge of the stock scene, in unscaled units (sph.hpp:173)


def by_id(xs: np.ndarray) -> np.ndarray:
    return xs[np.argsort(xs["id"], kind="stable")]


def warm(oracle, h, params, xs, steps, motion=None):
    """Advance `xs` in place with the Jacobi oracle for `steps` frames (optionally with the moving wall)."""
    for f in range(steps):
        p = motion(params, f) if motion else params
        oracle.step(h, p, xs)
    return xs


def frac_within(a: np.ndarray, b: np.ndarray, tol: float) -> float:
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return float((d <= tol).mean())


def demo_scene():
    """A sph::Scene (sph.hpp:75-80) that exercises every scene feature on the two-cube scenes: a well above the first
    cube, a source between the cubes (rate 20 -> a 4 x 5 sheet per call, ids = tag), a drain biting a corner of the
    first cube, and two queries (one inside the first cube, one in empty space)."""
    from pbf_sph_b200 import capi
    return capi.Scene(wells=[(7, (300.0, 200.0, 300.0), 5000.0)],
                      sources=[(99, (500.0, 100.0, 500.0), (0.0, 1.0, 0.0), (1.0, 0.0, 0.0, 1.0), 20.0)],
                      drains=[(1, (120.0, 20.0, 120.0), 40.0, 1.0)],
                      queries=[(11, (150.0, 60.0, 150.0)), (12, (900.0, 900.0, 900.0))])
