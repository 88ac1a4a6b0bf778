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
