#!/usr/bin/env python
"""Slab path on ONE GPU (LOCAL transport: every rank is a context of this process) — for `ncu` launch lists of the
multi-GPU bookkeeping kernels, which cannot be profiled under a multi-rank launch.
    python profiles/slab_local.py [world=2] [side=126] [settle=20] [steps=2]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pbf_sph_b200 import scenes
from pbf_sph_b200.dist import LocalGroup

world, side, settle, steps = (int(a) for a in (sys.argv[1:] + ["2", "126", "20", "2"][len(sys.argv) - 1:]))
p, xs = scenes.dam_break(side, 4)
with LocalGroup(scenes.H, [0] * world) as g:
    g.upload(xs)
    for _ in range(settle):
        g.step(p)
    g.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        g.step(p)
    g.sync()
    dt = (time.perf_counter() - t0) / steps
    print(f"world={world} n={len(xs)} {dt * 1e3:.3f} ms/step (ranks serialised on one GPU)", [r.stats() for r in g.ranks])
