#!/usr/bin/env python
"""Slab path on ONE GPU (LOCAL transport), per-step trace of every rank's owned / ghost counts — to see how the ghost sets
move between two plan steps (capacity sizing).   python profiles/tools/slab_trace.py [world=8] [side=200] [steps=120] [replan=8]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from pbf_sph_b200 import scenes
from pbf_sph_b200.dist import LocalGroup

world, side, steps, replan = (int(a) for a in (sys.argv[1:] + ["8", "200", "120", "8"][len(sys.argv) - 1:]))
p, xs = scenes.dam_break(side, 4)
with LocalGroup(scenes.H, [0] * world) as g:
    for r in g.ranks:
        r.set_replan(replan)
    g.upload(xs)
    for step in range(steps):
        try:
            g.step(p)
            g.sync()
            st = [r.stats() for r in g.ranks]
        except Exception as e:  # noqa: BLE001
            print("step", step, "FAILED:", e)
            break
        print(step, "own", [s["owned"] for s in st], "gh", [s["ghosts"] for s in st], "r1", [s["ghost_ring1"] for s in st], flush=True)
