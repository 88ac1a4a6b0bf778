#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the REAL reference (oracle/_ref/libpbf_ref_strict.so = the unmodified
/root/reference/src/omp/ompsph.hpp, built by oracle/Makefile with -O2 -ffp-contract=off, run on ONE thread so that its
in-place delta pass is a deterministic Gauss-Seidel sweep — SURVEY.md F2).  Needs /root/reference at build time only.

    python tests/golden/make_golden.py

Vectors (all from sph::omp_impl::Solver<size_t,float>(0.1).advance on simpleConfigWith2Cubes scenes, sph.hpp:160-186):
  small_2cubes.npz   2 x 12^3 = 3456 particles, 3 iterations, moving wall, marching cubes ON:
                     the input and output particle arrays of frame 0 (t=0 lattice) and frame 30 (warm), the
                     reference's own sort permutation (recovered from ids) and its mesh at both frames.
  scene_2cubes.npz   2 x 10^3 = 2000 particles, 3 iterations, with tests/helpers.py demo_scene() (a well, a source, a drain,
                     two queries; sph.hpp:56-80): input/output particle arrays and query answers of calls 0 and 5.
  stock_hashes.npz   the stock benchmark scene (18 522 particles, 6 iterations, MC on, moving wall, benchmark.cpp:22-58):
                     sha256 of the particle array and of the mesh after each of the first 6 frames, plus vertex counts.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

H = 0.1
OUT = Path(__file__).resolve().parent


def perm_from_ids(before, after):
    where = np.empty(len(before), np.int64)
    where[before["id"]] = np.arange(len(before))
    return where[after["id"]].astype(np.uint32)


def main():
    assert oracle.ref_available("strict"), "build oracle/_ref first (make -C oracle)"
    # ---- small scene, full vectors
    p, xs = oracle.ref_scene_2cubes(4000, 3)
    p.surface_enabled = 1
    keep = {"params": np.frombuffer(bytes(p), np.uint8)}
    a = xs.copy()
    for frame in range(31):
        pf = oracle.ref_apply_motion(p, frame)
        before = a.copy()
        r = oracle.ref_advance(H, pf, a, variant="strict", threads=1, mesh_cap=400000)
        if frame in (0, 30):
            keep[f"f{frame}_in"] = before
            keep[f"f{frame}_out"] = a.copy()
            keep[f"f{frame}_perm"] = perm_from_ids(before, a)
            keep[f"f{frame}_params"] = np.frombuffer(bytes(pf), np.uint8)
            keep[f"f{frame}_vs"], keep[f"f{frame}_ns"], keep[f"f{frame}_cs"] = r["mesh_vs"], r["mesh_ns"], r["mesh_cs"]
    np.savez_compressed(OUT / "small_2cubes.npz", **keep)
    # ---- scene dynamics (stable-sort build: the scene pinning is about wells/sources/drains/queries, not the sort)
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import demo_scene
    sc = demo_scene()
    p, xs = oracle.ref_scene_2cubes(2000, 3)
    keep = {"params": np.frombuffer(bytes(p), np.uint8)}
    a = xs.copy()
    for call in range(6):
        before = a.copy()
        a, answers = oracle.ref_advance_scene(H, p, sc, a, variant="strict_stable", threads=1)
        if call in (0, 5):
            keep[f"c{call}_in"], keep[f"c{call}_out"] = before, a.copy()
            for (qid, ids) in answers:
                keep[f"c{call}_q{qid}"] = ids
    np.savez_compressed(OUT / "scene_2cubes.npz", **keep)
    # ---- stock scene, hashes
    p, xs = oracle.ref_scene_2cubes(20000, 6)
    p.surface_enabled = 1
    a = xs.copy()
    hp, hm, nv = [], [], []
    for frame in range(6):
        pf = oracle.ref_apply_motion(p, frame)
        r = oracle.ref_advance(H, pf, a, variant="strict_stable", threads=1, mesh_cap=400000)
        hp.append(hashlib.sha256(a.tobytes()).hexdigest())
        hm.append(hashlib.sha256(r["mesh_vs"].tobytes() + r["mesh_ns"].tobytes() + r["mesh_cs"].tobytes()).hexdigest())
        nv.append(r["n_vertices"])
    np.savez_compressed(OUT / "stock_hashes.npz", particles=np.array(hp), mesh=np.array(hm), n_vertices=np.array(nv),
                        variant=np.array(oracle.ref_lib("strict_stable").pbf_ref_variant().decode()))
    for f in ("small_2cubes.npz", "scene_2cubes.npz", "stock_hashes.npz"):
        print(f, (OUT / f).stat().st_size, "bytes")


if __name__ == "__main__":
    main()
