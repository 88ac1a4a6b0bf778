"""Per-particle cost of ONE GPU on the dam-break family as the block grows (what bounds the cube family's weak-scaling
efficiency before any decomposition cost): for every side, settle 100 steps, time K resident steps, report ms/step,
ns per particle-iteration, and the candidates / in-radius neighbours per particle of the settled state.

    python profiles/tools/size_sweep.py [side ...]  ->  one JSON line per side"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))


def main():
    import torch
    from pbf_sph_b200 import FLAG_DEBUG_COUNTS, FLAG_PROFILE, Solver, capi, scenes
    sides = [int(a) for a in sys.argv[1:]] or [100, 126, 159, 200]
    stream = torch.cuda.Stream()
    for side in sides:
        p, xs = scenes.dam_break(side, 4)
        n = len(xs)
        with Solver(scenes.H, 0) as s:
            s.set_stream(stream.cuda_stream)
            s.upload(xs)
            del xs
            for _ in range(100):
                s.step(p)
            s.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k = 20
            e0.record(stream)
            for _ in range(k):
                s.step(p)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / k
            s.set_flags(FLAG_PROFILE)
            s.profile_mask(["lambda", "delta"])
            s.profile_reset()
            for _ in range(k):
                s.step(p)
            s.sync()
            fam = {f: round(v / k, 4) for f, v in s.profile()["ms"].items() if v > 0}
            s.profile_mask(None)
            s.set_flags(FLAG_DEBUG_COUNTS)
            s.step(p)
            cand = s.tap(capi.TAP_CAND_COUNT).astype(np.int64)
            hits = s.tap(capi.TAP_NBR_COUNT).astype(np.int64)
        print(json.dumps({"side": side, "particles": n, "ms_per_step": round(ms, 4), "ms_by_family": fam,
                          "ns_per_particle_iteration": round(ms * 1e6 / (n * 4), 4),
                          "candidates_per_particle": round(float(cand.mean()), 2),
                          "neighbours_per_particle": round(float(hits.mean()), 2)}), flush=True)


if __name__ == "__main__":
    main()
