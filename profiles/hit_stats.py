import sys, numpy as np
sys.path.insert(0, '/root/repo')
from pbf_sph_b200 import Solver, capi, scenes
p, xs = scenes.dam_break(100, 4)
with Solver(scenes.H, 0, capi.FLAG_DEBUG_COUNTS) as s:
    s.upload(xs)
    for f in range(301):
        s.step(p)
        if f in (0, 1, 5, 20, 50, 100, 150, 200, 300):
            nb, cd = s.tap(capi.TAP_NBR_COUNT), s.tap(capi.TAP_CAND_COUNT)
            print(f, 'hits mean %.1f p50 %d p90 %d p99 %d max %d >48: %.3f >64: %.3f | cand mean %.1f p99 %d max %d' % (
                nb.mean(), *np.percentile(nb, [50, 90, 99]), nb.max(), (nb > 48).mean(), (nb > 64).mean(), cd.mean(), np.percentile(cd, 99), cd.max()), flush=True)
