"""Parity on BASELINE.json's OWN configurations (the workloads bench.py times), GPU through the C ABI vs the CPU oracle.

  configs[0]  dam-break 40^3 = 64 000 particles, 4 iterations, 100 steps : aggregate statistics (mean density error,
              kinetic energy) — the dynamics are chaotic, SURVEY §8c
  configs[1]  dam-break 100^3 = 1 000 000 particles, 4 iterations         : the bench state (settled on the GPU), then ONE
              step on both sides: keys / stable permutation / cell table / candidate and in-radius counts bit-exact, the
              production neighbour list's hit counts, positions within 1e-5 of the domain (1e-7 under STRICT_FP)
  configs[3]  the same 1 M particles with marching cubes (lattice 191 x 103 x 103): NaN pattern, field, triangle count
              exact, mesh vertex by vertex

Reference semantics: ompsph.hpp:128-271 (step), :277-477 (surface), sph.hpp:119-125 (the call).
"""
import numpy as np
import pytest

from helpers import by_id, frac_within
from pbf_sph_b200 import FLAG_DEBUG_COUNTS, FLAG_STRICT_FP, Solver, capi, scenes
from test_parity_gpu import assert_integer_parity, run_gpu

pytestmark = pytest.mark.gpu
H = scenes.H
SETTLE = 100  # bench.py --settle


def box_edge(p) -> float:
    return float(max(p.max_bound[a] - p.min_bound[a] for a in range(3)))


@pytest.fixture(scope="module")
def dam_1m_settled(gpu):
    """dam(100) after bench.py's 100 settling steps, advanced on the GPU (resident path) and downloaded."""
    p, xs = scenes.dam_break(100, 4)
    with Solver(H, 0) as s:
        s.upload(xs)
        for _ in range(SETTLE):
            s.step(p)
        s.sync()
        snap = s.download()
    assert np.array_equal(np.sort(snap["id"]), np.arange(len(xs), dtype=np.uint64))
    return p, snap


@pytest.fixture(scope="module")
def dam_1m_oracle_step(dam_1m_settled, oracle_mod):
    p, snap = dam_1m_settled
    cpu = snap.copy()
    return cpu, oracle_mod.step(H, p, cpu, taps=True)


@pytest.mark.parametrize("flags,pos_tol", [(0, 1e-5), (FLAG_STRICT_FP, 1e-7)])
def test_dam_1m_one_step_against_oracle(gpu, dam_1m_settled, dam_1m_oracle_step, flags, pos_tol, record_property):
    """BASELINE configs[1] in the very state bench.py times."""
    p, snap = dam_1m_settled
    cpu, t_cpu = dam_1m_oracle_step
    gpu_xs, t_gpu, _ = run_gpu(p, snap, flags)
    assert t_gpu["grid"].grid_table_n == t_cpu["grid"].grid_table_n == 488063
    list_mismatch = assert_integer_parity(t_gpu, t_cpu, flags)
    assert np.array_equal(gpu_xs["id"], cpu["id"]), "output order (Z-sorted, ids carried)"
    assert np.array_equal(gpu_xs["colour"], cpu["colour"]), "diffused colours"
    domain = box_edge(p)  # 4600 units
    dp = np.abs(gpu_xs["position"].astype(np.float64) - cpu["position"])
    dv = np.abs(gpu_xs["velocity"].astype(np.float64) - cpu["velocity"])
    record_property("max_dx_over_domain", float(dp.max() / domain))
    record_property("max_dv", float(dv.max()))
    record_property("list_hit_mismatches", list_mismatch)
    print(f"dam-1m flags={flags}: max|dx|={dp.max():.3e} ({dp.max() / domain:.2e} of the domain), max|dv|={dv.max():.3e}, "
          f"hit-list mismatches={list_mismatch}, mean candidates={t_cpu['cand_count'].mean():.1f}, "
          f"mean neighbours={t_cpu['nbr_count'].mean():.1f}")
    # The bar (BASELINE.json north_star): >= 99.9 % of the position components within pos_tol x domain.  A settled 1 M-particle
    # dam break also holds about a hundred particles in violent neighbourhoods (clamped against a wall, 30-45 neighbours
    # inside h) where the solver amplifies the 1e-7 relative differences of the production arithmetic (FMA contraction,
    # rsqrt) from one iteration to the next; with PBF_FLAG_STRICT_FP the same step agrees with the oracle to 1.3e-7 of the
    # domain on every particle, so these are arithmetic, not logic.  They are bounded to 1 particle in 10^4 and reported.
    worst = dp.max(axis=1)
    out = np.flatnonzero(worst > 10 * pos_tol * domain)
    print(f"  components beyond {pos_tol:g} x domain: {int((dp > pos_tol * domain).sum())} of {dp.size}; particles beyond "
          f"{10 * pos_tol:g} x domain: {len(out)}")
    if len(out):
        from scipy.spatial import cKDTree
        before = by_id(snap)["position"].astype(np.float64) / float(p.scale)
        tree = cKDTree(before)
        ids = cpu["id"][out].astype(np.int64)
        near = tree.query(before[ids], k=2)[0][:, 1] / H  # nearest neighbour before the step, in units of h
        crowd = np.array([len(c) for c in tree.query_ball_point(before[ids], 1.0 * H)])
        print(f"  those particles: nearest neighbour {near.min():.2e} h .. {np.median(near):.2e} h (median), neighbours within h: "
              f"{crowd.min()} .. {crowd.max()}")
    assert frac_within(gpu_xs["position"], cpu["position"], pos_tol * domain) >= 0.999, dp.max()
    assert len(out) <= 1.5e-4 * len(worst), (len(out), dp.max())
    # a velocity is the step's displacement / dt * VD (ompsph.hpp:261): its tolerance follows from the position tolerance
    ok = worst <= 10 * pos_tol * domain
    vel_bound = 1.05 * (10 * pos_tol * domain / float(p.scale)) / float(p.dt) * 0.49
    assert dv[ok].max() <= vel_bound, (dv[ok].max(), vel_bound)
    lam_scale = np.abs(t_cpu["lambda"]).max()
    assert np.quantile(np.abs(t_gpu["lambda"] - t_cpu["lambda"]), 0.9999) <= (1e-6 if flags & FLAG_STRICT_FP else 1e-3) * lam_scale
    if flags & FLAG_STRICT_FP:
        assert len(out) == 0 and dp.max() <= 2 * pos_tol * domain


def test_dam_1m_surface_against_oracle(gpu, dam_1m_settled, oracle_mod):
    """BASELINE configs[3]: the 1 M-particle dam break with marching cubes (ompsph.hpp:277-477) on its full lattice."""
    p, snap = dam_1m_settled
    pm = p.copy()
    pm.surface_enabled = 1
    cpu, dev = snap.copy(), snap.copy()
    t = oracle_mod.step(H, pm, cpu, taps=True)
    with Solver(H, 0, FLAG_STRICT_FP) as s:
        res = s.advance(pm, dev)
        field, colour = s.tap(capi.TAP_MC_FIELD), s.tap(capi.TAP_MC_COLOUR)
        g = s.grid()
    assert list(g.sample_size) == list(t["grid"].sample_size) == [191, 103, 103]
    assert np.array_equal(np.isnan(field), np.isnan(t["mc_field"]))
    assert np.array_equal(np.isnan(colour), np.isnan(t["mc_colour"]))
    assert np.allclose(field[:, 0], t["mc_field"][:, 0], rtol=2e-5, atol=1e-4)
    assert np.allclose(field[:, 1:], t["mc_field"][:, 1:], rtol=0, atol=2e-4, equal_nan=True)
    assert np.allclose(colour, t["mc_colour"], rtol=1e-5, atol=1e-6, equal_nan=True)
    assert g.n_triangles * 3 == len(res.vs) == t["n_vertices"], (len(res.vs), t["n_vertices"])
    assert len(res.vs) > 100000, "the scene must produce a real surface"
    domain = box_edge(p)
    assert np.allclose(res.vs, t["mesh_vs"], rtol=0, atol=2e-5 * domain, equal_nan=True)
    assert np.allclose(res.ns, t["mesh_ns"], rtol=0, atol=2e-3, equal_nan=True)
    assert np.allclose(res.cs, t["mesh_cs"], rtol=0, atol=2e-4, equal_nan=True)


def test_dam_64k_100_steps_aggregates(gpu, oracle_mod):
    """BASELINE configs[0]: dam(40), 4 iterations, 100 steps — the configuration the reference's CPU benchmark runs.

    Mean density error within +-0.01 of the oracle's.  Kinetic energy: the wave is breaking around step 100 and the sum of
    v^2 is chaotic there — the ORACLE ITSELF, started from positions nudged by 1e-6 of the domain, ends 19 % away from its
    own unperturbed run (it is 1-2 % at step 50).  So the band at each checkpoint is max(10 %, 1.5 x the oracle's own
    sensitivity to that nudge), measured in the same test."""
    p, xs = scenes.dam_break(40, 4)
    cpu, nudged = xs.copy(), xs.copy()
    nudged["position"][::97, 0] += np.float32(1e-3)
    ke = lambda a: 0.5 * float((a["velocity"].astype(np.float64) ** 2).sum())
    got = {}
    with Solver(H, 0) as s:
        s.upload(xs)
        for f in range(100):
            s.step(p)
            if f + 1 in (50, 100):
                s.sync()
                got[f + 1] = s.download()
        g = got[100]
        rho_g = s.tap(capi.TAP_RHO)
    t, ref = None, {}
    for f in range(100):
        t = oracle_mod.step(H, p, cpu, taps=(f == 99))
        oracle_mod.step(H, p, nudged)
        if f + 1 in (50, 100):
            ref[f + 1] = (ke(cpu), ke(nudged))
    assert np.array_equal(np.sort(g["id"]), np.sort(cpu["id"]))
    dens_g, dens_c = float((rho_g / 6378.0 - 1).mean()), float((t["rho"] / 6378.0 - 1).mean())
    print(f"dam-64k after 100 steps: mean density error gpu {dens_g:+.4f} / oracle {dens_c:+.4f}")
    assert abs(dens_g - dens_c) <= 0.01, (dens_g, dens_c)
    for step in (50, 100):
        ke_c, ke_n = ref[step]
        band = max(0.10, 1.5 * abs(ke_n - ke_c) / ke_c)
        print(f"  step {step}: KE gpu {ke(got[step]):.1f} / oracle {ke_c:.1f} / nudged oracle {ke_n:.1f}  (band +-{band:.0%})")
        assert abs(ke(got[step]) - ke_c) <= band * ke_c, (step, ke(got[step]), ke_c, ke_n)
    lo, hi = np.array(p.min_bound[:]), np.array(p.max_bound[:])
    assert np.all(g["position"] >= lo - 1e-3) and np.all(g["position"] <= hi + 1e-3)
    # and one step from this warm state, float parity in both arithmetics (the figures DESIGN.md §2 quotes)
    snap = cpu.copy()
    ref1 = snap.copy()
    oracle_mod.step(H, p, ref1)
    for flags, tol in ((0, 1e-5), (FLAG_STRICT_FP, 1e-7)):
        out, _, _ = run_gpu(p, snap, flags, taps=False)
        dp = np.abs(out["position"].astype(np.float64) - ref1["position"])
        print(f"dam-64k one step flags={flags}: max|dx| = {dp.max():.3e} ({dp.max() / box_edge(p):.2e} of the domain)")
        assert np.array_equal(out["id"], ref1["id"])
        assert frac_within(out["position"], ref1["position"], tol * box_edge(p)) >= 0.999
        assert (dp.max(axis=1) > 10 * tol * box_edge(p)).sum() <= max(1, 1.5e-4 * len(dp))
