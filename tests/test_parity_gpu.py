"""GPU parity: libpbf_cuda.so (through the C ABI) against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star):
  * bit-exact: Morton keys, stable sort permutation, sorted keys, cell table, candidate and in-radius counts,
    ids and their output order, diffused colours (all-strict arithmetic in the same summation order);
  * floats after ONE step from a WARM snapshot (never from the t=0 lattice — SURVEY F4): >= 99.9 % of position
    components within 1e-5 * domain (domain = 1000 units) and all within 1e-4 * domain; |dv| <= 1e-2;
  * with PBF_FLAG_STRICT_FP the solver follows the oracle op for op (only pow(x,4) differs by design), so the
    tolerance is 100x tighter there;
  * multi-step runs: aggregate statistics (mean density error, kinetic energy), because the dynamics are chaotic.
"""
import numpy as np
import pytest

from helpers import DOMAIN, by_id, frac_within, warm
from pbf_sph_b200 import FLAG_DEBUG_COUNTS, FLAG_GLOBAL_NEIGHBOURS, FLAG_STRICT_FP, Solver, capi, scenes

pytestmark = pytest.mark.gpu
H = scenes.H


def run_gpu(params, xs, flags=0, taps=True, list_cap=None):
    out = xs.copy()
    with Solver(H, 0, flags | (FLAG_DEBUG_COUNTS if taps else 0)) as s:
        if list_cap is not None:
            s.set_list_capacity(list_cap)
        res = s.advance(params, out)
        t = {}
        if taps:
            for name, tap in (("keys_input", capi.TAP_KEYS_INPUT), ("perm", capi.TAP_PERM),
                              ("keys_sorted", capi.TAP_KEYS_SORTED), ("cell_table", capi.TAP_CELL_TABLE),
                              ("cand_count", capi.TAP_CAND_COUNT), ("nbr_count", capi.TAP_NBR_COUNT),
                              ("lambda", capi.TAP_LAMBDA), ("rho", capi.TAP_RHO)):
                t[name] = s.tap(tap)
            if not flags & FLAG_GLOBAL_NEIGHBOURS:
                t["list_hits"] = s.tap(capi.TAP_LIST_HITS)
            t["grid"] = s.grid()
    return out, t, res


INT_TAPS = ("keys_input", "perm", "keys_sorted", "cell_table", "cand_count", "nbr_count")


def assert_integer_parity(t_gpu, t_cpu, flags=None):
    """Bit-exact integer artefacts.  With `flags` given, the hit counts of the PRODUCTION neighbour list
    (PBF_TAP_LIST_HITS: what the lambda pass of the first solver iteration appended, neighbour_list.cu) are held to the
    oracle's in-radius counts as well: exactly under STRICT_FP; in the production arithmetic the distance test may be
    FMA-contracted, so a pair within an ulp of r = h can fall on the other side — at most 1 particle in 10^5 may differ,
    and by one neighbour."""
    for k in INT_TAPS:
        assert np.array_equal(t_gpu[k], t_cpu[k]), f"{k} differs from the oracle"
    if flags is not None and "list_hits" in t_gpu:
        diff = t_gpu["list_hits"].astype(np.int64) - t_cpu["nbr_count"].astype(np.int64)
        bad = int(np.count_nonzero(diff))
        if flags & FLAG_STRICT_FP:
            assert bad == 0, f"production hit list differs from the oracle's neighbour counts on {bad} particles"
        else:
            assert bad <= max(1, len(diff) // 100000) and (bad == 0 or np.abs(diff).max() <= 1), \
                f"production hit list: {bad} particles differ (max {np.abs(diff).max()})"
        return bad
    return 0


@pytest.mark.parametrize("flags", [0, FLAG_STRICT_FP, FLAG_GLOBAL_NEIGHBOURS, FLAG_GLOBAL_NEIGHBOURS | FLAG_STRICT_FP])
def test_t0_integers_stock_scene(gpu, oracle_mod, flags):
    """S0 at t=0 with the moving wall: every integer artefact is bit-exact (floats are NOT compared at t=0)."""
    p, xs = scenes.two_cubes(20000, 6)
    for frame in (0, 3):
        pf = scenes.apply_motion(p, frame)
        cpu = xs.copy()
        t_cpu = oracle_mod.step(H, pf, cpu, taps=True)
        gpu_xs, t_gpu, _ = run_gpu(pf, xs, flags)
        assert list(t_gpu["grid"].extent) == list(t_cpu["grid"].extent)
        assert t_gpu["grid"].grid_table_n == t_cpu["grid"].grid_table_n
        assert_integer_parity(t_gpu, t_cpu, flags)
        assert np.array_equal(gpu_xs["id"], cpu["id"]), "output order (Z-sorted, ids carried) differs"
        assert np.array_equal(gpu_xs["colour"], cpu["colour"]), "diffused colours must be bit-exact"
        assert np.array_equal(gpu_xs["mass"], cpu["mass"]) and np.all(gpu_xs["type"] == 0)


@pytest.fixture(scope="module")
def warm_s0(oracle_mod):
    p, xs = scenes.two_cubes(20000, 6)
    warm(oracle_mod, H, p, xs, 40, motion=scenes.apply_motion)
    return p, xs


@pytest.mark.parametrize("flags,pos_tol,vel_tol", [
    (0, 1e-5 * DOMAIN, 1e-2),
    (FLAG_GLOBAL_NEIGHBOURS, 1e-5 * DOMAIN, 1e-2),
    (FLAG_STRICT_FP, 1e-7 * DOMAIN, 1e-4),
    (FLAG_GLOBAL_NEIGHBOURS | FLAG_STRICT_FP, 1e-7 * DOMAIN, 1e-4),
])
def test_one_step_from_warm_snapshot(gpu, oracle_mod, warm_s0, flags, pos_tol, vel_tol):
    p, snap = warm_s0
    pf = scenes.apply_motion(p, 40)
    cpu = snap.copy()
    t_cpu = oracle_mod.step(H, pf, cpu, taps=True)
    gpu_xs, t_gpu, _ = run_gpu(pf, snap, flags)
    assert_integer_parity(t_gpu, t_cpu, flags)
    assert np.array_equal(gpu_xs["id"], cpu["id"])
    assert np.array_equal(gpu_xs["colour"], cpu["colour"])
    dp = np.abs(gpu_xs["position"].astype(np.float64) - cpu["position"])
    dv = np.abs(gpu_xs["velocity"].astype(np.float64) - cpu["velocity"])
    assert frac_within(gpu_xs["position"], cpu["position"], pos_tol) >= 0.999, dp.max()
    assert dp.max() <= 10 * pos_tol, dp.max()
    assert dv.max() <= vel_tol * max(1.0, np.abs(cpu["velocity"]).max() / 100.0), dv.max()
    lam_scale = np.abs(t_cpu["lambda"]).max()
    assert np.abs(t_gpu["lambda"] - t_cpu["lambda"]).max() <= (1e-6 if flags & FLAG_STRICT_FP else 1e-3) * lam_scale
    assert np.allclose(t_gpu["rho"], t_cpu["rho"], rtol=1e-6 if flags & FLAG_STRICT_FP else 1e-4, atol=1e-2)


def test_multi_step_aggregates(gpu, oracle_mod):
    """100 resident GPU steps vs 100 oracle steps: mean density error and kinetic energy (SURVEY §8c bands)."""
    p, xs = scenes.two_cubes(20000, 6)
    cpu = xs.copy()
    with Solver(H, 0) as s:
        s.upload(xs)
        for f in range(100):
            s.step(scenes.apply_motion(p, f))
        s.sync()
        g = s.download()
        rho_g = s.tap(capi.TAP_RHO)
    t = None
    for f in range(100):
        t = oracle_mod.step(H, scenes.apply_motion(p, f), cpu, taps=(f == 99))
    assert sorted(g["id"].tolist()) == sorted(cpu["id"].tolist())
    dens_g, dens_c = float((rho_g / 6378.0 - 1).mean()), float((t["rho"] / 6378.0 - 1).mean())
    ke_g = 0.5 * float((g["velocity"].astype(np.float64) ** 2).sum())
    ke_c = 0.5 * float((cpu["velocity"].astype(np.float64) ** 2).sum())
    assert abs(dens_g - dens_c) <= 0.01, (dens_g, dens_c)
    assert abs(ke_g - ke_c) <= 0.10 * ke_c, (ke_g, ke_c)
    lo, hi = np.array(p.min_bound[:]), np.array(p.max_bound[:])
    assert np.all(np.isfinite(g["position"])) and np.all(np.isfinite(g["velocity"]))


@pytest.mark.parametrize("n", [0, 1, 2, 31, 255, 256, 257, 4095, 4096, 4097])
def test_ragged_sizes(gpu, oracle_mod, n):
    """Empty, single-particle and tile-boundary sizes (sort tile = 4096, AoS staging block = 256)."""
    p, xs = scenes.two_cubes(20000, 3)
    xs = xs[:n].copy()
    cpu = xs.copy()
    t_cpu = oracle_mod.step(H, p, cpu, taps=True)
    if n == 0:
        with Solver(H, 0) as s:
            s.advance(p, xs)  # ompsph.hpp:122-126: nothing to do, no error
        return
    gpu_xs, t_gpu, _ = run_gpu(p, xs, FLAG_STRICT_FP)
    assert_integer_parity(t_gpu, t_cpu)
    assert np.array_equal(gpu_xs["id"], cpu["id"])


def test_particles_outside_the_grid(gpu, oracle_mod):
    """Fast particles leave the padded grid (key >= G or wrapped cell): still sorted on the full key, still
    processed as `a`, invisible as neighbours — exactly like the reference (sph.hpp:203-213)."""
    p, xs = scenes.two_cubes(2000, 2)
    rng = np.random.default_rng(7)
    fast = rng.choice(len(xs), 64, replace=False)
    xs["velocity"][fast] = rng.uniform(-60.0, 60.0, (64, 3)).astype(np.float32)  # |v*dt| up to 0.75 > 2h padding
    xs["velocity"][fast, 1] = np.abs(xs["velocity"][fast, 1])
    cpu = xs.copy()
    t_cpu = oracle_mod.step(H, p, cpu, taps=True)
    assert (t_cpu["keys_input"] >= t_cpu["grid"].grid_table_n).any(), "scenario must produce out-of-grid keys"
    gpu_xs, t_gpu, _ = run_gpu(p, xs, FLAG_STRICT_FP)
    assert_integer_parity(t_gpu, t_cpu)
    assert np.array_equal(gpu_xs["id"], cpu["id"])
    assert np.allclose(gpu_xs["position"], cpu["position"], rtol=0, atol=1e-4 * DOMAIN)


def test_dense_cell_collision(gpu, oracle_mod):
    """Thousands of particles clamped onto one wall cell (SURVEY §7 'pathological cells')."""
    p, xs = scenes.two_cubes(20000, 2)
    xs = xs[:6000].copy()
    xs["position"][:3000] = (0.0, 1000.0, 0.0)  # all in one corner cell
    xs["position"][:3000] += np.random.default_rng(3).uniform(0, 5, (3000, 3)).astype(np.float32)
    cpu = xs.copy()
    t_cpu = oracle_mod.step(H, p, cpu, taps=True)
    assert t_cpu["cand_count"].max() >= 3000
    gpu_xs, t_gpu, _ = run_gpu(p, xs, 0)
    assert_integer_parity(t_gpu, t_cpu)
    assert np.array_equal(gpu_xs["id"], cpu["id"])


@pytest.mark.parametrize("flags", [0, FLAG_STRICT_FP])
def test_piled_particles_wide_list_against_oracle_and_narrow_list(gpu, oracle_mod, flags):
    """While a dam break splashes, particles clamped onto the walls pile up: 97..229 in-radius neighbours at 1 M
    particles.  Clumps of 110 particles (more hits than the 96-deep list of the A/B modes, fewer than the production
    192) take the list path by default and the one-pass fallback with a 96-deep list (pbf_debug_set_list_capacity): same neighbour sets, same
    summation order, so lambda and the step agree with the oracle either way, and bit for bit with each other in the
    strict arithmetic."""
    rng = np.random.default_rng(11)
    p, xs = scenes.two_cubes(6000, 3)
    for k in range(6):  # six clumps of 110 particles, each inside a cube of 0.4 h (= 20 unscaled units)
        sel = slice(110 * k, 110 * (k + 1))
        centre = xs["position"][110 * k].copy()
        xs["position"][sel] = centre + rng.uniform(-10.0, 10.0, (110, 3)).astype(np.float32)
    cpu = xs.copy()
    t_cpu = oracle_mod.step(H, p, cpu, taps=True)
    assert 96 < t_cpu["nbr_count"].max() <= 192
    runs = {}
    for cap in (192, 96):
        runs[cap] = run_gpu(p, xs, flags, list_cap=cap)
        gpu_xs, t_gpu, _ = runs[cap]
        assert_integer_parity(t_gpu, t_cpu)
        tol = 1e-6 if flags & FLAG_STRICT_FP else 1e-4
        assert np.allclose(t_gpu["lambda"], t_cpu["lambda"], rtol=tol, atol=tol * np.abs(t_cpu["lambda"]).max()), cap
        assert np.allclose(t_gpu["rho"], t_cpu["rho"], rtol=tol, atol=tol * np.abs(t_cpu["rho"]).max()), cap
    if flags & FLAG_STRICT_FP:
        assert np.array_equal(runs[192][1]["lambda"], runs[96][1]["lambda"])
        assert runs[192][0].tobytes() == runs[96][0].tobytes()


def test_obstacle_rejected(gpu):
    p, xs = scenes.two_cubes(2000, 2)
    xs["type"][5] = 1
    before = xs.copy()
    with Solver(H, 0) as s:
        with pytest.raises(capi.PbfError):
            s.advance(p, xs)
    assert xs.tobytes() == before.tobytes(), "a rejected call must leave the caller's particles untouched"


def test_dam_break_parity_and_full_size_properties(gpu, oracle_mod):
    """dam(40) = 64 000 particles (BASELINE config 1): integer parity at t=0; then size-independent properties at
    the full 1 M size of config 2: sortedness, permutation validity, table = lower_bound, id conservation."""
    p, xs = scenes.dam_break(40, 4)
    cpu = xs.copy()
    t_cpu = oracle_mod.step(H, p, cpu, taps=True)
    gpu_xs, t_gpu, _ = run_gpu(p, xs, 0)
    assert_integer_parity(t_gpu, t_cpu)
    assert np.array_equal(gpu_xs["id"], cpu["id"])

    p, xs = scenes.dam_break(100, 4)
    with Solver(H, 0) as s:
        s.upload(xs)
        for _ in range(3):
            s.step(p)
        s.sync()
        keys, perm, table = s.tap(capi.TAP_KEYS_SORTED), s.tap(capi.TAP_PERM), s.tap(capi.TAP_CELL_TABLE)
        out = s.download()
    assert np.all(np.diff(keys.astype(np.int64)) >= 0), "keys must be sorted"
    assert np.array_equal(np.sort(perm), np.arange(len(xs), dtype=np.uint32)), "perm must be a permutation"
    assert np.array_equal(table, np.searchsorted(keys, np.arange(len(table)), side="left").astype(np.uint32))
    assert np.array_equal(np.sort(out["id"]), np.arange(len(xs), dtype=np.uint64)), "ids are conserved"
    lo, hi = np.array(p.min_bound[:]), np.array(p.max_bound[:])
    assert np.all(out["position"] >= lo - 1e-3) and np.all(out["position"] <= hi + 1e-3), "clamped to the box"


@pytest.mark.parametrize("seed,n,flags", [(1, 5000, FLAG_STRICT_FP), (2, 12345, 0), (3, 777, FLAG_STRICT_FP), (4, 40000, 0)])
def test_random_clouds(gpu, oracle_mod, seed, n, flags):
    """Not a lattice: uniformly random positions in a slab of the box, random velocities (some fast enough to leave the
    padded grid), random masses and colours, ids in random order.  Integer artefacts bit-exact, lambda and the step
    within the float tolerance, for both list depths."""
    rng = np.random.default_rng(seed)
    p, _ = scenes.two_cubes(2000, 3)
    xs = np.zeros(n, capi.PARTICLE)
    xs["id"] = rng.permutation(n).astype(np.uint64)
    xs["mass"] = rng.uniform(0.5, 1.5, n).astype(np.float32)
    side = 1000.0 * min(1.0, (n / 60000.0) ** (1 / 3))  # ~ the stock scene's density
    xs["position"] = rng.uniform(0.0, side, (n, 3)).astype(np.float32)
    xs["velocity"] = rng.normal(0.0, 2.0, (n, 3)).astype(np.float32)
    xs["velocity"][: n // 200] *= 20.0
    xs["colour"] = rng.uniform(0.0, 1.0, (n, 4)).astype(np.float32)
    cpu = xs.copy()
    t_cpu = oracle_mod.step(H, p, cpu, taps=True)
    for list_cap in (192, 96):
        gpu_xs, t_gpu, _ = run_gpu(p, xs, flags, list_cap=list_cap)
        assert_integer_parity(t_gpu, t_cpu, flags if list_cap == 192 else None)
        assert np.array_equal(gpu_xs["id"], cpu["id"]) and np.array_equal(gpu_xs["colour"], cpu["colour"])
        tol = 1e-6 if flags & FLAG_STRICT_FP else 1e-4
        assert np.allclose(t_gpu["lambda"], t_cpu["lambda"], rtol=tol, atol=tol * np.abs(t_cpu["lambda"]).max())
        assert np.allclose(t_gpu["rho"], t_cpu["rho"], rtol=tol, atol=tol * np.abs(t_cpu["rho"]).max())
        # a random cloud is as ill-conditioned as the t=0 lattice: positions only loosely in the fast mode
        pos_tol = (1e-6 if flags & FLAG_STRICT_FP else 1e-3) * DOMAIN
        assert frac_within(gpu_xs["position"], cpu["position"], pos_tol) >= 0.995


def test_advance_from_a_worker_thread(gpu):
    """visualise.cpp:85-109 calls advance() from a worker thread while the UI thread owns the process: every C entry
    point sets the device itself, so a context created on one thread works from another (serial calls)."""
    import threading
    p, xs = scenes.two_cubes(4000, 3)
    main = xs.copy()
    with Solver(H, 0) as s:
        for f in range(3):
            s.advance(scenes.apply_motion(p, f), main)
    worker, err = xs.copy(), []
    with Solver(H, 0) as s:  # created here ...
        def run():
            try:
                for f in range(3):
                    s.advance(scenes.apply_motion(p, f), worker)  # ... driven there
            except Exception as e:  # pragma: no cover
                err.append(e)
        t = threading.Thread(target=run)
        t.start()
        t.join()
    assert not err and worker.tobytes() == main.tobytes()
