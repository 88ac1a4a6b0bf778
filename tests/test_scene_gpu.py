"""Scene dynamics on the GPU (wells, sources, drains, queries — sph.hpp:56-80) through the C ABI against the CPU
oracle, which tests/test_oracle_vs_reference.py and tests/golden/scene_2cubes.npz pin to the real reference."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from helpers import DOMAIN, demo_scene, frac_within
from pbf_sph_b200 import FLAG_STRICT_FP, Solver, capi, scenes

pytestmark = pytest.mark.gpu
H = scenes.H
GOLD = Path(__file__).resolve().parent / "golden"


def oracle_call(oracle_mod, p, sc, xs):
    xs = oracle_mod.scene_edit(H, p, sc, xs)
    t = oracle_mod.step(H, p, xs, taps=True, scene=sc)
    answers = [(q.id, xs["id"][f:f + c].copy()) for q, f, c in zip(sc.queries, t["query_first"], t["query_count"])]
    return xs, answers, t


@pytest.mark.parametrize("flags,pos_tol,vel_tol", [(FLAG_STRICT_FP, 1e-7 * DOMAIN, 1e-4), (0, 1e-5 * DOMAIN, 1e-2)])
def test_advance_with_scene_matches_oracle(gpu, oracle_mod, flags, pos_tol, vel_tol):
    """Drop-in advance(config, scene, xs), call after call from the same inputs: particle count, output order, ids and
    query answers exact; the emitted sheet, the drained corner and the well's pull to the float tolerance of the
    parity tests (positions 1e-5 of the domain in the production arithmetic, 1e-7 following the oracle op for op)."""
    sc = demo_scene()
    p, xs = scenes.two_cubes(2000, 3)
    cpu = xs.copy()
    with Solver(H, 0, flags) as s:
        for call in range(6):
            start = cpu.copy()  # both sides start every call from the oracle's state: no drift between calls
            cpu, want_q, t = oracle_call(oracle_mod, p, sc, cpu)
            got, res = s.advance_scene(p, sc, start)
            assert len(got) == len(cpu) and np.array_equal(got["id"], cpu["id"]), f"call {call}"
            for (qid, ids), (wid, wids) in zip(res.queries, want_q):
                assert qid == wid and np.array_equal(ids, wids), f"call {call}: query {qid}"
            assert np.array_equal(got["colour"], cpu["colour"])
            if call >= 3:  # floats away from the t=0 lattice only
                assert frac_within(got["position"], cpu["position"], pos_tol) >= 0.999
                assert np.abs(got["position"] - cpu["position"]).max() <= 10 * pos_tol
                assert np.abs(got["velocity"] - cpu["velocity"]).max() <= vel_tol
    assert (cpu["id"] == 99).sum() > 0 and any(len(ids) for _, ids in want_q)


def test_keys_with_a_well_are_bit_exact(gpu, oracle_mod):
    """Wells act inside the prediction (ompsph.hpp:141-148), i.e. on the Morton keys: after the wells have pulled for a
    dozen steps, keys, permutation and cell table of the next step are still bit-exact."""
    sc = capi.Scene(wells=[(1, (200.0, 100.0, 200.0), 9000.0), (2, (700.0, 150.0, 700.0), -4000.0)])
    p, xs = scenes.two_cubes(20000, 2)
    state = xs.copy()
    for _ in range(12):
        oracle_mod.step(H, p, state, scene=sc)
    cpu = state.copy()
    t = oracle_mod.step(H, p, cpu, taps=True, scene=sc)
    plain = oracle_mod.step(H, p, state.copy(), taps=True)
    assert not np.array_equal(t["keys_input"], plain["keys_input"]), "the wells must matter in this scenario"
    with Solver(H, 0) as s:
        got, _ = s.advance_scene(p, sc, state)
        assert np.array_equal(s.tap(capi.TAP_KEYS_INPUT), t["keys_input"])
        assert np.array_equal(s.tap(capi.TAP_PERM), t["perm"])
        assert np.array_equal(s.tap(capi.TAP_CELL_TABLE), t["cell_table"])
    assert np.array_equal(got["id"], cpu["id"])
    assert frac_within(got["position"], cpu["position"], 1e-5 * DOMAIN) >= 0.999


def test_resident_path_applies_the_scene_every_step(gpu, oracle_mod):
    """pbf_set_scene + pbf_step: sources emit and drains remove on the resident state, like one advance() per step."""
    sc = demo_scene()
    p, xs = scenes.two_cubes(2000, 3)
    cpu = xs.copy()
    with Solver(H, 0, FLAG_STRICT_FP) as s:
        s.upload(xs)
        s.set_scene(sc)
        for step in range(8):
            cpu, want_q, _ = oracle_call(oracle_mod, p, sc, cpu)
            s.step(p)
            s.sync()
            assert s.count() == len(cpu), f"step {step}"
        got = s.download()
        answers = s.query_results(sc)
        assert np.array_equal(got["id"], cpu["id"])
        for (qid, ids), (wid, wids) in zip(answers, want_q):
            assert qid == wid and np.array_equal(ids, wids)
        assert frac_within(got["position"], cpu["position"], 1e-4 * DOMAIN) >= 0.995  # 8 chaotic steps apart
        s.set_scene(None)
        n = s.count()
        s.step(p)
        s.sync()
        assert s.count() == n, "an empty scene neither emits nor drains"


def test_everything_drained_and_capacity(gpu, oracle_mod):
    p, xs = scenes.two_cubes(2000, 2)
    everything = capi.Scene(drains=[(1, (500.0, 500.0, 500.0), 5000.0, 1.0)])
    with Solver(H, 0) as s:
        got, res = s.advance_scene(p, everything, xs)
        assert len(got) == 0 and res.queries == []  # "Particles depleted" (ompsph.hpp:122-126)
        # sources need room in the caller's array
        sc = demo_scene()
        buf = xs.copy()
        n_out, nv = C.c_uint64(0), C.c_uint64(0)
        rc = s._L.pbf_advance_scene_host(s._ctx, C.byref(p), C.byref(sc.struct), buf.ctypes.data, len(buf), len(buf),
                                         C.byref(n_out), C.byref(nv))
        assert rc == -5 and buf.tobytes() == xs.tobytes()  # PBF_ERR_CAPACITY, particles untouched


def test_scene_golden_from_the_reference(gpu, oracle_mod):
    """The GPU against the REAL reference's output (tests/golden/scene_2cubes.npz; Gauss-Seidel there, Jacobi here, so
    floats are compared loosely; counts, ids, order and query answers exactly)."""
    from test_oracle_golden import params_from
    g = np.load(GOLD / "scene_2cubes.npz")
    p = params_from(g["params"], oracle_mod)
    sc = demo_scene()
    with Solver(H, 0) as s:
        for call in (0, 5):
            got, res = s.advance_scene(p, sc, g[f"c{call}_in"].copy())
            want = g[f"c{call}_out"]
            assert len(got) == len(want) and np.array_equal(got["id"], want["id"])
            for qid, ids in res.queries:
                assert np.array_equal(ids, g[f"c{call}_q{qid}"])
            if call == 5:  # one Gauss-Seidel step (reference) vs one Jacobi step (device): same fluid, not the same bits
                assert frac_within(got["position"], want["position"], 1e-2 * DOMAIN) >= 0.99
