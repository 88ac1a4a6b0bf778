"""The oracle against the real reference build (oracle/_ref), when that build is present (it is built wherever
/root/reference exists and travels prebuilt to the GPU box; nothing here reads /root/reference at run time)."""
import numpy as np
import pytest

H = 0.1


def _need(oracle_mod, variant):
    if not oracle_mod.ref_available(variant):
        pytest.skip(f"oracle/_ref/libpbf_ref_{variant}.so not built here")
    try:
        oracle_mod.ref_lib(variant)
    except OSError as e:
        pytest.skip(f"prebuilt reference library not loadable on this host: {e}")


def test_gauss_seidel_mode_equals_reference_with_stable_sort(oracle_mod):
    _need(oracle_mod, "strict_stable")
    p, xs = oracle_mod.ref_scene_2cubes(6000, 4)
    p.surface_enabled = 1
    a, b = xs.copy(), xs.copy()
    for frame in range(4):
        pf = oracle_mod.ref_apply_motion(p, frame)
        r = oracle_mod.ref_advance(H, pf, a, variant="strict_stable", threads=1, mesh_cap=400000)
        o = oracle_mod.step(H, pf, b, mode=oracle_mod.GAUSS_SEIDEL)
        assert a.tobytes() == b.tobytes(), f"frame {frame}: particles differ from the reference"
        assert r["n_vertices"] == o["n_vertices"]
        for k in ("mesh_vs", "mesh_ns", "mesh_cs"):
            assert np.array_equal(r[k], o[k], equal_nan=True), k


def test_scene_dynamics_equal_reference(oracle_mod):
    """Wells, sources, drains and queries (sph.hpp:56-80, ompsph.hpp:91-120,141-148,167-186): the oracle's restatement
    against advance(config, scene, xs) of the real reference, several calls so that emitted particles get drained,
    pulled by the well and queried."""
    _need(oracle_mod, "strict_stable")
    from helpers import demo_scene
    sc = demo_scene()
    p, xs = oracle_mod.ref_scene_2cubes(2000, 3)
    a, b = xs.copy(), xs.copy()
    saw_hit = False
    for call in range(5):
        a, answers = oracle_mod.ref_advance_scene(H, p, sc, a, variant="strict_stable", threads=1)
        b = oracle_mod.scene_edit(H, p, sc, b)
        t = oracle_mod.step(H, p, b, mode=oracle_mod.GAUSS_SEIDEL, taps=True, scene=sc)
        assert a.tobytes() == b.tobytes(), f"call {call}: particles differ from the reference"
        for (qid, ids), first, cnt in zip(answers, t["query_first"], t["query_count"]):
            assert np.array_equal(ids, b["id"][first:first + cnt]), f"call {call}: query {qid}"
            saw_hit |= len(ids) > 0
    assert saw_hit and (b["id"] == 99).sum() > 0 and len(b) != len(xs)  # the scene did something


def test_unmodified_reference_via_recovered_permutation(oracle_mod):
    _need(oracle_mod, "strict")
    p, xs = oracle_mod.ref_scene_2cubes(6000, 4)
    a = xs.copy()
    for frame in range(3):
        pf = oracle_mod.ref_apply_motion(p, frame)
        before = a.copy()
        oracle_mod.ref_advance(H, pf, a, variant="strict", threads=1)
        where = np.empty(len(before), np.int64)
        where[before["id"]] = np.arange(len(before))
        b = before.copy()
        oracle_mod.step(H, pf, b, mode=oracle_mod.GAUSS_SEIDEL, forced_perm=where[a["id"]].astype(np.uint32))
        assert a.tobytes() == b.tobytes()


def test_scene_factory_motion_constants_and_morton_match_reference(oracle_mod):
    _need(oracle_mod, "strict")
    import ctypes as C
    from pbf_sph_b200 import capi, scenes
    pr, xr = oracle_mod.ref_scene_2cubes(20000, 6)
    p, xs = scenes.two_cubes(20000, 6)
    assert xs.tobytes() == xr.tobytes() and bytes(p) == bytes(pr)
    for f in (0, 1, 19, 63, 199):
        assert bytes(scenes.apply_motion(p, f)) == bytes(oracle_mod.ref_apply_motion(pr, f))
    R = oracle_mod.ref_lib("strict")
    rng = np.random.default_rng(0)
    for x, y, z in rng.integers(0, 1024, (2000, 3)):
        k = R.pbf_ref_morton_encode(int(x), int(y), int(z))
        assert k == capi.lib().pbf_host_morton_encode(int(x), int(y), int(z)) == \
            oracle_mod.lib().pbf_oracle_morton_encode(int(x), int(y), int(z))
    ref_c = np.zeros(3, np.float32)
    R.pbf_ref_constants(C.c_float(H), ref_c.ctypes.data_as(C.c_void_p))
    ours = np.zeros(5, np.float32)
    capi.lib().pbf_host_constants(C.c_float(H), ours.ctypes.data)
    assert np.array_equal(ref_c, ours[:3])
