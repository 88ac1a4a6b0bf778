"""Slab-decomposed (multi-rank) step against the single-device step — include/pbf_cuda.h "multi-GPU", DESIGN.md §6.

The ranks run as contexts of one process on cuda:0 (LOCAL transport: the same phases, kernels and message layout
as the NCCL transport; only the copies differ), so this runs on a 1-GPU box.  Every rank uploads a consecutive block
of the input array; the merged stable sort then reproduces the single-device particle order cell by cell, ghosts
carry identical values, and every sum is formed in the same order: the result must be BIT-IDENTICAL to the
single-device path (which tests/test_parity_gpu.py pins to the oracle), for any number of ranks."""
import numpy as np
import pytest

from helpers import by_id
from pbf_sph_b200 import Solver, capi, scenes
from pbf_sph_b200.dist import LocalGroup

pytestmark = pytest.mark.gpu
H = scenes.H


def run_single(p, xs, frames, motion):
    out = []
    with Solver(H, 0) as s:
        s.upload(xs)
        for f in range(frames):
            s.step(scenes.apply_motion(p, f) if motion else p)
            out.append(s.download())
    return out


def run_group(p, xs, frames, motion, world, replan):
    out, stats = [], []
    with LocalGroup(H, [0] * world) as g:
        g.ranks[0].set_replan(replan)
        g.upload(xs)
        for f in range(frames):
            g.step(scenes.apply_motion(p, f) if motion else p)
            out.append(g.download())
            stats.append([r.stats() for r in g.ranks])
    return out, stats


@pytest.mark.parametrize("world,replan", [(2, 16), (3, 2), (4, 1), (8, 3)])
def test_slab_group_is_bit_identical_to_single_device(gpu, world, replan):
    p, xs = scenes.two_cubes(20000, 4)
    frames = 8
    ref = run_single(p, xs.copy(), frames, True)
    got, stats = run_group(p, xs.copy(), frames, True, world, replan)
    for f in range(frames):
        a, b = ref[f], got[f]
        assert len(a) == len(b)
        # concatenating the ranks gives the global Z order, i.e. exactly the single-device output order
        assert np.array_equal(a["id"], b["id"]), f"frame {f}: particle order differs"
        for field in ("position", "velocity", "colour", "mass"):
            assert np.array_equal(a[field].view(np.uint32), b[field].view(np.uint32)), f"frame {f}: {field} differs"
    last = stats[-1]
    assert sum(s["owned"] for s in last) == len(xs)
    # the decomposition really is distributed: every rank owns particles and ghosts flow
    assert all(s["owned"] > 0 for s in last)
    assert sum(s["ghosts"] for s in last) > 0
    assert any(any(s["migrants_in"] for s in st) for st in stats)
    # key ranges tile [0, 2^30)
    assert last[0]["key_lo"] == 0 and last[-1]["key_hi"] == 1 << 30
    for a, b in zip(last[:-1], last[1:]):
        assert a["key_hi"] == b["key_lo"]


def test_slab_group_dam_break_balance_and_parity(gpu):
    p, xs = scenes.dam_break(40, 4)
    frames = 6
    ref = run_single(p, xs.copy(), frames, False)
    got, stats = run_group(p, xs.copy(), frames, False, 4, 2)
    a, b = ref[-1], got[-1]
    assert np.array_equal(a["id"], b["id"])
    assert np.array_equal(a["position"].view(np.uint32), b["position"].view(np.uint32))
    owned = [s["owned"] for s in stats[-1]]
    assert max(owned) < 1.4 * len(xs) / 4, owned  # histogram splits balance the (density-weighted) particle counts
    ring1 = sum(s["ghost_ring1"] for s in stats[-1])
    ghosts = sum(s["ghosts"] for s in stats[-1])
    assert 0 < ring1 < ghosts  # lambda is computed for the inner ghost ring only


def test_slab_rank_with_uneven_and_empty_uploads(gpu):
    """Any initial distribution is legal: here rank 0 uploads everything and rank 1 nothing."""
    p, xs = scenes.two_cubes(4000, 3)
    ref = run_single(p, xs.copy(), 3, False)
    with LocalGroup(H, [0, 0]) as g:
        g.ranks[0].upload(xs)
        g.ranks[1].upload(xs[:0])
        for _ in range(3):
            g.step(p)
        out = g.download()
    assert np.array_equal(by_id(ref[-1])["position"].view(np.uint32), by_id(out)["position"].view(np.uint32))


def test_slab_path_rejects_single_device_calls(gpu):
    p, xs = scenes.two_cubes(2000, 2)
    with LocalGroup(H, [0, 0]) as g:
        g.upload(xs)
        with pytest.raises(capi.PbfError):
            g.solvers[0].step(p)  # pbf_step on a slab rank


@pytest.mark.parametrize("world", [2, 4])
def test_group_advance_equals_single_device_advance(gpu, world):
    """pbf_dist_advance_host (what sph::cuda_impl::Solver(h, {devices...})::advance calls): the caller's array goes in,
    comes back in the global Z order — byte for byte what pbf_advance_host returns on one device, frame after frame
    (the array returned by one call is the next call's input, as in the reference's drivers, benchmark.cpp:22-58)."""
    p, xs = scenes.two_cubes(20000, 4)
    one, many = xs.copy(), xs.copy()
    with Solver(H, 0) as s, LocalGroup(H, [0] * world) as g:
        for f in range(6):
            pf = scenes.apply_motion(p, f)
            s.advance(pf, one)
            g.advance(pf, many)
            assert many.tobytes() == one.tobytes(), f"frame {f}"
        # a different array size mid-run (the caller dropped particles): blocks are re-cut, still identical
        one, many = one[:9000].copy(), many[:9000].copy()
        s.advance(p, one)
        g.advance(p, many)
        assert many.tobytes() == one.tobytes()
    xs["type"][7] = 1
    with LocalGroup(H, [0] * world) as g, pytest.raises(capi.PbfError):
        g.advance(p, xs)


def test_group_on_distinct_devices(gpu):
    """One process driving two different GPUs (pbf_dist_init_local with distinct ordinals): per-device kernel attributes
    (the tiled diffusion opts in to > 48 KB of shared memory per DEVICE) and peer copies."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p, xs = scenes.two_cubes(20000, 4)
    ref = run_single(p, xs.copy(), 4, True)
    out = []
    with LocalGroup(H, [0, 1]) as g:
        g.upload(xs)
        for f in range(4):
            g.step(scenes.apply_motion(p, f))
        out = g.download()
    assert out.tobytes() == ref[-1].tobytes()


@pytest.mark.parametrize("world,replan", [(2, 4), (3, 1), (4, 2)])
def test_slab_group_surface_is_the_single_device_mesh(gpu, world, replan):
    """Marching cubes on the slab path (ompsph.hpp:277-477): lattice NaN pattern, triangle count and every vertex, normal
    and colour equal the single-device surface bit for bit, on the moving-wall scene (the lattice changes size between
    frames, which re-plans the arenas)."""
    p, xs = scenes.two_cubes(20000, 3)
    frames = 5
    ref = []
    with Solver(H, 0) as s:
        s.upload(xs.copy())
        for f in range(frames):
            pf = scenes.apply_motion(p, f)
            pf.surface_enabled = 1
            s.step(pf)
            ref.append((s.download(), s.mesh(), s.tap(capi.TAP_MC_FIELD).copy()))
    with LocalGroup(H, [0] * world) as g:
        g.ranks[0].set_replan(replan)
        g.upload(xs.copy())
        for f in range(frames):
            pf = scenes.apply_motion(p, f)
            pf.surface_enabled = 1
            g.step(pf)
            got, mesh = g.download(), g.mesh()
            a, m, field = ref[f]
            assert np.array_equal(a["id"], got["id"]) and np.array_equal(a["position"].view(np.uint32), got["position"].view(np.uint32))
            lattice = g.solvers[0].tap(capi.TAP_MC_FIELD)
            assert np.array_equal(field.view(np.uint32), lattice.view(np.uint32)), f"frame {f}: lattice differs"
            assert len(mesh.vs) == len(m.vs) and len(m.vs) > 0, f"frame {f}: {len(mesh.vs)} vertices, one device has {len(m.vs)}"
            for name in ("vs", "ns", "cs"):
                assert np.array_equal(getattr(m, name).view(np.uint32), getattr(mesh, name).view(np.uint32)), f"frame {f}: mesh.{name} differs"
            # only rank 0 holds the mesh
            assert all(s_.grid().n_triangles == 0 for s_ in g.solvers[1:])


def test_slab_group_advance_returns_the_mesh(gpu):
    """pbf_dist_advance_host (the multi-device sph::Solver::advance) with a surface: vertex count and mesh of one device."""
    p, xs = scenes.two_cubes(20000, 3)
    p.surface_enabled = 1
    one = xs.copy()
    with Solver(H, 0) as s:
        res = s.advance(p, one)
    many = xs.copy()
    with LocalGroup(H, [0, 0, 0]) as g:
        nv = g.advance(p, many)
        mesh = g.mesh()
    assert nv == len(res.vs) and nv > 0
    assert np.array_equal(one["id"], many["id"])
    assert np.array_equal(res.vs.view(np.uint32), mesh.vs.view(np.uint32))


def test_slab_group_replans_early_when_a_capacity_fills(gpu):
    """No scheduled re-plan at all (pbf_dist_set_replan(0): the arenas are sized once, on the first step, from the block of
    the dam break standing in its corner).  The dam then collapses across the tank: the owned and ghost counts of every key
    range drift far from what the first plan sized for.  The ranks must notice from the count matrices (SlabDyn::hot),
    re-plan together without talking, and stay bit-identical to one device."""
    p, xs = scenes.dam_break(64, 3)  # 262 144 particles: the fixed slack of a capacity (4 096) is small beside a rank's share
    frames = 90
    with Solver(H, 0) as s:
        s.upload(xs.copy())
        for f in range(frames):
            s.step(p)
        ref = s.download()
    with LocalGroup(H, [0] * 4) as g:
        g.ranks[0].set_replan(0)
        g.upload(xs.copy())
        first = None
        for f in range(frames):
            g.step(p)
            if f == 0:
                first = [r.stats() for r in g.ranks]
        got = g.download()
        last = [r.stats() for r in g.ranks]
    assert np.array_equal(ref["id"], got["id"])
    assert np.array_equal(ref["position"].view(np.uint32), got["position"].view(np.uint32))
    # the ranks re-planned although no re-plan was scheduled, all of them on the same steps
    assert first[0]["plan_steps"] == 1 and first[0]["early_plans"] == 0
    assert last[0]["early_plans"] >= 1, last
    assert len({(s["plan_steps"], s["early_plans"]) for s in last}) == 1
