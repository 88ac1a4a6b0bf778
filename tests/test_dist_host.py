"""Host-side logic of the multi-GPU path, on CPU: split planning (pbf_host_plan_splits) and, with torch.distributed
over gloo at world_size 2, the launcher-side plumbing bench.py / dist.py rely on (id broadcast, sharding, max-over-ranks
timing, histogram all-reduce -> identical splits on every rank)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from pbf_sph_b200 import capi, scenes
from pbf_sph_b200.dist import shard


def plan(hist, shift, world):
    h = np.ascontiguousarray(hist, dtype=np.uint64)
    out = np.zeros(world + 1, np.uint32)
    rc = capi.lib().pbf_host_plan_splits(h.ctypes.data, len(h), shift, world, out.ctypes.data)
    assert rc == 0
    return out


def test_plan_splits_balances_and_tiles_key_space():
    rng = np.random.default_rng(7)
    hist = rng.integers(0, 50, 4096)
    hist[1000:1400] = 0  # an empty stretch of the curve
    for world in (1, 2, 3, 4, 8):
        sp = plan(hist, 3, world)
        assert sp[0] == 0 and sp[-1] == 1 << 30
        assert np.all(np.diff(sp.astype(np.int64)) >= 0)
        assert np.all(sp[1:-1] % 8 == 0)  # splits fall on bucket boundaries (2x2x2 cell blocks for shift 3)
        cum = np.concatenate([[0], np.cumsum(hist)])
        owned = [cum[min(len(hist), sp[r + 1] >> 3)] - cum[min(len(hist), sp[r] >> 3)] for r in range(world)]
        assert sum(owned) == hist.sum()
        assert max(owned) <= hist.sum() / world + hist.max()


def test_work_weights_move_the_split_towards_the_dense_half():
    """The slab splits balance density-weighted work (pbf_host_work_weights), not particle counts: with a dense and a
    sparse half holding the same number of particles, the dense half is the heavier one, so the split moves into it."""
    shift, per_bucket = 3, 8
    hist = np.zeros(513, np.uint32)          # 512 buckets of 8 cells + the bucket of everything >= G
    hist[:128] = 12 * per_bucket             # dense quarter: 12 particles per cell
    hist[128:512] = 4 * per_bucket           # sparse three quarters: 4 per cell — the same particle count in total
    hist[512] = 50                           # particles outside the grid: counted, but no density
    w = np.zeros(len(hist), np.uint64)
    assert capi.lib().pbf_host_work_weights(hist.ctypes.data, len(hist), shift, w.ctypes.data) == 0
    slots = 1 << shift
    assert np.array_equal(w[:512], hist[:512].astype(np.uint64) * (np.uint64(int(4.4 * slots)) + hist[:512].astype(np.uint64)))
    assert w[512] == 50 * int(4.4 * slots)
    by_count = plan(hist.astype(np.uint64), shift, 2)[1] >> shift
    by_work = plan(w, shift, 2)[1] >> shift
    assert by_count == 129                   # half of the particles: the dense quarter (+ one bucket of rounding)
    assert by_work < by_count                # half of the WORK is reached earlier, inside the dense quarter
    dense_cost, sparse_cost = 128 * 96 * (35 + 96), 384 * 32 * (35 + 32)
    assert abs(int(w[:by_work].sum()) - (dense_cost + sparse_cost + int(w[512])) / 2) <= w[:512].max()
    assert capi.lib().pbf_host_work_weights(None, 4, 3, w.ctypes.data) != 0


def test_plan_splits_degenerate_inputs():
    assert list(plan([0, 0, 0, 0], 0, 2)) == [0, 0, 1 << 30]            # no particles at all
    assert list(plan([10, 0, 0, 0], 4, 4)) == [0, 16, 16, 16, 1 << 30]  # everything in one bucket
    assert capi.lib().pbf_host_plan_splits(None, 4, 0, 2, None) != 0


def test_shard_partitions_the_input():
    _, xs = scenes.two_cubes(2000, 2)
    for world in (1, 2, 3, 8):
        parts = [shard(xs, r, world) for r in range(world)]
        assert np.array_equal(np.concatenate(parts)["id"], xs["id"])
        assert max(len(q) for q in parts) - min(len(q) for q in parts) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the 128-byte id blob travels from rank 0 to everyone (bench: NCCL unique id)
        blob = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            blob.copy_(torch.arange(capi.NCCL_ID_BYTES, dtype=torch.uint8))
        dist.broadcast(blob, 0)
        # 2. each rank holds its shard; a global coarse key histogram (all-reduce) gives every rank the same splits
        p, xs = scenes.dam_break(12, 2)
        mine = shard(xs, rank, world)
        g = capi.GridInfo()
        capi.lib().pbf_host_grid(C.c_float(scenes.H), C.byref(p), C.byref(g))
        cell = np.floor((mine["position"] / p.scale - np.array(g.min_extent[:], np.float32)) / np.float32(scenes.H)).astype(np.uint32)
        keys = np.array([capi.lib().pbf_host_morton_encode(int(x), int(y), int(z)) for x, y, z in cell], np.uint32)
        shift, nb = 3, (g.grid_table_n >> 3) + 2
        hist = torch.from_numpy(np.bincount(np.minimum(keys >> shift, nb - 1), minlength=nb).astype(np.int64))
        dist.all_reduce(hist)
        sp = plan(hist.numpy(), shift, world)
        owner = np.searchsorted(sp[1:-1], keys, side="right")
        sent = torch.tensor([int((owner == d).sum()) for d in range(world)])
        got = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(got, sent)
        # 3. max-over-ranks timing reduction
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, bytes(blob.numpy().tobytes()), sp.tolist(), torch.stack(got).numpy().tolist(), float(t.item()), len(xs)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_launcher_plumbing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q, port, world = ctx.Queue(), _free_port(), 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    (_, blob0, sp0, m0, t0, n), (_, blob1, sp1, m1, t1, _) = res
    assert blob0 == blob1 == bytes(range(capi.NCCL_ID_BYTES))
    assert sp0 == sp1 and sp0[0] == 0 and sp0[-1] == 1 << 30
    assert m0 == m1                      # both ranks see the same migration matrix
    m = np.array(m0)
    assert m.sum() == n                  # every particle has exactly one owner
    owned = m.sum(axis=0)
    assert abs(int(owned[0]) - int(owned[1])) <= 0.2 * n
    assert t0 == t1 == 2.0
