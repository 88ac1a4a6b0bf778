"""Multi-GPU slab path over the C ABI (include/pbf_cuda.h "multi-GPU"): one rank per GPU, Z-curve slab decomposition.

* ``SlabRank``    one rank of an NCCL job (one process per GPU; torch.distributed carries the NCCL unique id).
* ``LocalGroup``  all ranks as contexts of this process (device-to-device copies) — the 1-GPU parity tests.
* ``shard``       the share of a particle array a rank uploads (any split works: step 1 migrates to the owners).
* ``bench_main``  the N > 1 arm of bench.py.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import capi
from .capi import PARTICLE, DistStats, Params, check, lib
from .solver import Solver


def shard(xs: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Contiguous block `rank` of `world` of the input array (sizes differ by at most one)."""
    n = len(xs)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    return np.ascontiguousarray(xs[lo:hi])


def stats_dict(s: DistStats) -> dict:
    return {k: int(getattr(s, k)) for k, _ in DistStats._fields_}


class _RankOps:
    """Calls shared by both transports; `self.s` is the rank's Solver (context)."""

    s: Solver

    def upload(self, xs: np.ndarray) -> None:
        assert xs.dtype == PARTICLE and xs.flags.c_contiguous
        self.s._ck(self.s._L.pbf_dist_upload(self.s._ctx, xs.ctypes.data, len(xs)))

    def download(self) -> np.ndarray:
        xs = np.zeros(self.s.count(), PARTICLE)
        n = C.c_uint64(0)
        self.s._ck(self.s._L.pbf_dist_download(self.s._ctx, xs.ctypes.data, len(xs), C.byref(n)))
        return xs[: n.value]

    def stats(self) -> dict:
        st = DistStats()
        self.s._ck(self.s._L.pbf_dist_stats_read(self.s._ctx, C.byref(st)))
        return stats_dict(st)

    def set_replan(self, steps: int) -> None:
        self.s._ck(self.s._L.pbf_dist_set_replan(self.s._ctx, steps))


class SlabRank(_RankOps):
    """One rank of the NCCL transport.  `id_bytes` = the 128-byte id from ``unique_id()`` on rank 0."""

    def __init__(self, h: float, device: int, rank: int, world: int, id_bytes: bytes, flags: int = 0):
        self.s = Solver(h, device, flags)
        self.rank, self.world = rank, world
        buf = (C.c_uint8 * capi.NCCL_ID_BYTES).from_buffer_copy(id_bytes)
        self.s._ck(self.s._L.pbf_dist_init(self.s._ctx, buf, rank, world))

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * capi.NCCL_ID_BYTES)()
        check(lib().pbf_dist_unique_id(buf))
        return bytes(buf)

    def step(self, params: Params) -> None:
        self.s._ck(self.s._L.pbf_dist_step(self.s._ctx, C.byref(params)))

    def close(self) -> None:
        self.s.close()


class _LocalRank(_RankOps):
    def __init__(self, s: Solver):
        self.s = s


class LocalGroup:
    """`world` slab ranks inside this process (devices[i] = CUDA ordinal of rank i; they may coincide)."""

    def __init__(self, h: float, devices: list[int], flags: int = 0):
        self.solvers = [Solver(h, d, flags) for d in devices]
        self.ranks = [_LocalRank(s) for s in self.solvers]
        arr = (C.c_void_p * len(devices))(*[s._ctx for s in self.solvers])
        check(lib().pbf_dist_init_local(arr, len(devices)), self.solvers[0]._ctx)

    @property
    def world(self) -> int:
        return len(self.ranks)

    def upload(self, xs: np.ndarray) -> None:
        for r, rank in enumerate(self.ranks):
            rank.upload(shard(xs, r, self.world))

    def step(self, params: Params) -> None:
        s = self.solvers[0]
        s._ck(s._L.pbf_dist_step(s._ctx, C.byref(params)))

    def advance(self, params: Params, xs: np.ndarray) -> int:
        """sph::Solver::advance over the group (returns the mesh vertex count): `xs` in place, back in the global Z order (pbf_dist_advance_host)."""
        assert xs.dtype == PARTICLE and xs.flags.c_contiguous
        s = self.solvers[0]
        nv = C.c_uint64(0)
        s._ck(s._L.pbf_dist_advance_host(s._ctx, C.byref(params), xs.ctypes.data, len(xs), C.byref(nv)))
        return int(nv.value)

    def mesh(self):
        """The surface of the last step (params.surface_enabled): every rank fills the lattice points it owns, rank 0
        extracts the triangles — the single-device mesh."""
        return self.solvers[0].mesh()

    def sync(self) -> None:
        for s in self.solvers:
            s.sync()

    def download(self) -> np.ndarray:
        """All particles, rank 0 first: the global Z order."""
        return np.concatenate([r.download() for r in self.ranks])

    def close(self) -> None:
        for s in self.solvers:
            s.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ------------------------------------------------------------------------------------------------- bench.py, N > 1
def _new_rank(torch, dist, device, rank, world, flags):
    """A SlabRank on a fresh NCCL communicator (rank 0 makes the id, torch.distributed carries it)."""
    from . import scenes
    idt = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(SlabRank.unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    return SlabRank(scenes.H, device, rank, world, bytes(idt.cpu().numpy().tobytes()), flags)


def nccl_parity(torch, dist, device, rank, world, frames=6) -> dict:
    """`frames` moving-wall frames of the stock two-cube scene (18 522 particles) on the N-rank NCCL group against the
    same frames on one device (rank 0): particle order, positions, velocities and colours must agree bit for bit."""
    from . import scenes
    p, xs = scenes.two_cubes(20000, 4)
    sr = _new_rank(torch, dist, device, rank, world, 0)
    sr.set_replan(2)
    sr.upload(shard(xs, rank, world))
    for f in range(frames):
        sr.step(scenes.apply_motion(p, f))
    mine, st = sr.download(), sr.stats()
    parts = [None] * world
    dist.all_gather_object(parts, (mine.tobytes(), st["owned"], st["ghosts"]))
    sr.close()
    verdict = None
    if rank == 0:
        got = np.concatenate([np.frombuffer(b, dtype=PARTICLE) for b, _, _ in parts])
        with Solver(scenes.H, device) as s:
            s.upload(xs)
            for f in range(frames):
                s.step(scenes.apply_motion(p, f))
            ref = s.download()
        same = (len(ref) == len(got) and np.array_equal(ref["id"], got["id"])
                and all(np.array_equal(ref[k].view(np.uint32), got[k].view(np.uint32)) for k in ("position", "velocity", "colour")))
        verdict = {"nccl_vs_single": "bit-identical" if same else "MISMATCH", "ranks": world, "frames": frames,
                   "scene": "stock two-cube scene, 18 522 particles, moving wall, 4 solver iterations, splits re-planned every 2 steps",
                   "owned": [o for _, o, _ in parts], "ghosts": [g for _, _, g in parts]}
    out = [verdict]
    dist.broadcast_object_list(out, 0)
    return out[0]


def _timed_group_steps(torch, dist, sr, stream, p, steps):
    """`steps` slab steps, barrier on both sides, CUDA events on the rank's stream, max over ranks -> ms total."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record(stream)
    for _ in range(steps):
        sr.step(p)
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def _profiled_group_steps(torch, dist, sr, p, steps, flags):
    sr.s.set_flags(capi.FLAG_PROFILE | flags)
    sr.s.profile_reset()
    torch.cuda.synchronize()
    dist.barrier()
    for _ in range(steps):
        sr.step(p)
    torch.cuda.synchronize()
    dist.barrier()
    prof = sr.s.profile()
    sr.s.set_flags(flags)
    return prof


def secondary_multi(torch, dist, args, device, rank, world, UNIT) -> list:
    """BASELINE.json configs[2] (dam-8m split over the N GPUs: strong scaling) and configs[4] (dam-weak-8m: ~8 M particles
    per GPU, 8 solver iterations: weak scaling), few steps each.  Rank 0 first measures the single-GPU figure each entry's
    efficiency is quoted against (dam(200) with 4 and with 8 iterations) while the other ranks wait.  A failure inside one
    entry is reported in that entry (`error`) on every rank alike — it never takes the primary line down with it."""
    from . import scenes
    k = max(3, args.secondary_steps)
    single = {}
    if rank == 0:
        try:
            p1, xs1 = scenes.dam_break(200, 4)
            stream = torch.cuda.Stream()
            with Solver(scenes.H, device, args.flags) as s:
                s.set_stream(stream.cuda_stream)
                s.upload(xs1)
                del xs1
                for _ in range(args.settle):
                    s.step(p1)
                for iters in (4, 8):
                    p1.iteration = iters
                    for _ in range(3):
                        s.step(p1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record(stream)
                    for _ in range(k):
                        s.step(p1)
                    e1.record(stream)
                    torch.cuda.synchronize()
                    single[iters] = e0.elapsed_time(e1) / k
        except Exception as e:  # noqa: BLE001
            single = {"error": repr(e)}
    box = [single]
    dist.broadcast_object_list(box, 0)
    single = box[0]
    out = []
    for name, side, iters, scaling in (("dam-8m", 200, 4, "strong"),
                                       ("dam-weak-8m", int(round((8_000_000 * world) ** (1 / 3))), 8, "weak")):
        sr, mine_result, err, n_total = None, None, None, side ** 3
        try:
            p, mine, n_total = scenes.dam_break_shard(side, iters, rank, world)
            sr = _new_rank(torch, dist, device, rank, world, args.flags)
            stream = torch.cuda.Stream()
            sr.s.set_stream(stream.cuda_stream)
            sr.upload(mine)
            del mine
            for _ in range(args.settle + 3):
                sr.step(p)
            # no torch collective inside the guarded region (a failed rank would leave its peers inside it): every rank times
            # its own stream — the slab barriers inside the steps keep the ranks together — and the maximum is taken below
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(k):
                sr.step(p)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            sr.s.set_flags(capi.FLAG_PROFILE | args.flags)
            sr.s.profile_reset()
            for _ in range(k):
                sr.step(p)
            torch.cuda.synchronize()
            prof = sr.s.profile()
            sr.s.set_flags(args.flags)
            st = sr.stats()
            mine_result = {"rank": rank, "owned": st["owned"], "ghosts": st["ghosts"], "ms_total": ms,
                           "ms": {f: round(v / k, 4) for f, v in prof["ms"].items() if v > 0}}
        except Exception as e:  # noqa: BLE001  (a rank that fails makes its peers' barriers time out: they land here too)
            err = repr(e)
        try:
            if sr is not None:
                sr.close()
        except Exception:  # noqa: BLE001
            pass
        ranks = [None] * world
        dist.all_gather_object(ranks, mine_result if err is None else {"rank": rank, "error": err})
        workload_desc = f"dam-break {side}^3 = {n_total} particles, {iters} solver iterations, {world} GPUs"
        if any("error" in r for r in ranks) or "error" in single:
            out.append({"name": name, "workload": workload_desc, "n_gpus": world, "scaling": scaling,
                        "error": [r["error"] for r in ranks if "error" in r] or single.get("error")})
            continue
        ms_step = max(r.pop("ms_total") for r in ranks) / k
        one = single[iters]  # dam(200) = 8 M particles on one GPU with the same iteration count
        if scaling == "strong":
            eff = one / (world * ms_step)
        else:  # weak: per-GPU work ~ constant; compare particle-iterations/s per GPU with the single-GPU run
            eff = (n_total * iters / ms_step / world) / (8_000_000 * iters / one)
        out.append({"name": name, "workload": workload_desc,
                    "n_gpus": world, "scaling": scaling, "particles": n_total, "solver_iterations": iters, "steps": k,
                    "settle_steps": args.settle, "ms_per_step": ms_step, "value": n_total * iters / (ms_step * 1e-3), "unit": UNIT,
                    "single_gpu": {"workload": f"dam-break 200^3 = 8000000 particles, {iters} solver iterations", "ms_per_step": one,
                                   "value": 8_000_000 * iters / (one * 1e-3)},
                    "efficiency_vs_single_gpu": eff, "ranks": ranks})
    return out


def bench_main(args, workload, ClockSampler, METRIC, UNIT, roofline_of=None) -> None:
    """One rank per GPU under torchrun: weak scaling (the dam-break block grows so that every GPU holds ~1 M
    particles), device time by CUDA events, max over ranks, rank 0 prints the JSON line."""
    import sys
    import torch
    import torch.distributed as dist
    from . import scenes

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    parity = nccl_parity(torch, dist, local, rank, world)
    if parity["nccl_vs_single"] != "bit-identical":
        if rank == 0:
            print(json.dumps({"metric": METRIC, "n_gpus": world, "parity": parity, "error": "NCCL slab path differs from one device"}))
        dist.destroy_process_group()
        sys.exit(3)
    name, desc, p, xs = workload(args.workload, world)
    n_total, iters = len(xs), int(p.iteration)

    sr = _new_rank(torch, dist, local, rank, world, args.flags)
    stream = torch.cuda.Stream()
    sr.s.set_stream(stream.cuda_stream)
    mine = shard(xs, rank, world)
    del xs
    sr.upload(mine)
    for _ in range(args.settle + args.warmup):
        sr.step(p)
    sr.s.sync()
    sr.s.profile_reset()
    l0 = sr.s.launch_count()
    # NVML queries go through the driver: one sampler (rank 0, 10 Hz) is enough and keeps the other ranks' launches quiet
    with ClockSampler(local, period=0.1, enabled=(rank == 0)) as clk:
        ms_total = _timed_group_steps(torch, dist, sr, stream, p, args.steps)
    launches = sr.s.launch_count() - l0
    # the same steps again with the library's per-family CUDA events on (they stretch the step, so `value` is
    # timed without them): per-rank phase times and the roofline's launch durations
    prof = _profiled_group_steps(torch, dist, sr, p, args.steps, args.flags)
    st = sr.stats()

    # end to end: every step uploads the rank's particles from pinned host memory and reads them back into it
    import time
    L = lib()
    n_cur = sr.s.count()
    cap = int(n_cur * 1.5) + 4096
    ptr = L.pbf_host_alloc(cap * PARTICLE.itemsize)
    n_out = C.c_uint64(0)
    sr.s._ck(L.pbf_dist_download(sr.s._ctx, C.c_void_p(ptr), cap, C.byref(n_out)))
    e2e_steps = max(3, min(args.steps, 10))
    moved = 0

    def roundtrip(n: int) -> int:
        sr.s._ck(L.pbf_dist_upload(sr.s._ctx, C.c_void_p(ptr), n))
        sr.step(p)
        sr.s._ck(L.pbf_dist_download(sr.s._ctx, C.c_void_p(ptr), cap, C.byref(n_out)))
        return int(n_out.value)

    n_cur = int(n_out.value)
    for _ in range(2):
        n_cur = roundtrip(n_cur)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        moved += n_cur
        n_cur = roundtrip(n_cur)
    torch.cuda.synchronize(); dist.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    moved_t = torch.tensor([moved], device="cuda", dtype=torch.float64)
    dist.all_reduce(moved_t, op=dist.ReduceOp.SUM)
    L.pbf_host_free(ptr)

    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, **st, "launches": launches,
                                      "ms": {k: round(v / args.steps, 4) for k, v in prof["ms"].items() if v > 0}})
    clocks = clk.summary()
    sr.close()
    secondary = None
    if not args.no_secondary and args.workload == "auto":
        secondary = secondary_multi(torch, dist, args, local, rank, world, UNIT)
    if rank == 0:
        value = n_total * iters * args.steps / (ms_total * 1e-3)
        e2e_val = n_total * iters * e2e_steps / float(e2e_s.item())
        per_step_bytes = float(moved_t.item()) / e2e_steps * PARTICLE.itemsize
        sums = [sum(v for k, v in g["ms"].items() if k not in ("diffuse", "slab_setup", "slab_barrier", "slab_iterations")) for g in gathered]
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak" if name.startswith("dam-weak") else "strong",  # a fixed-size --workload split over the ranks is strong scaling
            "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "name": name, "particles": n_total, "solver_iterations": iters,
                       "settle_steps": args.settle, "decomposition": "Z-curve slabs; migrants, ghosts and the per-iteration halo (16 B per ghost) move by kernels storing into peer memory over NVLink (arenas mapped through CUDA IPC), flag barriers on the stream; NCCL for the bootstrap and the plan steps' histogram all-reduce",
                       "l2": "no flush: the per-step working set exceeds the 126 MB L2"},
            "particle_steps_per_sec": n_total * args.steps / (ms_total * 1e-3),
            # rank 0's dominant kernel family over the particles rank 0 processes (owned + ghosts)
            "roofline": roofline_of(prof, st["owned"] + st["ghosts"], iters, args.steps) if roofline_of else None,
            "cpu_baseline": None,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(per_step_bytes),
                    "d2h_bytes_per_step": int(per_step_bytes), "steps": e2e_steps,
                    "api": "pbf_dist_upload (host AoS) -> pbf_dist_step -> pbf_dist_download, every step, all ranks"},
            "gpu_launches": int(sum(g["launches"] for g in gathered)), "parity": parity,
            "rank_ms_spread": {"min": min(sums), "max": max(sums), "note": "per-rank sum of the kernel families (events pass; without the side-stream diffusion, the barrier waits and the two spans slab_setup / slab_iterations, which overlap the families), ms/step"},
            "ranks": gathered, "clocks": clocks, "secondary": secondary,
        }))
    dist.destroy_process_group()
