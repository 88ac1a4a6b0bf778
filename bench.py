#!/usr/bin/env python
"""bench.py — particle-iterations/s of the PBF step on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload dam-1m]

* N = 1: BASELINE.json configs[1] — dam-break, 100^3 = 1 000 000 particles, 4 solver iterations, no surface
  extraction.  A "step" is one whole PBF step (predict/key, sort, reorder, cell table, diffuse, 4 x (lambda, delta),
  finalise) over all particles.  `value` = particles x iterations x K / device time, state resident in HBM.
* N > 1 (torchrun, one rank per GPU): Z-curve slab decomposition with NCCL ghost/migrant exchange, weak scaling
  (dam-break sized so every GPU holds ~1 M particles... see --workload), max-over-ranks device time.
* `e2e` = the same metric through the drop-in call (pbf_advance_host == sph::Solver::advance): pinned HOST buffers,
  H2D + step + D2H inside the timed region.
* `roofline` = the dominant kernel family (CUDA events recorded by the library on its own stream during the timed
  region) against MEASURED_PEAKS.json; `cpu_baseline` = the reference OpenMP backend (oracle/_ref, built from the
  unmodified sources) on this box's host cores on a bounded sample.
* --impl reference times that same reference CPU implementation as the driver's comparison arm.
* `secondary` = BASELINE.json's other named configurations measured in the same run with few steps:
  N = 1: dam-1m-mc (configs[3]), dam-8m on one GPU (configs[2]'s baseline), dam-8m with 8 iterations (configs[4]'s
  first point); N > 1: dam-8m split N ways (configs[2], strong scaling) and dam-weak-8m (configs[4], ~8 M per GPU,
  8 iterations), each with its own single-GPU figure measured by rank 0 in the same job.
* N > 1: `parity` = 6 moving-wall frames of the stock scene on the N-rank NCCL group against one device, bit for bit.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "particle_iterations_per_sec"
UNIT = "particle-iterations/s"

# algorithmic bytes per particle per launch of each kernel family
ALG_BYTES = {
    "predict_key": 36,   # R pos4 16 + vel4 16, W key 4
    "sort": 3 * 20,      # 3 passes x (hist R 4 + scatter R 8 + W 8)   [first pass reads no value: -4]
    "reorder": 148,      # R perm 4 + pos4/vel4/col4 48 + id 8, W pos4/vel4/col4/pstar4 64 + id 8 ... see DESIGN.md
    "cell_table": 6,
    "diffuse": 36,       # R key 4 + col4 16, W col4 16
    "lambda": 40,        # R pstar4 16 + key 4 + mass 4, W pstar4|lambda 16
    "delta": 36,         # R pstar4|lambda 16 + key 4, W pstar4 16
    "finalise": 80,      # R pstar4 16 + pos4 16 + vel4 16, W pos4 16 + vel4 16
}


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md 'clocks' line): an NVML polling thread
    (20 ms period, started before the region, stopped after it); `nvidia-smi -lms` as the fallback when NVML cannot
    be loaded."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, period: float = 0.02, enabled: bool = True):
        self.index, self.rows, self.proc, self.stop, self.how = index, [], None, threading.Event(), None
        self.period, self.enabled = period, enabled

    def _nvml_loop(self, nv, h, mx):
        bits = [getattr(nv, n, 0) for n in ("nvmlClocksEventReasonHwSlowdown", "nvmlClocksEventReasonHwThermalSlowdown",
                                            "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksEventReasonSwPowerCap")]
        alt = [getattr(nv, n, 0) for n in ("nvmlClocksThrottleReasonHwSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown",
                                           "nvmlClocksThrottleReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwPowerCap")]
        bits = [b or a for b, a in zip(bits, alt)]
        while True:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if (b and r & b) else "Not Active" for b in bits])
            except Exception:
                pass
            if self.stop.wait(self.period):
                break

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            import pynvml as nv
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: go through the PCI bus id of the CUDA device
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(nv.nvmlDeviceGetCount()):
                    hi = nv.nvmlDeviceGetHandleByIndex(i)
                    if int(nv.nvmlDeviceGetPciInfo(hi).bus) == int(bus):
                        h = hi
                        break
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.how = "nvml"
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h, mx), daemon=True)
            self.t.start()
            return self
        except Exception:
            pass
        try:
            self.how = "nvidia-smi"
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif getattr(self, "t", None):
            self.t.join(timeout=1)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in list(self.rows):
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(self.NAMES, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.how}


def workload(name: str, world: int):
    from pbf_sph_b200 import scenes
    if name == "auto":
        name = "dam-1m" if world == 1 else "dam-weak"
    if name == "dam-1m":
        p, xs = scenes.dam_break(100, 4)
        return name, "dam-break 100^3 = 1 000 000 particles, 4 solver iterations, no surface extraction", p, xs
    if name == "dam-1m-mc":  # BASELINE.json configs[3]: the mesh-generation path
        p, xs = scenes.dam_break(100, 4)
        p.surface_enabled = 1
        return name, "dam-break 100^3 = 1 000 000 particles, 4 solver iterations, marching-cubes surface every step", p, xs
    if name == "dam-64k":
        p, xs = scenes.dam_break(40, 4)
        return name, "dam-break 40^3 = 64 000 particles, 4 solver iterations", p, xs
    if name == "dam-8m":
        p, xs = scenes.dam_break(200, 4)
        return name, "dam-break 200^3 = 8 000 000 particles, 4 solver iterations", p, xs
    if name == "dam-weak":  # ~1 M particles per GPU
        side = int(round((1_000_000 * world) ** (1 / 3)))
        p, xs = scenes.dam_break(side, 4)
        return name, f"dam-break {side}^3 = {side ** 3} particles ({world} GPUs, ~1 M per GPU), 4 solver iterations", p, xs
    if name == "dam-weak-8m":  # BASELINE.json configs[4]: weak scaling 8 M -> 64 M particles, 8 solver iterations
        side = int(round((8_000_000 * world) ** (1 / 3)))
        p, xs = scenes.dam_break(side, 8)
        return name, f"dam-break {side}^3 = {side ** 3} particles ({world} GPUs, ~8 M per GPU), 8 solver iterations", p, xs
    raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------------------------------------------ reference arm
def reference_cpu(params, xs, steps, warmup, settle_steps=100, budget_s=150.0):
    """The reference's own CPU implementation of the path (oracle/_ref: unmodified ompsph.hpp, all host threads),
    on a bounded sample of the workload.  Returns (PI/s, ms/step, cores, kind, sample description)."""
    import oracle
    from pbf_sph_b200 import scenes
    cores = os.cpu_count() or 1
    use_ref = oracle.ref_available("fast")
    if use_ref:
        L = oracle.ref_lib("best")
        L.pbf_ref_set_threads(cores)
        kind, variant = "reference", L.variant_name

        def advance(p, a):
            oracle.ref_advance(scenes.H, p, a, variant=variant)
    else:
        kind, variant = "port", "oracle Jacobi restatement"

        def advance(p, a):
            oracle.step(scenes.H, p, a)
    # Size the sample: time one step of a 64 K block and extrapolate linearly in N.  Like our arm, the fluid is settled
    # first (throughput depends on the state: ~266 candidates per particle on the initial lattice, ~130 once settled),
    # here by the reference itself, so settle + warm-up + timed steps must all fit the budget.
    probe_p, probe = scenes.dam_break(40, int(params.iteration))
    t0 = time.perf_counter(); advance(probe_p, probe); t_probe = time.perf_counter() - t0
    t0 = time.perf_counter(); advance(probe_p, probe); t_probe = min(t_probe, time.perf_counter() - t0)
    per_particle = t_probe / len(probe)
    settle = settle_steps
    n_budget = budget_s / max(1, steps + warmup + settle) / per_particle
    same = n_budget >= len(xs)
    if same:
        p, a, sample = params, xs.copy(), f"the full workload ({len(xs)} particles)"
    else:
        side = max(24, int(n_budget ** (1 / 3)))
        p, a = scenes.dam_break(side, int(params.iteration))
        sample = (f"dam-break {side}^3 = {side ** 3} particles (same scene family, spacing and iterations as the workload; "
                  f"sized so that settle + warm-up + timed steps fit {budget_s:.0f} s of CPU time)")
    for _ in range(settle + warmup):
        advance(p, a)
    t0 = time.perf_counter()
    for _ in range(steps):
        advance(p, a)
    dt = time.perf_counter() - t0
    pis = len(a) * int(p.iteration) * steps / dt
    note = ("; each call constructs the reference's solver and copies the particle vector in and out (a few % at 1 M); the "
            "reference settles its own fluid with its in-place (Gauss-Seidel, racy) delta pass, the GPU arm with the Jacobi form, "
            "so the two arms time statistically equal but not identical fluid states")
    return pis, dt / steps * 1e3, cores, kind, (f"{sample}, settled {settle} steps by the reference itself, then {warmup} warm-up "
                                                f"+ {steps} timed steps; {variant}; {cores} threads" + note), same, len(a)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, desc, p, xs = workload(args.workload, max(1, args.gpus))
    pis, ms, cores, kind, sample, same, n_timed = reference_cpu(p, xs, args.steps, args.warmup, args.settle, budget_s=args.ref_budget)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": pis, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "name": name},
        # false = the CPU arm timed a smaller block of the same scene family than the GPU arm's workload (bounded sample)
        "same_config": bool(same), "particles_timed": int(n_timed), "particles_workload": int(len(xs)),
        "cpu_baseline": {"value": pis, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": pis, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def roofline_of(prof, n, iters, steps):
    """Roofline object of the dominant kernel family: algorithmic bytes per launch (ALG_BYTES x particles the launch
    processes) / the family's average launch duration from the library's CUDA events on its own stream, against the
    measured HBM copy peak.  `traffic` = DRAM bytes per launch of that kernel from the committed `ncu --set full` capture
    of the same workload (profiles/ncu_traffic.json), or null when the family has no capture."""
    peak, peak_src = peaks()
    fam = max(("lambda", "delta"), key=lambda k: prof["ms"][k])
    n_launch = max(1, prof["launches"][fam])
    avg_ms = prof["ms"][fam] / n_launch
    achieved = n * ALG_BYTES[fam] / (avg_ms * 1e-3) / 1e9
    breakdown = {k: round(v / steps, 4) for k, v in prof["ms"].items() if v > 0}
    per_kernel = {}
    for k, b in ALG_BYTES.items():
        if prof["launches"].get(k) and prof["ms"][k] > 0:
            per_kernel[k] = round(n * b / (prof["ms"][k] / prof["launches"][k] * 1e-3) / 1e9, 1) if k in ("lambda", "delta") \
                else round(n * b / (prof["ms"][k] / steps * 1e-3) / 1e9, 1)
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "ncu_traffic.json"
    if tf.exists():
        t = json.loads(tf.read_text())
        if fam in t.get("dram_bytes_per_launch", {}) and t.get("particles") == n:
            traffic, traffic_src = t["dram_bytes_per_launch"][fam], t.get("source")
    return {"bound": "hbm", "kernel": fam, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "bytes_per_particle": ALG_BYTES[fam], "avg_launch_ms": avg_ms,
            "note": "the neighbour passes are bound by the latency of their gathers, the L1 data pipe and instruction issue "
                    "(~195 candidate pairs per particle, one 16-byte position through L1 per pair and lane), not by HBM; "
                    "`issue` is the fraction of the warp-instruction issue slots; see DESIGN.md §4 and "
                    "profiles/r01c_search_experiments.txt, r02b_block_size.txt, r02c_list_layout.txt", "ms_per_step_by_family": breakdown,
            "achieved_GBps_by_family": per_kernel}


# ------------------------------------------------------------------------------------------------ our arm, 1 GPU
def time_resident(torch, s, stream, p, steps):
    """K resident steps between two CUDA events on the solver's stream -> (ms total, launches)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = s.launch_count()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        s.step(p)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), s.launch_count() - l0


def profiled_pass(torch, s, stream, p, steps, flags, families=None):
    """The same K steps with the library's CUDA events around the given kernel families (None = all)."""
    from pbf_sph_b200 import FLAG_PROFILE
    s.set_flags(FLAG_PROFILE | flags)
    s.profile_mask(families)
    s.profile_reset()
    ms, _ = time_resident(torch, s, stream, p, steps)
    prof = s.profile()
    s.set_flags(flags)
    s.profile_mask(None)
    return ms, prof


def secondary_single(torch, s, stream, args, which):
    """BASELINE.json's other configurations on ONE GPU, few steps each: settle, warm up, time K resident steps, then one
    profiled pass for the per-family milliseconds."""
    out = []
    k = max(3, args.secondary_steps)
    for name, iters_list in which:
        wname, desc, p, xs = workload(name, 1)
        s.upload(xs)
        n = len(xs)
        del xs
        for _ in range(args.settle):
            s.step(p)
        for iters in iters_list:
            p.iteration = iters
            for _ in range(3):
                s.step(p)
            ms, launches = time_resident(torch, s, stream, p, k)
            _, prof = profiled_pass(torch, s, stream, p, k, args.flags)
            entry = {"name": wname if iters == iters_list[0] else f"{wname}-{iters}it", "workload": desc if iters == iters_list[0]
                     else desc.replace(f"{iters_list[0]} solver iterations", f"{iters} solver iterations"),
                     "n_gpus": 1, "particles": n, "solver_iterations": iters, "steps": k, "settle_steps": args.settle,
                     "ms_per_step": ms / k, "value": n * iters * k / (ms * 1e-3), "unit": UNIT,
                     "us_per_particle_step": ms / k * 1e3 / n, "gpu_launches": launches,
                     "ms_per_step_by_family": {f: round(v / k, 4) for f, v in prof["ms"].items() if v > 0}}
            if p.surface_enabled:
                s.sync()
                entry["triangles"] = int(s.grid().n_triangles)
            out.append(entry)
    return out


def run_single(args):
    import torch
    from pbf_sph_b200 import FLAG_DEBUG_COUNTS, PARTICLE, Solver, capi, scenes
    name, desc, p, xs = workload(args.workload, 1)
    n, iters = len(xs), int(p.iteration)
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream()
    s = Solver(scenes.H, dev, args.flags)
    s.set_stream(stream.cuda_stream)
    s.upload(xs)
    # settle the fluid first (throughput depends on the state: ~266 candidates/particle on the initial lattice,
    # ~130-190 once settled) — these steps are outside both the warm-up and the timed region
    for _ in range(args.settle):
        s.step(p)
    for _ in range(args.warmup):
        s.step(p)
    s.sync()
    with ClockSampler(dev, period=0.05) as clk:
        ms_total, launches = time_resident(torch, s, stream, p, args.steps)
    value = n * iters * args.steps / (ms_total * 1e-3)
    # Kernel durations: (a) every family timed (~30 event records per step — they stretch a 1.8 ms step by up to a fifth,
    # which is why `value` is timed without them): the per-family breakdown; (b) the two solver families alone (16 records
    # per step): the launch duration the roofline uses.
    ms_all, prof = profiled_pass(torch, s, stream, p, args.steps, args.flags)
    ms_dom, prof_dom = profiled_pass(torch, s, stream, p, args.steps, args.flags, ["lambda", "delta"])
    for f in ("lambda", "delta"):
        prof["ms"][f], prof["launches"][f] = prof_dom["ms"][f], prof_dom["launches"][f]
    roofline = roofline_of(prof, n, iters, args.steps)
    roofline["region"] = (f"{args.steps} further steps with CUDA events recorded by the library on its stream around the lambda / delta "
                          f"launches only ({ms_dom / args.steps:.4f} ms/step; {ms_total / args.steps:.4f} without any events, "
                          f"{ms_all / args.steps:.4f} with every family timed — the other families' figures come from that pass)")
    # what the neighbour passes are really limited by: pair tests / evaluations per second and issue slots
    s.set_flags(FLAG_DEBUG_COUNTS | args.flags)
    s.step(p)
    s.sync()
    cand, hits = s.tap(capi.TAP_CAND_COUNT).astype(np.int64), s.tap(capi.TAP_NBR_COUNT).astype(np.int64)
    s.set_flags(args.flags)
    lam_ms = prof["ms"]["lambda"] / max(1, prof["launches"]["lambda"])
    del_ms = prof["ms"]["delta"] / max(1, prof["launches"]["delta"])
    roofline["limiter"] = "latency / issue (not HBM): see `pairs` and `issue`"
    roofline["pairs"] = {
        "candidates_per_particle": float(cand.mean()), "neighbours_per_particle": float(hits.mean()),
        "pair_tests_per_s": float(cand.sum() / (lam_ms * 1e-3)),
        "pair_evaluations_per_s": float(2 * hits.sum() / ((lam_ms + del_ms) * 1e-3)),
        "note": "tests = 27-cell candidates the lambda pass's search examines per launch / its launch duration (the search "
                "shares the launch with the sums over the hits); evaluations = kernel-function evaluations of the lambda and "
                "delta passes (one per in-radius pair each) / their combined duration"}
    tf = ROOT / "profiles" / "ncu_traffic.json"
    if tf.exists():
        t = json.loads(tf.read_text())
        inst = t.get("warp_inst_per_launch", {}).get(roofline["kernel"])
        clocks = clk.summary()
        if inst and t.get("particles") == n and clocks.get("sm_mhz"):
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak_issue = sms * 4 * clocks["sm_mhz"] * 1e6 / 1e9
            got = inst / (roofline["avg_launch_ms"] * 1e-3) / 1e9
            roofline["issue"] = {"achieved": got, "peak": peak_issue, "unit": "G warp-inst/s", "frac": got / peak_issue,
                                 "warp_inst_per_launch": inst,
                                 "source": "smsp__inst_executed.sum of one launch (" + t.get("source", "ncu") + "); peak = "
                                           f"{sms} SMs x 4 schedulers x {clocks['sm_mhz']:.0f} MHz"}
        if t.get("particles") == n and t.get("fma_pipe_pct", {}).get(roofline["kernel"]) is not None:
            roofline["fp32_pipe_utilisation_pct"] = {"value": t["fma_pipe_pct"][roofline["kernel"]],
                                                      "source": "sm__inst_executed_pipe_fma (ncu capture, " + t.get("source", "") + ")"}

    # end to end through the drop-in call, pinned host buffers
    snap = s.download()
    L = capi.lib()
    nbytes = n * PARTICLE.itemsize
    ptr = L.pbf_host_alloc(nbytes)
    host = np.frombuffer((__import__("ctypes").c_char * nbytes).from_address(ptr), dtype=PARTICLE)
    host[:] = snap
    e2e_steps = max(3, min(args.steps, 20)) if not args.no_e2e else 1
    for _ in range(3 if not args.no_e2e else 0):
        s.advance_ptr(p, ptr, n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.advance_ptr(p, ptr, n)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": n * iters * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nbytes,
           "d2h_bytes_per_step": nbytes, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
           "api": "pbf_advance_host (== sph::Solver::advance): pinned host AoS -> H2D -> step -> D2H"}
    del host
    L.pbf_host_free(ptr)

    # CPU baseline: the reference OpenMP backend on this box's host cores, same snapshot, bounded sample
    cpu = None
    if not args.no_cpu_baseline:
        import oracle
        cores = os.cpu_count() or 1
        if oracle.ref_available("fast"):
            Lr = oracle.ref_lib("best")
            Lr.pbf_ref_set_threads(cores)
            a = snap.copy()
            k, t_spent = 0, 0.0
            oracle.ref_advance(scenes.H, p, a, variant=Lr.variant_name)  # untimed first call (page faults)
            while k < 8 and t_spent < 15.0:
                t0 = time.perf_counter()
                oracle.ref_advance(scenes.H, p, a, variant=Lr.variant_name)
                t_spent += time.perf_counter() - t0
                k += 1
            cpu = {"value": n * iters * k / t_spent, "unit": UNIT, "cores": cores, "kind": "reference",
                   "ms_per_step": t_spent / k * 1e3,
                   "sample": f"{k} steps of the full workload from the same settled snapshot; unmodified ompsph.hpp, "
                             f"{Lr.variant_name} build, {cores} OpenMP threads"}
        else:
            a = snap.copy()
            t0 = time.perf_counter(); oracle.step(scenes.H, p, a); t1 = time.perf_counter() - t0
            cpu = {"value": n * iters / t1, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "1 step of the full workload; oracle Jacobi restatement (oracle/_ref absent)"}
    del snap
    secondary = None
    if not args.no_secondary and args.workload in ("auto", "dam-1m"):
        # BASELINE.json configs[3] (surface every step), configs[2] on one GPU, and configs[4]'s first point (8 M, 8 iterations)
        try:  # a failure here is reported in `secondary`, it never takes the primary line down with it
            secondary = secondary_single(torch, s, stream, args, [("dam-1m-mc", [4]), ("dam-8m", [4, 8])])
            base = ms_total / args.steps * 1e3 / n
            for e in secondary:
                e["us_per_particle_step_vs_dam_1m"] = e["us_per_particle_step"] / base if e["solver_iterations"] == iters else None
        except Exception as e:  # noqa: BLE001
            secondary = [{"error": repr(e)}]
    s.close()
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "name": name, "particles": n, "solver_iterations": iters,
                   "settle_steps": args.settle,
                   "l2": "no flush: the per-step working set (8 float4 + 2 u64 + 5 u32 arrays ~ 170 MB at 1 M) exceeds "
                         "the 126 MB L2"},
        "particle_steps_per_sec": n * args.steps / (ms_total * 1e-3),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
        "clocks": clk.summary(), "secondary": secondary,
    }
    print(json.dumps(out))


def run_multi(args):
    from pbf_sph_b200 import dist
    dist.bench_main(args, workload, ClockSampler, METRIC, UNIT, roofline_of)


def _protect_stdout():
    """Native libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line on stdout.
    Point fd 1 at stderr for the run and keep the real stdout for the result line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--settle", type=int, default=100, help="untimed steps that settle the fluid before warm-up")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip BASELINE's other configurations (`secondary`)")
    ap.add_argument("--secondary-steps", type=int, default=10)
    ap.add_argument("--ref-budget", type=float, default=240.0, help="--impl reference: seconds of CPU time for settle + warm-up + timed steps")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--flags", type=int, default=0, help="extra PBF_FLAG_* bits (1 strict fp, 8 global-memory neighbours)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.gpus > 1 and "RANK" not in os.environ:
        # launched by hand without torchrun: re-launch one rank per GPU the way the driver does
        os.dup2(sys.stdout.fileno(), 1)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"),
                                   str(Path(__file__).resolve())] + sys.argv[1:])
    if args.impl == "reference":
        run_reference(args)
    elif args.gpus <= 1 and int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        run_single(args)
    else:
        run_multi(args)


if __name__ == "__main__":
    main()
