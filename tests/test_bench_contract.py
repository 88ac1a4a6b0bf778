"""bench.py's host-side logic on CPU: the roofline object, the clock sampler when disabled / without NVML, the
workload table and the reference arm's JSON line (on a tiny sample) — the keys the driver's contract names."""
import importlib.util
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_roofline_object(bench):
    fams = ["predict_key", "sort", "reorder", "cell_table", "diffuse", "lambda", "delta", "finalise"]
    prof = {"ms": {k: 0.0 for k in fams}, "launches": {k: 0 for k in fams}}
    steps, iters, n = 10, 4, 1_000_000
    prof["ms"].update({"lambda": 13.0, "delta": 4.0, "reorder": 0.26, "sort": 1.1})
    prof["launches"].update({"lambda": steps * iters, "delta": steps * iters, "reorder": steps, "sort": steps * 15})
    r = bench.roofline_of(prof, n, iters, steps)
    assert r["bound"] == "hbm" and r["kernel"] == "lambda" and r["unit"] == "GB/s"
    assert r["avg_launch_ms"] == pytest.approx(13.0 / 40)
    assert r["achieved"] == pytest.approx(n * bench.ALG_BYTES["lambda"] / (13.0 / 40 * 1e-3) / 1e9)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"]) and r["peak"] > 1000
    assert r["achieved_GBps_by_family"]["reorder"] == pytest.approx(n * 148 / (0.026e-3) / 1e9, rel=1e-3)
    assert r["traffic"] is None or r["traffic"] > 1e6  # the committed ncu capture is for 1 M particles
    assert bench.roofline_of(prof, 12345, iters, steps)["traffic"] is None  # no capture for other sizes


def test_clock_sampler_disabled_and_summary(bench):
    with bench.ClockSampler(0, enabled=False) as c:
        pass
    s = c.summary()
    assert s["samples"] == 0 and s["sm_mhz"] is None and s["reasons"] == []
    c.rows = [["1965", "1965", "Not Active", "Not Active", "Not Active", "Active"], ["1800", "1965"] + ["Not Active"] * 4, ["x", "y"]]
    s = c.summary()
    assert s["samples"] == 2 and s["sm_max_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"]


def test_workloads(bench):
    name, desc, p, xs = bench.workload("auto", 1)
    assert name == "dam-1m" and len(xs) == 1_000_000 and p.iteration == 4 and not p.surface_enabled
    name, desc, p, xs = bench.workload("auto", 8)
    assert name == "dam-weak" and len(xs) == 200 ** 3
    assert bench.workload("dam-1m-mc", 1)[2].surface_enabled == 1
    with pytest.raises(SystemExit):
        bench.workload("nope", 1)


def test_reference_arm_line(oracle_mod):
    """--impl reference on a tiny budget: one JSON line with the contract's keys (times the real reference, or the port)."""
    code = ("import bench, sys, json\n"
            "from pbf_sph_b200 import scenes\n"
            "p, xs = scenes.dam_break(16, 2)\n"
            "print(json.dumps(bench.reference_cpu(p, xs, 2, 1, settle_steps=2, budget_s=5.0)))\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    pis, ms, cores, kind, sample, same, n_timed = json.loads(out.stdout.strip().splitlines()[-1])
    assert pis > 0 and ms > 0 and cores >= 1 and kind in ("reference", "port") and "settled" in sample
    assert isinstance(same, bool) and 0 < n_timed <= 16 ** 3
