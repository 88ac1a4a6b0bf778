"""The `benchmark` driver (src/benchmark.cpp): the reference's flags and summary lines (reference benchmark.cpp:91-101,
args.cpp:7-50) through the C++ adaptor sph::cuda_impl::Solver (include/pbf/cudasph.hpp) — i.e. the C++ host path of the
drop-in boundary, including Result/mesh hand-off and the cloud.ply / mesh.obj writers."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from pbf_sph_b200 import Solver, scenes

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "build" / "benchmark"


def run(*args):
    if not EXE.exists():
        subprocess.run(["make", "-C", str(ROOT / "pbf_sph_b200" / "csrc")], check=True, capture_output=True)
    return subprocess.run([str(EXE), *args], capture_output=True, text=True, timeout=300)


def test_benchmark_stock_scene_matches_python_mirror(gpu, tmp_path):
    out = run("-i", "cuda", "-n", "4", "-w", "3", "-o", str(tmp_path / "out_{impl}_{type}_{iter}"))
    assert out.returncode == 0, out.stderr
    for label in ("Benchmark completed after 4 frames", "Runtime", "Framerate", "Frame-time min", "Frame-time max",
                  "Frame-time mean", "Frame-time stdDev", "Final Vertex count", "Final Particle count : 18522", "Results flushed."):
        assert label in out.stdout, out.stdout
    d = tmp_path / "out_cuda_fp32_4"
    assert (d / "cloud.ply").exists() and (d / "mesh.obj").exists()
    # the same frames through the Python mirror of the same C ABI: identical final particles
    p, xs = scenes.two_cubes(20000, 6)
    p.surface_enabled = 1
    with Solver(scenes.H, 0) as s:
        for f in list(range(3)) + list(range(4)):  # warm-up frames 0..2, then timed frames 0..3 (benchmark.cpp:31-54)
            res = s.advance(scenes.apply_motion(p, f), xs)
    raw = (d / "cloud.ply").read_bytes()
    body = raw[raw.index(b"end_header\n") + len(b"end_header\n"):]
    rec = np.frombuffer(body, dtype=np.dtype([("xyz", "<f4", 3), ("rgba", "<f4", 4), ("id", "<u4")]))
    assert len(rec) == len(xs)
    assert np.array_equal(rec["id"], xs["id"].astype(np.uint32))
    assert np.array_equal(rec["xyz"], xs["position"])
    nv = int(re.search(r"Final Vertex count\s*:\s*(\d+)", out.stdout).group(1))
    assert nv == len(res.vs) and nv > 1000
    assert sum(1 for line in (d / "mesh.obj").read_text().splitlines() if line.startswith("v ")) == nv


def test_benchmark_flags(gpu, tmp_path):
    assert run("--fp64").returncode != 0                      # fp32 only, like the OpenCL backend
    assert run("-i", "omp").returncode != 0
    assert "CUDA device 0" in run("-l").stdout
    out = run("--scene=dam", "--particles=27000", "--solver-iters=4", "--surface=off", "--resident", "-n3", "-w2", "-o", "")
    assert out.returncode == 0 and "Final Particle count : 27000" in out.stdout and "Final Vertex count   : 0" in out.stdout


def test_benchmark_with_a_scene_matches_python_mirror(gpu, tmp_path):
    """--fountain: a non-empty sph::Scene through sph::cuda_impl::Solver::advance (sources resize the caller's vector,
    drains shrink it, Result::queries is filled) — same frames through the Python mirror give the same particles."""
    from helpers import demo_scene
    out = run("--fountain", "--particles=2000", "--solver-iters=3", "--surface=off", "-n", "4", "-w", "2", "-o",
              str(tmp_path / "f_{iter}"))
    assert out.returncode == 0, out.stdout + out.stderr
    sc = demo_scene()
    p, xs = scenes.two_cubes(2000, 3)
    with Solver(scenes.H, 0) as s:
        for f in list(range(2)) + list(range(4)):
            xs, res = s.advance_scene(scenes.apply_motion(p, f), sc, xs)
    assert f"Final Particle count : {len(xs)} " in out.stdout, out.stdout
    want = " ".join(f"{qid}:{len(ids)}" for qid, ids in res.queries)
    assert f"Query answers        : {want}" in out.stdout, out.stdout
    raw = (tmp_path / "f_4" / "cloud.ply").read_bytes()
    body = raw[raw.index(b"end_header\n") + len(b"end_header\n"):]
    rec = np.frombuffer(body, dtype=np.dtype([("xyz", "<f4", 3), ("rgba", "<f4", 4), ("id", "<u4")]))
    assert np.array_equal(rec["id"], xs["id"].astype(np.uint32)) and np.array_equal(rec["xyz"], xs["position"])
    assert run("--fountain", "--resident", "-n1", "-w0", "-o", "").returncode != 0


def test_checkpoint_resume_is_bit_exact(gpu, tmp_path):
    """--save-state / --load-state: 3 + 3 frames through a checkpoint file equal 6 frames in one go, bit for bit (the
    step is deterministic and the checkpoint holds exactly what crosses advance())."""
    common = ["--scene=dam", "--particles=27000", "--solver-iters=3", "--surface=off", "-w", "0"]
    assert run(*common, "-n", "6", "--save-state", str(tmp_path / "six.bin"), "-o", "").returncode == 0
    assert run(*common, "-n", "3", "--save-state", str(tmp_path / "three.bin"), "-o", "").returncode == 0
    out = run(*common, "-n", "3", "--load-state", str(tmp_path / "three.bin"), "--save-state", str(tmp_path / "resumed.bin"), "-o", "")
    assert out.returncode == 0 and "Final Particle count : 27000" in out.stdout
    six, resumed = (tmp_path / "six.bin").read_bytes(), (tmp_path / "resumed.bin").read_bytes()
    assert six[:8] == b"PBFSTATE" and len(six) == 16 + 27000 * 56
    assert six == resumed
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"nonsense")
    assert run(*common, "-n", "1", "--load-state", str(bad), "-o", "").returncode != 0


def test_pinned_caller_memory_gives_the_same_particles(gpu, tmp_path):
    """PBF_FLAG_PIN_HOST (the benchmark driver's default, --no-pin turns it off): the library page-locks the caller's
    array for the copies of advance().  Same bytes out, through the C ABI with pageable numpy arrays (also when the
    array changes between calls and after an explicit unpin) and through the C++ adaptor's std::vector."""
    from pbf_sph_b200 import FLAG_PIN_HOST
    p, xs = scenes.dam_break(40, 3)  # 64 000 particles = 3.6 MB: above the 1 MB registration threshold
    plain, pinned, other = xs.copy(), xs.copy(), xs.copy()
    with Solver(scenes.H, 0) as s:
        for _ in range(3):
            s.advance(p, plain)
    with Solver(scenes.H, 0, FLAG_PIN_HOST) as s:
        s.advance(p, pinned)
        s.advance(p, pinned)          # the same array again: the lock is reused
        s.advance(p, other)           # another array: the lock moves
        s.unpin_host()
        s.advance(p, pinned)
        del other                     # (released above: freeing is safe)
    assert pinned.tobytes() == plain.tobytes()
    outs = []
    for extra in ((), ("--no-pin",)):
        tag = "nopin" if extra else "pin"
        out = run("--scene=dam", "--particles=64000", "--solver-iters=3", "--surface=off", "-n3", "-w2",
                  "-o", str(tmp_path / (tag + "_{iter}")), *extra)
        assert out.returncode == 0, out.stderr
        outs.append((tmp_path / (tag + "_3") / "cloud.ply").read_bytes())
    assert outs[0] == outs[1]


def test_device_list_runs_the_slab_path_and_matches_one_device(gpu, tmp_path):
    """-d is a device LIST like the reference's (args.cpp:20-23); more than one entry = the Z-curve slab decomposition
    behind the same sph::Solver::advance.  On this 1-GPU box the four ranks share device 0: cloud.ply must be
    byte-identical to the single-device run's."""
    common = ["--scene=dam", "--particles=64000", "--solver-iters=4", "--surface=off", "-n", "4", "-w", "3"]
    one = run(*common, "-d", "0", "-o", str(tmp_path / "one_{iter}"))
    four = run(*common, "-d", "0,0", "-d", "0", "-d0", "-o", str(tmp_path / "four_{iter}"))
    assert one.returncode == 0 and four.returncode == 0, one.stderr + four.stderr
    assert "Final Particle count : 64000" in four.stdout
    a, b = (tmp_path / "one_4" / "cloud.ply").read_bytes(), (tmp_path / "four_4" / "cloud.ply").read_bytes()
    assert len(a) > 64000 * 32 and a == b
    # the stock scene with its moving wall and the surface on (the stock benchmark, benchmark.cpp:29): particles AND mesh of
    # the two-rank run are the single-device files byte for byte
    stock = ["--particles=20000", "--solver-iters=4", "-n", "3", "-w", "2"]
    one = run(*stock, "-o", str(tmp_path / "s1_{iter}"))
    two = run(*stock, "-d", "0,0", "-o", str(tmp_path / "s2_{iter}"))
    assert one.returncode == 0 and two.returncode == 0, one.stderr + two.stderr
    assert (tmp_path / "s1_3" / "cloud.ply").read_bytes() == (tmp_path / "s2_3" / "cloud.ply").read_bytes()
    mesh = (tmp_path / "s1_3" / "mesh.obj").read_bytes()
    assert len(mesh) > 10000 and mesh == (tmp_path / "s2_3" / "mesh.obj").read_bytes()
    assert run("--gpus=2", "--resident", "-n1", "-w0", "-o", "").returncode != 0
    assert run("-d", "0,99", "-n1", "-w0", "-o", "").returncode != 0  # no such device
