"""The C-ABI library loads without a GPU, exports every symbol include/pbf_cuda.h declares, and its host-side
logic (grid set-up, Morton curve, scene factory, constants) matches the reference's definitions."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from pbf_sph_b200 import capi, scenes

ROOT = Path(__file__).resolve().parent.parent


def test_every_declared_symbol_is_exported():
    header = (ROOT / "include" / "pbf_cuda.h").read_text()
    declared = set(re.findall(r"\b(pbf_[a-z0-9_]+)\s*\(", header)) - {"pbf_ctx"}
    L = capi.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    assert L.pbf_abi_version() == 2


def test_struct_layouts():
    assert capi.PARTICLE.itemsize == 56  # sizeof(sph::Particle<size_t,float,glm::vec>)
    assert [capi.PARTICLE.fields[f][1] for f in ("id", "type", "mass", "position", "velocity", "colour")] == \
        [0, 8, 12, 16, 28, 40]
    assert C.sizeof(capi.Params) == 80 and C.sizeof(capi.McParams) == 16


def test_ctypes_mirrors_match_the_header(tmp_path):
    """Every struct of include/pbf_cuda.h that crosses the boundary has the size and field offsets of its ctypes mirror
    (compiled as plain C: the header is the contract a cgo / JNI / ctypes binding reads)."""
    import subprocess
    pairs = {"pbf_params": capi.Params, "pbf_mc_params": capi.McParams, "pbf_dist_stats": capi.DistStats,
             "pbf_profile": capi.Profile, "pbf_well": capi.Well, "pbf_source": capi.Source, "pbf_drain": capi.Drain,
             "pbf_query": capi.Query, "pbf_grid_info": capi.GridInfo}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "pbf_cuda.h"', "int main(void) {"]
    for cname, mirror in pairs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ['  printf("PH %d %d %d\\n", PBF_PH_HALO, PBF_PH_SLAB_ITERATIONS, PBF_PH_COUNT);', "  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    for line, (cname, mirror) in zip(out, pairs.items()):
        got = line.split()
        assert got[0] == cname
        want = [C.sizeof(mirror)] + [getattr(mirror, f).offset for f, _ in mirror._fields_]
        assert [int(v) for v in got[1:]] == want, (cname, got[1:], want)
    ph = out[len(pairs)].split()
    assert ph[0] == "PH" and int(ph[1]) == capi.PHASES.index("halo") and int(ph[2]) == capi.PHASES.index("slab_iterations")
    assert int(ph[3]) == capi.PH_COUNT == len(capi.PHASES)


def test_no_gpu_fails_loudly():
    L = capi.lib()
    ctx = C.c_void_p()
    rc = L.pbf_create(C.byref(ctx), C.c_float(0.1), 0)
    if rc == 0:
        L.pbf_destroy(ctx)
        pytest.skip("a GPU is present")
    assert rc == -2 and b"no CPU fallback" in L.pbf_last_error(None)
    with pytest.raises(capi.PbfError):
        from pbf_sph_b200 import Solver
        Solver(0.1, 0)


def test_morton_known_answers_and_round_trip():
    L = capi.lib()
    assert L.pbf_host_morton_encode(24, 24, 24) == 32256  # G of the stock scene
    assert L.pbf_host_morton_encode(23, 24, 24) == 31817
    assert L.pbf_host_morton_encode(1, 0, 0) == 1 and L.pbf_host_morton_encode(0, 1, 0) == 2
    assert L.pbf_host_morton_encode(0, 0, 1) == 4 and L.pbf_host_morton_encode(1023, 1023, 1023) == (1 << 30) - 1
    assert L.pbf_host_morton_encode(0xFFFFFFFF, 0, 0) == L.pbf_host_morton_encode(1023, 0, 0)  # x-1 at 0 wraps
    assert L.pbf_host_morton_encode(1024, 5, 5) == L.pbf_host_morton_encode(0, 5, 5)           # x+1 at 1023 wraps
    xyz = (C.c_uint32 * 3)()
    rng = np.random.default_rng(1)
    for x, y, z in np.vstack([rng.integers(0, 1024, (3000, 3)), [[0, 0, 0], [1023, 0, 1023]]]):
        L.pbf_host_morton_decode(L.pbf_host_morton_encode(int(x), int(y), int(z)), xyz)
        assert tuple(xyz) == (x, y, z)


def test_scene_factory_known_answers():
    p, xs = scenes.two_cubes(20000, 6)
    assert len(xs) == 18522 and np.array_equal(xs["id"], np.arange(18522))  # 2 x 21^3, sph.hpp:165-166
    assert tuple(xs["position"][0]) == (100.0, 0.0, 100.0) and tuple(xs["position"][1]) == (100.0, 0.0, 122.0)
    assert tuple(xs["position"][9261]) == (600.0, 0.0, 600.0)
    assert abs(p.dt - 0.01245) < 1e-7 and p.scale == 500.0 and p.iteration == 6
    p2, d = scenes.dam_break(100, 4)
    assert len(d) == 1_000_000
    g = capi.GridInfo()
    assert capi.lib().pbf_host_grid(C.c_float(0.1), C.byref(p2), C.byref(g)) == 0
    assert list(g.extent) == [95, 51, 51] and g.grid_table_n == 488063
    g0 = capi.GridInfo()
    capi.lib().pbf_host_grid(C.c_float(0.1), C.byref(p), C.byref(g0))
    assert list(g0.extent) == [24, 24, 24] and g0.grid_table_n == 32256


def test_host_grid_and_constants_equal_oracle(oracle_mod):
    p, _ = scenes.two_cubes(20000, 6)
    p.surface_enabled = 1
    for f in range(0, 200, 7):  # the moving wall flips the x extent between 23 and 24
        pf = scenes.apply_motion(p, f)
        g = capi.GridInfo()
        capi.lib().pbf_host_grid(C.c_float(0.1), C.byref(pf), C.byref(g))
        o = oracle_mod.grid(0.1, pf)
        assert list(g.extent) == list(o.extent) and g.grid_table_n == o.grid_table_n
        assert list(g.min_extent) == list(o.min_extent) and list(g.sample_size) == list(o.sample_size)
    ours, theirs = np.zeros(5, np.float32), np.zeros(3, np.float32)
    capi.lib().pbf_host_constants(C.c_float(0.1), ours.ctypes.data)
    oracle_mod.lib().pbf_oracle_constants(C.c_float(0.1), theirs.ctypes.data_as(C.c_void_p))
    assert np.array_equal(ours[:3], theirs)
    assert np.sqrt(np.float32(ours[3])) <= np.float32(0.1) < np.sqrt(np.nextafter(np.float32(ours[3]), np.float32(1)))
