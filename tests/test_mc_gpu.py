"""Marching-cubes path (SphParams::surface) on the GPU against the oracle — ompsph.hpp:277-477.

Lattice field / normals / colours within float tolerance (CUDA powf vs glibc powf differ by ulps), NaN pattern identical
(lattice points with no particle in range carry NaN normals and colours in the reference), per-cube triangle counts and
the total exact, and the mesh — which both sides emit in ascending cube order — compared vertex by vertex."""
import numpy as np
import pytest

from helpers import warm
from pbf_sph_b200 import Solver, capi, scenes

pytestmark = pytest.mark.gpu
H = scenes.H


@pytest.mark.parametrize("count,frames", [(4000, 8), (20000, 25)])
def test_surface_extraction_matches_oracle(gpu, oracle_mod, count, frames):
    p, xs = scenes.two_cubes(count, 3)
    warm(oracle_mod, H, p, xs, frames, motion=scenes.apply_motion)
    pf = scenes.apply_motion(p, frames)
    pf.surface_enabled = 1
    cpu, dev = xs.copy(), xs.copy()
    t = oracle_mod.step(H, pf, cpu, taps=True)
    with Solver(H, 0, capi.FLAG_STRICT_FP) as s:
        res = s.advance(pf, dev)
        field, colour = s.tap(capi.TAP_MC_FIELD), s.tap(capi.TAP_MC_COLOUR)
        g = s.grid()
    assert list(g.sample_size) == list(t["grid"].sample_size)
    assert np.array_equal(np.isnan(field), np.isnan(t["mc_field"]))
    assert np.array_equal(np.isnan(colour), np.isnan(t["mc_colour"]))
    assert np.allclose(field[:, 0], t["mc_field"][:, 0], rtol=2e-5, atol=1e-4)
    assert np.allclose(field[:, 1:], t["mc_field"][:, 1:], rtol=0, atol=2e-4, equal_nan=True)  # unit normals
    assert np.allclose(colour, t["mc_colour"], rtol=1e-5, atol=1e-6, equal_nan=True)
    # triangle counts: a lattice value within float noise of the isolevel could flip a cube; report if it does
    assert g.n_triangles * 3 == len(res.vs)
    assert len(res.vs) == t["n_vertices"], (len(res.vs), t["n_vertices"])
    assert len(res.vs) > 1000, "scene must actually produce a surface"
    assert np.allclose(res.vs, t["mesh_vs"], rtol=0, atol=2e-2, equal_nan=True)   # positions, domain = 1000
    assert np.allclose(res.ns, t["mesh_ns"], rtol=0, atol=2e-3, equal_nan=True)
    assert np.allclose(res.cs, t["mesh_cs"], rtol=0, atol=2e-4, equal_nan=True)


def test_surface_off_gives_empty_mesh(gpu):
    p, xs = scenes.two_cubes(2000, 2)
    with Solver(H, 0) as s:
        res = s.advance(p, xs)
        assert len(res.vs) == 0 and s.grid().n_triangles == 0


def test_device_mesh_handoff(gpu):
    """pbf_mesh_device: the mesh where it was produced (for a renderer mapping CUDA memory) equals the host download."""
    import ctypes as C
    p, xs = scenes.two_cubes(4000, 3)
    p.surface_enabled = 1
    with Solver(H, 0) as s:
        for f in range(6):
            res = s.advance(scenes.apply_motion(p, f), xs)
        vs, ns, cs, nv = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64(0)
        s._ck(s._L.pbf_mesh_device(s._ctx, C.byref(vs), C.byref(ns), C.byref(cs), C.byref(nv)))
        assert nv.value == len(res.vs) > 0 and vs.value and ns.value and cs.value
        rt = None
        for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                rt = C.CDLL(name)
                break
            except OSError:
                continue
        assert rt is not None, "CUDA runtime not found"
        for ptr, want in ((vs, res.vs), (ns, res.ns), (cs, res.cs)):
            host = np.empty_like(want)
            assert rt.cudaMemcpy(C.c_void_p(host.ctypes.data), ptr, C.c_size_t(host.nbytes), 2) == 0  # DeviceToHost
            assert np.array_equal(host, want, equal_nan=True)
