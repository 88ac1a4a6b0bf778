"""The opt-in XSPH viscosity / vorticity confinement extension (PBF_FLAG_XSPH, PBF_FLAG_VORTICITY; csrc/xsph.cu).

No reference backend implements either term (src/sph_constants.h:13-14 only declares the constants — SURVEY F1), so
parity is against this repo's own definition in oracle/pbf_oracle.c (modes XSPH / VORTICITY), which the kernels
follow operation for operation, plus properties the definition must have."""
import numpy as np
import pytest

from helpers import DOMAIN, frac_within, warm
from pbf_sph_b200 import FLAG_STRICT_FP, FLAG_VORTICITY, FLAG_XSPH, Solver, capi, scenes

pytestmark = pytest.mark.gpu
H = scenes.H


@pytest.fixture(scope="module")
def warm_state(oracle_mod):
    p, xs = scenes.two_cubes(8000, 4)
    warm(oracle_mod, H, p, xs, 25, motion=scenes.apply_motion)
    return p, xs


@pytest.mark.parametrize("flags,mode", [(FLAG_XSPH, "XSPH"), (FLAG_VORTICITY, "VORTICITY"), (FLAG_XSPH | FLAG_VORTICITY, "BOTH")])
def test_extension_matches_oracle(gpu, oracle_mod, warm_state, flags, mode):
    p, snap = warm_state
    pf = scenes.apply_motion(p, 25)
    omode = {"XSPH": oracle_mod.XSPH, "VORTICITY": oracle_mod.VORTICITY, "BOTH": oracle_mod.XSPH | oracle_mod.VORTICITY}[mode]
    cpu, plain = snap.copy(), snap.copy()
    oracle_mod.step(H, pf, cpu, mode=omode)
    oracle_mod.step(H, pf, plain)
    assert np.array_equal(cpu["position"], plain["position"]), "the extension acts on velocities only"
    assert not np.array_equal(cpu["velocity"], plain["velocity"])
    got = snap.copy()
    with Solver(H, 0, flags | FLAG_STRICT_FP) as s:
        s.advance(pf, got)
    assert np.array_equal(got["id"], cpu["id"])
    # strict arithmetic: only pow(x, 4) of the delta pass differs from the oracle by design (pair_math.cuh)
    assert frac_within(got["position"], cpu["position"], 1e-7 * DOMAIN) >= 0.999
    dv_ext = np.abs(cpu["velocity"] - plain["velocity"]).max()
    assert np.abs(got["velocity"] - cpu["velocity"]).max() <= max(1e-4, 1e-3 * dv_ext)
    # and the production arithmetic stays within the usual one-step tolerance
    fast = snap.copy()
    with Solver(H, 0, flags) as s:
        s.advance(pf, fast)
    assert frac_within(fast["position"], cpu["position"], 1e-5 * DOMAIN) >= 0.999
    assert np.abs(fast["velocity"] - cpu["velocity"]).max() <= 1e-2


def test_default_is_off_and_xsph_properties(gpu, warm_state):
    """Flags off: bit-identical to the plain step.  XSPH on: total momentum is conserved (symmetric weights,
    antisymmetric velocity differences) and the kinetic energy of the relative motion does not grow."""
    p, snap = warm_state
    pf = scenes.apply_motion(p, 25)
    outs = {}
    for name, flags in (("plain", 0), ("zero", 0), ("xsph", FLAG_XSPH)):
        xs = snap.copy()
        with Solver(H, 0, flags) as s:
            s.advance(pf, xs)
        outs[name] = xs
    assert outs["plain"].tobytes() == outs["zero"].tobytes()
    v0, v1 = outs["plain"]["velocity"].astype(np.float64), outs["xsph"]["velocity"].astype(np.float64)
    assert np.array_equal(outs["plain"]["position"], outs["xsph"]["position"])
    assert np.allclose(v0.sum(0), v1.sum(0), rtol=0, atol=1e-4 * np.abs(v0).sum(0).max())
    ke = lambda v: 0.5 * ((v - v.mean(0)) ** 2).sum()
    assert ke(v1) < ke(v0)


def test_extension_rejected_on_the_slab_path(gpu):
    import ctypes as C
    p, xs = scenes.two_cubes(2000, 2)
    L = capi.lib()
    ctxs = (C.c_void_p * 2)()
    for i in range(2):
        assert L.pbf_create(C.byref(C.c_void_p.from_buffer(ctxs, i * C.sizeof(C.c_void_p))), C.c_float(H), 0) == 0
    try:
        assert L.pbf_dist_init_local(ctxs, 2) == 0
        half = len(xs) // 2
        assert L.pbf_dist_upload(ctxs[0], xs[:half].ctypes.data, half) == 0
        assert L.pbf_dist_upload(ctxs[1], xs[half:].ctypes.data, len(xs) - half) == 0
        assert L.pbf_set_flags(ctxs[0], FLAG_XSPH) == 0
        assert L.pbf_dist_step(ctxs[0], C.byref(p)) == -4  # PBF_ERR_STATE
        assert L.pbf_set_flags(ctxs[0], 0) == 0
        assert L.pbf_dist_step(ctxs[0], C.byref(p)) == 0
    finally:
        for i in range(2):
            L.pbf_destroy(ctxs[i])
