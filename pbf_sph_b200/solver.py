"""Host-side mirror of the reference's solver interface (src/sph.hpp:119-125) over the C ABI.

``Solver(h, device).advance(params, xs)`` is ``sph::Solver::advance``: one PBF step on a host particle array, which
comes back in Z-sorted order with the optional marching-cubes mesh.  ``upload/step/download`` is the resident path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import PARTICLE, GridInfo, Params, Profile, check, lib


@dataclass
class Result:  # sph::Result — sph.hpp:114-117
    vs: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    ns: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    cs: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.float32))
    queries: list = field(default_factory=list)  # [(query id, ids of the fluid particles in its cell)] — sph.hpp:27-31


class Solver:
    def __init__(self, h: float = 0.1, device: int = 0, flags: int = 0):
        self._L = lib()
        self._ctx = C.c_void_p()
        check(self._L.pbf_create(C.byref(self._ctx), C.c_float(h), device))
        self.h = h
        if flags:
            self.set_flags(flags)

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._L.pbf_destroy(self._ctx)
            self._ctx = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc: int) -> None:
        check(rc, self._ctx)

    def unpin_host(self) -> None:
        """Release the page-lock FLAG_PIN_HOST holds on the caller's array (before freeing or reallocating it)."""
        self._ck(self._L.pbf_unpin_host(self._ctx))

    def set_flags(self, flags: int) -> None:
        self._ck(self._L.pbf_set_flags(self._ctx, flags))

    def set_list_capacity(self, hits: int) -> None:
        """Parity tests only: depth of the per-iteration neighbour list (192 by default, 96 the other compiled depth)."""
        self._ck(self._L.pbf_debug_set_list_capacity(self._ctx, hits))

    def debug_scan(self, values: np.ndarray):
        """Parity tests only: the device exclusive prefix sum on a host u32 array -> (prefix, total)."""
        v = np.ascontiguousarray(values, np.uint32)
        out = np.empty_like(v)
        total = C.c_uint32(0)
        self._ck(self._L.pbf_debug_scan_u32(self._ctx, v.ctypes.data_as(C.c_void_p), len(v), out.ctypes.data_as(C.c_void_p),
                                            C.byref(total)))
        return out, int(total.value)

    def debug_sort_pairs(self, keys: np.ndarray):
        """Parity tests only: the device radix sort of (key, index) pairs -> (sorted keys, stable permutation)."""
        k = np.ascontiguousarray(keys, np.uint32)
        ko, po = np.empty_like(k), np.empty_like(k)
        self._ck(self._L.pbf_debug_sort_pairs(self._ctx, k.ctypes.data_as(C.c_void_p), len(k), ko.ctypes.data_as(C.c_void_p),
                                              po.ctypes.data_as(C.c_void_p)))
        return ko, po

    def set_stream(self, cuda_stream: int) -> None:
        self._ck(self._L.pbf_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    # ---- drop-in: sph::Solver::advance ---------------------------------------------------------------------
    def advance(self, params: Params, xs: np.ndarray, mesh: bool = True) -> Result:
        assert xs.dtype == PARTICLE and xs.flags.c_contiguous
        nv = C.c_uint64(0)
        self._ck(self._L.pbf_advance_host(self._ctx, C.byref(params), xs.ctypes.data, len(xs), C.byref(nv)))
        return self.mesh() if (mesh and nv.value) else Result()

    def advance_scene(self, params: Params, scene, xs: np.ndarray, mesh: bool = True):
        """advance(config, scene, xs) with a non-empty sph::Scene (capi.Scene): returns (xs after the call — sources
        may have appended, drains removed —, Result with the query answers)."""
        assert xs.dtype == PARTICLE and xs.flags.c_contiguous
        cap = len(xs) + scene.emitted()
        buf = np.zeros(cap, PARTICLE)
        buf[: len(xs)] = xs
        nv, n_out = C.c_uint64(0), C.c_uint64(0)
        self._ck(self._L.pbf_advance_scene_host(self._ctx, C.byref(params), C.byref(scene.struct), buf.ctypes.data, len(xs),
                                                cap, C.byref(n_out), C.byref(nv)))
        res = self.mesh() if (mesh and nv.value) else Result()
        res.queries = self.query_results(scene) if n_out.value else [(q.id, np.zeros(0, np.uint64)) for q in scene.queries]
        return buf[: n_out.value], res

    def set_scene(self, scene) -> None:
        """Scene of the following step() calls on the resident path (None = empty)."""
        self._ck(self._L.pbf_set_scene(self._ctx, C.byref(scene.struct) if scene is not None else None))

    def query_results(self, scene) -> list:
        out = []
        for i, q in enumerate(scene.queries):
            cnt = C.c_uint64(0)
            self._L.pbf_query_result(self._ctx, i, None, 0, C.byref(cnt))  # count only (CAPACITY status when > 0)
            ids = np.zeros(cnt.value, np.uint64)
            if cnt.value:
                self._ck(self._L.pbf_query_result(self._ctx, i, ids.ctypes.data, cnt.value, C.byref(cnt)))
            out.append((q.id, ids))
        return out

    def advance_ptr(self, params: Params, ptr: int, n: int) -> int:
        """advance() on a raw host pointer (e.g. pinned memory); returns the mesh vertex count."""
        nv = C.c_uint64(0)
        self._ck(self._L.pbf_advance_host(self._ctx, C.byref(params), C.c_void_p(ptr), n, C.byref(nv)))
        return nv.value

    def mesh(self) -> Result:
        n = self.grid().n_triangles * 3
        r = Result(np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 4), np.float32))
        if n:
            self._ck(self._L.pbf_mesh_download(self._ctx, r.vs.ctypes.data, r.ns.ctypes.data, r.cs.ctypes.data, n))
        return r

    # ---- resident path ------------------------------------------------------------------------------------------
    def upload(self, xs: np.ndarray) -> None:
        assert xs.dtype == PARTICLE and xs.flags.c_contiguous
        self._ck(self._L.pbf_upload(self._ctx, xs.ctypes.data, len(xs)))

    def step(self, params: Params) -> None:
        self._ck(self._L.pbf_step(self._ctx, C.byref(params)))

    def sync(self) -> None:
        self._ck(self._L.pbf_sync(self._ctx))

    def count(self) -> int:
        n = C.c_uint64(0)
        self._ck(self._L.pbf_particle_count(self._ctx, C.byref(n)))
        return n.value

    def download(self) -> np.ndarray:
        xs = np.zeros(self.count(), PARTICLE)
        n = C.c_uint64(0)
        self._ck(self._L.pbf_download(self._ctx, xs.ctypes.data, len(xs), C.byref(n)))
        return xs

    # ---- introspection ------------------------------------------------------------------------------------------
    def grid(self) -> GridInfo:
        g = GridInfo()
        self._ck(self._L.pbf_grid(self._ctx, C.byref(g)))
        return g

    _TAP_DTYPE = {capi.TAP_LAMBDA: np.float32, capi.TAP_RHO: np.float32, capi.TAP_IDS: np.uint64,
                  capi.TAP_MC_FIELD: np.float32, capi.TAP_MC_COLOUR: np.float32}

    def tap(self, tap: int) -> np.ndarray:
        g = self.grid()
        n = int(g.n_particles)
        if tap == capi.TAP_CELL_TABLE:
            shape = (g.grid_table_n,)
        elif tap in (capi.TAP_MC_FIELD, capi.TAP_MC_COLOUR):
            shape = (int(g.sample_size[0]) * int(g.sample_size[1]) * int(g.sample_size[2]), 4)
        else:
            shape = (n,)
        out = np.zeros(shape, self._TAP_DTYPE.get(tap, np.uint32))
        self._ck(self._L.pbf_debug_read(self._ctx, tap, out.ctypes.data, out.nbytes))
        return out

    def profile_reset(self) -> None:
        self._ck(self._L.pbf_profile_reset(self._ctx))

    def profile_mask(self, families=None) -> None:
        """Families timed under FLAG_PROFILE (names from capi.PHASES; None = all)."""
        mask = 0xFFFFFFFF if families is None else sum(1 << capi.PHASES.index(f) for f in families)
        self._ck(self._L.pbf_profile_set_mask(self._ctx, mask))

    def profile(self) -> dict:
        p = Profile()
        self._ck(self._L.pbf_profile_read(self._ctx, C.byref(p)))
        return {"steps": p.steps, "ms": {name: p.ms[i] for i, name in enumerate(capi.PHASES)},
                "launches": {name: p.launches[i] for i, name in enumerate(capi.PHASES)}}

    def launch_count(self) -> int:
        return self._L.pbf_launch_count(self._ctx)
