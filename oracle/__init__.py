"""ctypes binding of the CPU checker (TEST INFRASTRUCTURE — never imported by the product package).

``liboracle.so``            oracle/pbf_oracle.c, the C restatement of ompsph.hpp:85-485
``_ref/libpbf_ref_*.so``    the unmodified reference OpenMP backend (oracle/ref/ref_driver.cpp), built in
                            the container that has /root/reference and shipped prebuilt to the GPU box.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

import sys

HERE = Path(__file__).resolve().parent
if str(HERE.parent) not in sys.path:
    sys.path.insert(0, str(HERE.parent))

# the POD types of include/pbf_cuda.h have ONE Python definition, in the product's binding
from pbf_sph_b200.capi import PARTICLE, GridInfo, McParams, Params, SceneStruct  # noqa: E402,F401

GAUSS_SEIDEL = 1
SKIP_DIFFUSE = 2
XSPH = 4       # extension modes: not part of any reference backend
VORTICITY = 8


class OracleIO(C.Structure):
    _fields_ = [("keys_input", C.c_void_p), ("perm", C.c_void_p), ("keys_sorted", C.c_void_p),
                ("cell_table", C.c_void_p), ("cell_table_cap", C.c_uint64),
                ("cand_count", C.c_void_p), ("nbr_count", C.c_void_p), ("lambda_", C.c_void_p), ("rho", C.c_void_p),
                ("mc_field", C.c_void_p), ("mc_colour", C.c_void_p), ("mc_lattice_cap", C.c_uint64),
                ("mesh_vs", C.c_void_p), ("mesh_ns", C.c_void_p), ("mesh_cs", C.c_void_p),
                ("mesh_cap_vertices", C.c_uint64), ("forced_perm", C.c_void_p),
                ("grid", GridInfo), ("n_vertices", C.c_uint64),
                ("scene", C.POINTER(SceneStruct)), ("query_first", C.c_void_p), ("query_count", C.c_void_p)]


def build(verbose: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists).  Building is not using."""
    subprocess.run(["make", "-C", str(HERE), "-j8"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL, stderr=None if verbose else subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = HERE / "liboracle.so"
        if not path.exists():
            build()
        _lib = C.CDLL(str(path))
        _lib.pbf_oracle_step.argtypes = [C.c_float, C.POINTER(Params), C.c_void_p, C.c_uint64, C.c_uint32,
                                         C.POINTER(OracleIO)]
        _lib.pbf_oracle_step.restype = C.c_int
        _lib.pbf_oracle_grid.argtypes = [C.c_float, C.POINTER(Params), C.POINTER(GridInfo)]
        _lib.pbf_oracle_scene_edit.argtypes = [C.c_float, C.POINTER(Params), C.POINTER(SceneStruct), C.c_void_p,
                                               C.c_uint64, C.c_uint64]
        _lib.pbf_oracle_scene_edit.restype = C.c_int64
        _lib.pbf_oracle_morton_encode.argtypes = [C.c_uint32] * 3
        _lib.pbf_oracle_morton_encode.restype = C.c_uint32
    return _lib


def grid(h: float, params: Params) -> GridInfo:
    g = GridInfo()
    lib().pbf_oracle_grid(C.c_float(h), C.byref(params), C.byref(g))
    return g


def scene_edit(h: float, params: Params, scene, xs: np.ndarray) -> np.ndarray:
    """Sources then drains on the particle list (ompsph.hpp:91-120); returns the edited list."""
    cap = len(xs) + scene.emitted()
    buf = np.zeros(cap, PARTICLE)
    buf[: len(xs)] = xs
    n = lib().pbf_oracle_scene_edit(C.c_float(h), C.byref(params), C.byref(scene.struct), buf.ctypes.data, len(xs), cap)
    if n < 0:
        raise RuntimeError(f"pbf_oracle_scene_edit failed: {n}")
    return buf[:n].copy()


def step(h: float, params: Params, xs: np.ndarray, mode: int = 0, taps: bool = False, mesh: bool = True,
         forced_perm: np.ndarray | None = None, scene=None) -> dict:
    """One oracle step.  ``xs`` (PARTICLE array) is advanced IN PLACE into Z-sorted order.

    Returns a dict with the grid info and, when ``taps``, every intermediate the parity tests compare."""
    assert xs.dtype == PARTICLE and xs.flags.c_contiguous
    n = len(xs)
    g = grid(h, params)
    io = OracleIO()
    keep = {}

    def buf(name, shape, dtype):
        a = np.zeros(shape, dtype=dtype)
        keep[name] = a
        return a.ctypes.data

    if taps:
        io.keys_input = buf("keys_input", n, np.uint32)
        io.perm = buf("perm", n, np.uint32)
        io.keys_sorted = buf("keys_sorted", n, np.uint32)
        io.cell_table = buf("cell_table", g.grid_table_n, np.uint32)
        io.cell_table_cap = g.grid_table_n
        io.cand_count = buf("cand_count", n, np.uint32)
        io.nbr_count = buf("nbr_count", n, np.uint32)
        io.lambda_ = buf("lambda", n, np.float32)
        io.rho = buf("rho", n, np.float32)
    if params.surface_enabled:
        L = int(g.sample_size[0]) * int(g.sample_size[1]) * int(g.sample_size[2])
        if taps:
            io.mc_field = buf("mc_field", (L, 4), np.float32)
            io.mc_colour = buf("mc_colour", (L, 4), np.float32)
            io.mc_lattice_cap = L
        if mesh:
            cap = 15 * max(1, (int(g.sample_size[0]) - 1) * (int(g.sample_size[1]) - 1) * (int(g.sample_size[2]) - 1))
            cap = min(cap, 6_000_000)
            io.mesh_vs = buf("mesh_vs", (cap, 3), np.float32)
            io.mesh_ns = buf("mesh_ns", (cap, 3), np.float32)
            io.mesh_cs = buf("mesh_cs", (cap, 4), np.float32)
            io.mesh_cap_vertices = cap
    if scene is not None:  # wells act in the prediction, queries are answered from the cell table
        io.scene = C.pointer(scene.struct)
        if scene.queries:
            io.query_first = buf("query_first", len(scene.queries), np.uint32)
            io.query_count = buf("query_count", len(scene.queries), np.uint32)
    if forced_perm is not None:
        fp = np.ascontiguousarray(forced_perm, dtype=np.uint32)
        keep["_forced"] = fp
        io.forced_perm = fp.ctypes.data
    rc = lib().pbf_oracle_step(C.c_float(h), C.byref(params), xs.ctypes.data, n, mode, C.byref(io))
    if rc != 0:
        raise RuntimeError(f"pbf_oracle_step failed: {rc}")
    out = {k: v for k, v in keep.items() if not k.startswith("_")}
    nv = int(io.n_vertices)
    for k in ("mesh_vs", "mesh_ns", "mesh_cs"):
        if k in out:
            out[k] = out[k][:nv]
    out["n_vertices"] = nv
    out["grid"] = io.grid
    return out


# ---------------------------------------------------------------- the real reference (oracle/_ref)
_ref = {}


def ref_available(variant: str = "strict") -> bool:
    return (HERE / "_ref" / f"libpbf_ref_{variant}.so").exists()


def ref_lib(variant: str = "strict") -> C.CDLL:
    """variant: strict | strict_stable | fast | native | best (native when this CPU has its ISA, else fast)."""
    if variant == "best":
        variant = "fast"
        flags_file = HERE / "_ref" / "native_cpu_flags.txt"
        if flags_file.exists() and (HERE / "_ref" / "libpbf_ref_native.so").exists():
            need = set(flags_file.read_text().split())
            have = set()
            for line in open("/proc/cpuinfo"):
                if line.startswith("flags"):
                    have = set(line.split(":", 1)[1].split())
                    break
            if need <= have:
                variant = "native"
    if variant not in _ref:
        path = HERE / "_ref" / f"libpbf_ref_{variant}.so"
        L = C.CDLL(str(path))
        L.pbf_ref_advance.argtypes = [C.c_float, C.POINTER(Params), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.pbf_ref_advance.restype = C.c_int
        L.pbf_ref_variant.restype = C.c_char_p
        L.pbf_ref_scene_2cubes.argtypes = [C.c_uint64, C.c_uint64, C.c_float, C.POINTER(Params), C.c_void_p, C.c_uint64]
        L.pbf_ref_scene_2cubes.restype = C.c_uint64
        L.pbf_ref_apply_motion.argtypes = [C.POINTER(Params), C.c_uint64, C.POINTER(Params)]
        L.pbf_ref_morton_encode.argtypes = [C.c_uint32] * 3
        L.pbf_ref_morton_encode.restype = C.c_uint32
        L.pbf_ref_set_threads.argtypes = [C.c_int]
        L.variant_name = variant
        _ref[variant] = L
    return _ref[variant]


def ref_advance(h: float, params: Params, xs: np.ndarray, variant: str = "strict", threads: int | None = None,
                mesh_cap: int = 0) -> dict:
    """sph::omp_impl::Solver<size_t,float>(h).advance(params, {}, xs) on the real reference; xs in place."""
    L = ref_lib(variant)
    if threads is not None:
        L.pbf_ref_set_threads(threads)
    nv = C.c_uint64(0)
    vs = np.zeros((mesh_cap, 3), np.float32)
    ns = np.zeros((mesh_cap, 3), np.float32)
    cs = np.zeros((mesh_cap, 4), np.float32)
    rc = L.pbf_ref_advance(C.c_float(h), C.byref(params), xs.ctypes.data, len(xs), vs.ctypes.data, ns.ctypes.data,
                           cs.ctypes.data, mesh_cap, C.byref(nv))
    if rc != 0:
        raise RuntimeError(f"pbf_ref_advance failed: {rc}")
    k = min(int(nv.value), mesh_cap)
    return {"n_vertices": int(nv.value), "mesh_vs": vs[:k], "mesh_ns": ns[:k], "mesh_cs": cs[:k]}


def ref_advance_scene(h: float, params: Params, scene, xs: np.ndarray, variant: str = "strict", threads: int | None = None):
    """advance(params, scene, xs) on the real reference; returns (particles after the call, [(query id, ids)])."""
    L = ref_lib(variant)
    if threads is not None:
        L.pbf_ref_set_threads(threads)
    cap = len(xs) + scene.emitted()
    buf = np.zeros(cap, PARTICLE)
    buf[: len(xs)] = xs
    n_out = C.c_uint64(0)
    q_ids = np.zeros(max(1, cap * max(1, len(scene.queries))), np.uint64)
    q_counts = np.zeros(max(1, len(scene.queries)), np.uint64)
    L.pbf_ref_advance_scene.argtypes = [C.c_float, C.POINTER(Params), C.POINTER(SceneStruct), C.c_void_p, C.c_uint64,
                                        C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64, C.c_void_p]
    rc = L.pbf_ref_advance_scene(C.c_float(h), C.byref(params), C.byref(scene.struct), buf.ctypes.data, len(xs), cap,
                                 C.byref(n_out), q_ids.ctypes.data, len(q_ids), q_counts.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"pbf_ref_advance_scene failed: {rc}")
    res, w = [], 0
    for i, q in enumerate(scene.queries):
        c = int(q_counts[i])
        res.append((q.id, q_ids[w:w + c].copy()))
        w += c
    return buf[: n_out.value].copy(), res


def ref_scene_2cubes(count: int, solver_iter: int, scaling: float = 500.0, variant: str = "strict"):
    L = ref_lib(variant)
    p = Params()
    n = L.pbf_ref_scene_2cubes(count, solver_iter, C.c_float(scaling), C.byref(p), None, 0)
    xs = np.zeros(n, PARTICLE)
    L.pbf_ref_scene_2cubes(count, solver_iter, C.c_float(scaling), C.byref(p), xs.ctypes.data, n)
    return p, xs


def ref_apply_motion(params: Params, frame: int, variant: str = "strict") -> Params:
    out = Params()
    ref_lib(variant).pbf_ref_apply_motion(C.byref(params), frame, C.byref(out))
    return out
