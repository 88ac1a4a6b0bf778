"""ctypes binding of include/pbf_cuda.h (libpbf_cuda.so: hand-written sm_100a kernels + C ABI).

There is no CPU fallback: if the shared library is missing it is built with nvcc, and if that fails, or no CUDA
device is present when a context is created, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libpbf_cuda.so"

# sph::Particle<size_t,float,glm::vec> — reference src/sph.hpp:36-54 (56 bytes)
PARTICLE = np.dtype(
    [("id", "<u8"), ("type", "u1"), ("_pad", "u1", (3,)), ("mass", "<f4"), ("position", "<f4", (3,)),
     ("velocity", "<f4", (3,)), ("colour", "<f4", (4,))]
)
assert PARTICLE.itemsize == 56

PBF_OK = 0
FLAG_STRICT_FP = 1 << 0
FLAG_DEBUG_COUNTS = 1 << 1
FLAG_PROFILE = 1 << 2
FLAG_GLOBAL_NEIGHBOURS = 1 << 3
FLAG_XSPH = 1 << 4       # extensions, default off (no reference backend has them)
FLAG_VORTICITY = 1 << 5
FLAG_PIN_HOST = 1 << 6  # page-lock the caller's array in the drop-in call (see include/pbf_cuda.h for the contract)

TAP_KEYS_INPUT, TAP_PERM, TAP_KEYS_SORTED, TAP_CELL_TABLE, TAP_CAND_COUNT, TAP_NBR_COUNT = range(6)
TAP_LAMBDA, TAP_RHO, TAP_IDS, TAP_MC_FIELD, TAP_MC_COLOUR, TAP_LIST_HITS = range(6, 12)

PHASES = ["predict_key", "sort", "reorder", "cell_table", "diffuse", "lambda", "delta", "finalise", "mc_field",
          "mc_count_scan", "mc_emit", "pack", "halo", "slab_setup", "slab_barrier", "slab_iterations"]
PH_COUNT = 16
NCCL_ID_BYTES = 128


class McParams(C.Structure):  # sph::McParams — sph.hpp:82-95
    _fields_ = [("resolution", C.c_float), ("isolevel", C.c_float), ("particle_size", C.c_float),
                ("particle_influence", C.c_float)]


class Params(C.Structure):  # sph::SphParams — sph.hpp:97-103
    _fields_ = [("dt", C.c_float), ("scale", C.c_float), ("iteration", C.c_uint64),
                ("constant_force", C.c_float * 3), ("min_bound", C.c_float * 3), ("max_bound", C.c_float * 3),
                ("wait", C.c_int32), ("surface_enabled", C.c_int32), ("surface", McParams)]

    def copy(self) -> "Params":
        out = Params()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(Params))
        return out


class GridInfo(C.Structure):
    _fields_ = [("min_extent", C.c_float * 3), ("extent", C.c_uint32 * 3), ("grid_table_n", C.c_uint32),
                ("key_bits", C.c_uint32), ("radix_passes", C.c_uint32), ("n_particles", C.c_uint64),
                ("sample_size", C.c_uint32 * 3), ("n_triangles", C.c_uint32)]


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * PH_COUNT), ("launches", C.c_uint64 * PH_COUNT), ("steps", C.c_uint64)]


class DistStats(C.Structure):
    _fields_ = [("owned", C.c_uint64), ("ghosts", C.c_uint64), ("migrants_out", C.c_uint64),
                ("migrants_in", C.c_uint64), ("halo_bytes_per_iteration", C.c_uint64), ("key_lo", C.c_uint32),
                ("key_hi", C.c_uint32), ("ghost_ring1", C.c_uint32), ("boundary", C.c_uint32), ("plan_steps", C.c_uint32),
                ("early_plans", C.c_uint32), ("capacity_owned", C.c_uint32), ("capacity_ghosts", C.c_uint32)]


class Well(C.Structure):  # sph::Well — sph.hpp:56-60
    _fields_ = [("tag", C.c_uint64), ("centre", C.c_float * 3), ("force", C.c_float)]


class Source(C.Structure):  # sph::Source — sph.hpp:62-67
    _fields_ = [("tag", C.c_uint64), ("centre", C.c_float * 3), ("velocity", C.c_float * 3), ("colour", C.c_float * 4),
                ("rate", C.c_float)]


class Drain(C.Structure):  # sph::Drain — sph.hpp:69-73
    _fields_ = [("tag", C.c_uint64), ("centre", C.c_float * 3), ("width", C.c_float), ("depth", C.c_float)]


class Query(C.Structure):  # sph::Query — sph.hpp:22-25
    _fields_ = [("id", C.c_uint64), ("point", C.c_float * 3)]


class SceneStruct(C.Structure):  # pbf_scene
    _fields_ = [("wells", C.POINTER(Well)), ("n_wells", C.c_uint32), ("sources", C.POINTER(Source)),
                ("n_sources", C.c_uint32), ("drains", C.POINTER(Drain)), ("n_drains", C.c_uint32),
                ("queries", C.POINTER(Query)), ("n_queries", C.c_uint32)]


class Scene:
    """sph::Scene — sph.hpp:75-80: python lists in, one pbf_scene (with the arrays kept alive) out."""

    def __init__(self, wells=(), sources=(), drains=(), queries=()):
        self.wells = [Well(t, (C.c_float * 3)(*c), f) for t, c, f in wells]
        self.sources = [Source(t, (C.c_float * 3)(*c), (C.c_float * 3)(*v), (C.c_float * 4)(*col), r)
                        for t, c, v, col, r in sources]
        self.drains = [Drain(t, (C.c_float * 3)(*c), w, d) for t, c, w, d in drains]
        self.queries = [Query(i, (C.c_float * 3)(*pt)) for i, pt in queries]
        self._arrays = ((Well * len(self.wells))(*self.wells), (Source * len(self.sources))(*self.sources),
                        (Drain * len(self.drains))(*self.drains), (Query * len(self.queries))(*self.queries))
        a = self._arrays
        self.struct = SceneStruct(C.cast(a[0], C.POINTER(Well)), len(self.wells), C.cast(a[1], C.POINTER(Source)),
                                  len(self.sources), C.cast(a[2], C.POINTER(Drain)), len(self.drains),
                                  C.cast(a[3], C.POINTER(Query)), len(self.queries))

    def emitted(self) -> int:
        """Particles the sources emit per call: floor(sqrt(rate)) * ceil(sqrt(rate)) each (ompsph.hpp:93-104)."""
        import math
        total = 0
        for s in self.sources:
            side = float(np.sqrt(np.float32(s.rate)))
            total += int(math.floor(side)) * int(math.ceil(side))
        return total


# every symbol include/pbf_cuda.h declares (tests/test_abi.py checks the library exports each one)
EXPORTS = [
    "pbf_create", "pbf_destroy", "pbf_last_error", "pbf_abi_version", "pbf_set_flags", "pbf_set_stream",
    "pbf_advance_host", "pbf_advance_scene_host", "pbf_unpin_host", "pbf_query_result", "pbf_set_scene", "pbf_mesh_download",
    "pbf_mesh_device", "pbf_upload",
    "pbf_step", "pbf_sync", "pbf_download",
    "pbf_particle_count", "pbf_device_state", "pbf_grid", "pbf_debug_read", "pbf_debug_set_list_capacity", "pbf_debug_scan_u32", "pbf_debug_sort_pairs", "pbf_profile_reset", "pbf_profile_read", "pbf_profile_set_mask",
    "pbf_launch_count", "pbf_dist_unique_id", "pbf_dist_init", "pbf_dist_init_local", "pbf_dist_upload",
    "pbf_dist_step", "pbf_dist_advance_host", "pbf_dist_download", "pbf_dist_set_replan", "pbf_dist_stats_read", "pbf_host_alloc", "pbf_host_free", "pbf_host_grid",
    "pbf_host_plan_splits", "pbf_host_work_weights", "pbf_host_constants", "pbf_host_morton_encode", "pbf_host_morton_decode",
    "pbf_host_apply_motion",
]


def build(verbose: bool = False) -> None:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> pbf_sph_b200/libpbf_cuda.so (in-tree)."""
    proc = subprocess.run(["make", "-C", str(HERE / "csrc"), "-j8"], capture_output=not verbose, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libpbf_cuda.so failed:\n" + (proc.stdout or "") + (proc.stderr or ""))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        build()
    L = C.CDLL(str(LIB_PATH))
    vp, u64, u32, i32, f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_float
    P = C.POINTER
    sig = {
        "pbf_create": ([P(vp), f32, i32], i32),
        "pbf_destroy": ([vp], None),
        "pbf_last_error": ([vp], C.c_char_p),
        "pbf_abi_version": ([], i32),
        "pbf_set_flags": ([vp, u32], i32),
        "pbf_set_stream": ([vp, vp], i32),
        "pbf_advance_host": ([vp, P(Params), vp, u64, P(u64)], i32),
        "pbf_advance_scene_host": ([vp, P(Params), P(SceneStruct), vp, u64, u64, P(u64), P(u64)], i32),
        "pbf_query_result": ([vp, u32, vp, u64, P(u64)], i32),
        "pbf_set_scene": ([vp, P(SceneStruct)], i32),
        "pbf_mesh_download": ([vp, vp, vp, vp, u64], i32),
        "pbf_mesh_device": ([vp, P(vp), P(vp), P(vp), P(u64)], i32),
        "pbf_upload": ([vp, vp, u64], i32),
        "pbf_step": ([vp, P(Params)], i32),
        "pbf_sync": ([vp], i32),
        "pbf_unpin_host": ([vp], i32),
        "pbf_download": ([vp, vp, u64, P(u64)], i32),
        "pbf_particle_count": ([vp, P(u64)], i32),
        "pbf_device_state": ([vp, P(vp), P(vp), P(vp), P(vp)], i32),
        "pbf_grid": ([vp, P(GridInfo)], i32),
        "pbf_debug_read": ([vp, i32, vp, u64], i32),
        "pbf_debug_set_list_capacity": ([vp, u32], i32),
        "pbf_debug_scan_u32": ([vp, vp, u64, vp, P(u32)], i32),
        "pbf_debug_sort_pairs": ([vp, vp, u32, vp, vp], i32),
        "pbf_profile_reset": ([vp], i32),
        "pbf_profile_read": ([vp, P(Profile)], i32),
        "pbf_profile_set_mask": ([vp, u32], i32),
        "pbf_launch_count": ([vp], u64),
        "pbf_dist_unique_id": ([vp], i32),
        "pbf_dist_init": ([vp, vp, i32, i32], i32),
        "pbf_dist_init_local": ([P(vp), i32], i32),
        "pbf_dist_set_replan": ([vp, u32], i32),
        "pbf_dist_upload": ([vp, vp, u64], i32),
        "pbf_dist_step": ([vp, P(Params)], i32),
        "pbf_dist_advance_host": ([vp, P(Params), vp, u64, P(u64)], i32),
        "pbf_dist_download": ([vp, vp, u64, P(u64)], i32),
        "pbf_dist_stats_read": ([vp, P(DistStats)], i32),
        "pbf_host_alloc": ([u64], vp),
        "pbf_host_free": ([vp], None),
        "pbf_host_grid": ([f32, P(Params), P(GridInfo)], i32),
        "pbf_host_plan_splits": ([vp, u32, u32, i32, vp], i32),
        "pbf_host_work_weights": ([vp, u32, u32, vp], i32),
        "pbf_host_constants": ([f32, vp], None),
        "pbf_host_morton_encode": ([u32, u32, u32], u32),
        "pbf_host_morton_decode": ([u32, vp], None),
        "pbf_host_apply_motion": ([P(Params), u64, P(Params)], None),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


class PbfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pbf_cuda error {code}: {msg}")
        self.code = code


def check(rc: int, ctx=None) -> None:
    if rc != PBF_OK:
        msg = lib().pbf_last_error(ctx)
        raise PbfError(rc, msg.decode() if msg else "")
