"""pbf_sph_b200 — B200-native (sm_100a) Position-Based-Fluids step behind pbf-sph's solver interface.

The product is ``libpbf_cuda.so`` (csrc/: hand-written CUDA kernels + the C ABI of include/pbf_cuda.h); this
package is the thin host-side mirror of the reference's ``sph::Solver`` interface used by the tests and bench.
"""
from .capi import (FLAG_DEBUG_COUNTS, FLAG_GLOBAL_NEIGHBOURS, FLAG_PROFILE, FLAG_STRICT_FP, FLAG_PIN_HOST, FLAG_VORTICITY, FLAG_XSPH,
                   PARTICLE, Params, PbfError)
from .solver import Result, Solver
from . import capi, scenes

__all__ = ["Solver", "Result", "Params", "PARTICLE", "PbfError", "scenes", "capi", "FLAG_STRICT_FP",
           "FLAG_DEBUG_COUNTS", "FLAG_PROFILE", "FLAG_GLOBAL_NEIGHBOURS", "FLAG_XSPH", "FLAG_VORTICITY", "FLAG_PIN_HOST"]
