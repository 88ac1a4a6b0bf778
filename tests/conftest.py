import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import ctypes as C
        from pbf_sph_b200 import capi
        ctx = C.c_void_p()
        rc = capi.lib().pbf_create(C.byref(ctx), C.c_float(0.1), 0)
        if rc == 0:
            capi.lib().pbf_destroy(ctx)
        return rc == 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not _has_gpu():
        pytest.fail("no usable CUDA device: the -m gpu tests exercise libpbf_cuda.so and have no CPU fallback")
    return True


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()
    return oracle
