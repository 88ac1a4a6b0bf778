"""The two device primitives every compaction and the Z-order sort go through (csrc/sort_scan.cu), against numpy on the
sizes where their launch shapes change: the exclusive scan (one tile / one wide block up to 128 K entries / three-level
above) and the stable LSD radix sort of (key, index) pairs (the std::sort of ompsph.hpp:158 made stable)."""
import numpy as np
import pytest

from pbf_sph_b200 import Solver, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 31, 2047, 2048, 2049, 4096, 16384, 16385, 50000, 131071, 131072, 131073, 1 << 20, 5_000_003])
def test_exclusive_scan_matches_numpy(gpu, n):
    rng = np.random.default_rng(n)
    v = rng.integers(0, 300, n, dtype=np.uint32)
    v[rng.integers(0, n, max(1, n // 7))] = 0  # runs of empty tiles, as the slab path's counters have
    with Solver(scenes.H, 0) as s:
        got, total = s.debug_scan(v)
    want = np.concatenate([[0], np.cumsum(v, dtype=np.uint64)[:-1]]).astype(np.uint32)
    assert np.array_equal(got, want)
    assert total == int(v.sum(dtype=np.uint64))


@pytest.mark.parametrize("n", [1, 5, 4095, 4096, 4097, 100_000, 1_000_000])
def test_radix_sort_is_the_stable_sort(gpu, n):
    rng = np.random.default_rng(n + 1)
    keys = rng.integers(0, 488_063, n, dtype=np.uint32)          # dam-1m's cell-table size: many duplicates per key
    far = rng.integers(0, n, max(1, n // 500))
    keys[far] |= rng.integers(0, 1 << 10, len(far), dtype=np.uint32) << 20  # particles predicted outside the grid: 30-bit keys
    with Solver(scenes.H, 0) as s:
        ks, perm = s.debug_sort_pairs(keys)
    want = np.argsort(keys, kind="stable").astype(np.uint32)
    assert np.array_equal(perm, want)
    assert np.array_equal(ks, keys[want])
