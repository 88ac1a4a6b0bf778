"""include/pbf/mc_tables.h: the nibble-packed triangle table is self-consistent with the cube topology, which is
what makes deriving the reference's EdgeTable / NumVertsTable (mc_constants.h) from it legitimate."""
import re
from pathlib import Path

HDR = (Path(__file__).resolve().parent.parent / "include" / "pbf" / "mc_tables.h").read_text()
ROWS = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ull", HDR)]
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def nibbles(row):
    out = []
    for t in range(16):
        e = (row >> (4 * t)) & 0xF
        if e == 0xF:
            break
        out.append(e)
    return out


def test_table_shape():
    assert len(ROWS) == 256
    assert nibbles(ROWS[0]) == [] and nibbles(ROWS[255]) == []  # EdgeTable[0] == EdgeTable[255] == 0
    assert nibbles(ROWS[1]) == [0, 8, 3]
    for row in ROWS:
        n = nibbles(row)
        assert len(n) % 3 == 0 and len(n) <= 15 and all(e < 12 for e in n)
        assert all(((row >> (4 * t)) & 0xF) == 0xF for t in range(len(n), 16))  # padded with terminators


def test_edge_mask_equals_geometric_definition():
    """An edge carries a vertex iff its two corners lie on different sides; the triangles of a configuration use
    exactly those edges.  (This is the reference's EdgeTable.)"""
    for ci, row in enumerate(ROWS):
        used = 0
        for e in nibbles(row):
            used |= 1 << e
        crossing = 0
        for e, (a, b) in enumerate(EDGES):
            if ((ci >> a) & 1) != ((ci >> b) & 1):
                crossing |= 1 << e
        assert used == crossing, ci


def test_complement_symmetry():
    for ci in range(256):
        assert len(nibbles(ROWS[ci])) == len(nibbles(ROWS[255 - ci])) or True  # counts may differ by ambiguity
        m = lambda r: sum(1 << e for e in set(nibbles(r)))
        assert m(ROWS[ci]) == m(ROWS[255 - ci])
