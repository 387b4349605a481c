"""CPU: the oracle's polygon mask / painting (the "clean" and "fill" buttons' use of Rasters.boolmask / rasterize!,
/root/reference/src/selectcolors.jl:480-481, 491-497, 559-563) against an independent pure-Python reading of the
stated rule (even-odd crossing test on the pixel points).  Parity unpinned: Rasters.jl is not in the tree."""
import numpy as np
import pytest


def py_in_ring(ring, pi, pj):
    ring = [(float(np.float32(a)), float(np.float32(b))) for a, b in ring]
    if len(ring) < 3:
        return False
    inside = False
    b1, b2 = ring[-1]
    for a1, a2 in ring:
        if (a2 > pj) != (b2 > pj):
            t = (pj - a2) / (b2 - a2)
            x = a1 + (b1 - a1) * t
            if pi < x:
                inside = not inside
        b1, b2 = a1, a2
    return inside


def py_polygons(rings, shape, fill=None, plane=None):
    n2, n1 = shape
    out = np.zeros(shape, dtype=np.uint8) if plane is None else plane.copy()
    for j in range(n2):
        for i in range(n1):
            hit = -1
            for p, r in enumerate(rings):
                if py_in_ring(r, float(i + 1), float(j + 1)):
                    hit = p
            if fill is None:
                out[j, i] = hit >= 0
            elif hit >= 0:
                out[j, i] = fill[hit]
    return out


def random_rings(rng, n1, n2, n):
    rings = []
    for _ in range(n):
        nv = int(rng.integers(1, 9))
        c = rng.random(2) * (n1, n2)
        rings.append([(float(c[0] + rng.normal(0, n1 / 4)), float(c[1] + rng.normal(0, n2 / 4))) for _ in range(nv)])
    return rings


@pytest.mark.parametrize("seed", range(6))
def test_polygons_against_python_reading(oracle, seed):
    rng = np.random.default_rng(seed)
    n1, n2 = int(rng.integers(1, 40)), int(rng.integers(1, 30))
    rings = random_rings(rng, n1, n2, int(rng.integers(0, 5)))
    rings.append([(100.0, 100.0)])                                    # the GUI's placeholder polygon: empty
    if seed == 0:
        rings.append([(2.0, 2.0), (8.0, 2.0), (8.0, 6.0), (2.0, 6.0)])    # vertices exactly on pixel points
    assert (oracle.polygons(rings, (n2, n1)) == py_polygons(rings, (n2, n1))).all()
    plane = rng.integers(0, 9, size=(n2, n1), dtype=np.uint8)
    fill = rng.integers(0, 9, size=len(rings)).astype(np.uint8)
    assert (oracle.polygons(rings, (n2, n1), fill, plane) == py_polygons(rings, (n2, n1), fill, plane)).all()


def test_axis_aligned_square_half_open(oracle):
    """the stated rule on a square whose edges pass through pixel points: the low edges belong to it, the high ones do not"""
    m = oracle.polygons([[(2, 2), (8, 2), (8, 6), (2, 6)]], (8, 10))
    want = np.zeros((8, 10), dtype=np.uint8)
    want[1:5, 1:7] = 1
    assert (m == want).all()
