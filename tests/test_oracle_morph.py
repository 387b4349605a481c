"""CPU: the oracle's categorical erosion / dilation (`_erode`, `_dilate`, `_shrink_grow`,
/root/reference/src/clean.jl:96-140) against an independent pure-Python reading of the same lines with the stated
decisions (Moore(1) = 8 neighbours in CartesianIndices order, 0 outside).  Parity unpinned."""
import numpy as np
import pytest

OFFS = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]     # (d1, d2), d1 fastest


def hood(A, i, j):
    n2, n1 = A.shape
    return [int(A[j + d2, i + d1]) if 0 <= i + d1 < n1 and 0 <= j + d2 < n2 else 0 for d1, d2 in OFFS]


def py_erode(A):
    out = A.copy()
    for j in range(A.shape[0]):
        for i in range(A.shape[1]):
            x = int(A[j, i])
            out[j, i] = 0 if all(v == x for v in hood(A, i, j)) else x
    return out


def py_dilate(A):
    out = A.copy()
    for j in range(A.shape[0]):
        for i in range(A.shape[1]):
            if A[j, i] != 0:
                continue
            h = hood(A, i, j)
            counts = [sum(1 for q in h if q == g) for g in h]
            v = h[counts.index(max(counts))]                      # findmax: first maximum
            if v == 0:
                v = next((q for q in h if q != 0), 0)              # findfirst(!=(missingval), hood)
            out[j, i] = v
    return out


def random_labels(rng, n2, n1, K, holes):
    lab = rng.integers(1, K + 1, size=(n2 // 4 + 1, n1 // 4 + 1)).repeat(4, axis=0).repeat(4, axis=1)[:n2, :n1]
    lab = lab.astype(np.uint8)
    lab[rng.random((n2, n1)) < holes] = 0
    return np.ascontiguousarray(lab)


@pytest.mark.parametrize("seed", range(6))
def test_erode_dilate_against_python_reading(oracle, seed):
    rng = np.random.default_rng(seed)
    n2, n1 = int(rng.integers(1, 25)), int(rng.integers(1, 30))
    A = random_labels(rng, n2, n1, 5, float(rng.choice([0.0, 0.1, 0.6])))
    assert (oracle.shrink_grow(A, 1, 0) == py_erode(A)).all()
    assert (oracle.shrink_grow(A, 0, 1) == py_dilate(A)).all()
    assert (oracle.shrink_grow(A, 2, 3) == py_dilate(py_dilate(py_dilate(py_erode(py_erode(A)))))).all()
    assert (oracle.shrink_grow(A, 0, 0) == A).all()
