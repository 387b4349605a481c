"""CPU: the oracle's fill_gaps (oracle/maprast_oracle.c, restating /root/reference/src/clean.jl:2-54) against
hand-worked cases and an independent pure-Python reading of the same source lines.  Parity unpinned by the
reference (no fixture for fill_gaps exists there); these tests pin the oracle to the decisions in DESIGN.md."""
import math

import numpy as np
import pytest

from util import mismatches

OFFS = [(d1, d2) for d2 in range(-2, 3) for d1 in range(-2, 3) if 1 <= abs(d1) + abs(d2) <= 2]   # CartesianIndices order


def py_fill(lab, px, mean, cats, known=None, mask=None):
    """one pass, Python floats (IEEE binary64, same operation order as clean.jl:43-54)"""
    n2, n1 = lab.shape
    out = np.zeros_like(lab)
    cats = list(cats)
    for j in range(n2):
        for i in range(n1):
            v = int(lab[j, i])
            if mask is not None and mask[j, i]:
                out[j, i] = v
                continue
            kn = int(known[j, i]) if known is not None else 0
            if kn != 0 and kn in cats:
                out[j, i] = kn
                continue
            if v != 0:
                out[j, i] = v if v in cats else 0
                continue
            hood = []
            for d1, d2 in OFFS:
                a, b = i + d1, j + d2
                hood.append(int(lab[b, a]) if 0 <= a < n1 and 0 <= b < n2 else 0)
            if sum(h == 0 for h in hood) > 12 - 12 // 4:
                continue
            counts = []
            for c in cats:
                acc = 0.0
                for (d1, d2), h in zip(OFFS, hood):
                    if h != 0 and h == c:
                        acc += 1 / math.sqrt(d1 * d1 + d2 * d2)
                counts.append(acc)
            if not counts:
                continue
            im = max(range(len(cats)), key=lambda k: (counts[k], -k))
            if not counts[im] > 1:
                continue
            sec, isx = 0, 0
            for k, x in enumerate(counts):
                if k == im:
                    continue
                if x > sec:
                    sec, isx = x, k
            pick = im
            if sec > 1:
                x = [float(np.float64(t) / np.float64(255.0)) for t in px[j, i]]

                def err(k):
                    m = mean[cats[k] - 1]
                    e = (x[0] - m[0]) ** 2 + (x[1] - m[1]) ** 2
                    e = e + (x[2] - m[2]) ** 2
                    return e
                pick = im if err(im) <= err(isx) else isx
            out[j, i] = cats[pick]
    return out


def holes(rng, n2, n1, K, frac):
    f = rng.integers(1, K + 1, size=(n2 // 8 + 2, n1 // 8 + 2)).astype(np.uint8)
    lab = np.kron(f, np.ones((8, 8), dtype=np.uint8))[:n2, :n1].copy()
    lab[rng.random((n2, n1)) < frac] = 0
    for _ in range(3):   # bigger gaps
        y, x = int(rng.integers(0, n2)), int(rng.integers(0, n1))
        lab[y:y + 6, x:x + 9] = 0
    return lab


def test_hand_cases(oracle):
    mean = np.array([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]])
    px = np.zeros((5, 5, 3), np.uint8)
    lab = np.zeros((5, 5), np.uint8)
    lab[2, 1] = 1
    lab[2, 3] = 1
    lab[1, 2] = 2
    out = oracle.fill_gaps(lab, px, mean, [1, 2])
    assert out[2, 2] == 1            # weights 2.0 (category 1) against 1.0 (category 2, not > 1)
    assert out[3, 2] == 1            # two diagonal neighbours: 2/sqrt(2) = 1.414 > 1, ten of twelve... nine missing
    assert out[0, 0] == 0            # mostly missing
    # tie between two categories with weight 2 each: the original colour decides
    lab = np.zeros((5, 5), np.uint8)
    lab[2, 1] = lab[2, 3] = 1
    lab[1, 2] = lab[3, 2] = 2
    px[2, 2] = (250, 250, 250)
    assert oracle.fill_gaps(lab, px, mean, [1, 2])[2, 2] == 2
    px[2, 2] = (3, 3, 3)
    assert oracle.fill_gaps(lab, px, mean, [1, 2])[2, 2] == 1
    assert oracle.fill_gaps(lab, px, mean, [2, 1])[2, 2] == 1       # colour, not list order
    px[2, 2] = (128, 128, 128)       # 128/255 is farther from 0 than from 1
    assert oracle.fill_gaps(lab, px, mean, [1, 2])[2, 2] == 2
    # labels that are not kept become 0 and are not counted; known and mask rules
    lab = np.full((4, 6), 3, np.uint8)
    out = oracle.fill_gaps(lab, np.zeros((4, 6, 3), np.uint8), np.zeros((3, 3)), [1, 2])
    assert (out == 0).all()
    known = np.zeros((4, 6), np.uint8)
    known[1, 1] = 2
    known[2, 2] = 3
    mask = np.zeros((4, 6), np.uint8)
    mask[0, 0] = 1
    out = oracle.fill_gaps(lab, np.zeros((4, 6, 3), np.uint8), np.zeros((3, 3)), [1, 2], known=known, mask=mask)
    assert out[1, 1] == 2 and out[2, 2] == 0 and out[0, 0] == 3


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (7, 1), (24, 37), (40, 33)])
def test_against_python_reading(oracle, seed, shape):
    rng = np.random.default_rng(500 + seed)
    n2, n1 = shape
    K = 5
    lab = holes(rng, n2, n1, K, 0.25)
    px = rng.integers(0, 256, size=(n2, n1, 3), dtype=np.uint8)
    mean = rng.random((K, 3))
    known = np.where(rng.random((n2, n1)) < 0.03, rng.integers(1, K + 1, size=(n2, n1)), 0).astype(np.uint8)
    mask = (rng.random((n2, n1)) < 0.05).astype(np.uint8)
    for cats in ([1, 2, 3, 4, 5], [4, 2, 5], [3], []):
        want = py_fill(lab, px, mean, cats, known, mask)
        got = oracle.fill_gaps(lab, px, mean, cats, known=known, mask=mask)
        assert mismatches(got, want) == 0
        want2 = py_fill(py_fill(lab, px, mean, cats), px, mean, cats)
        assert mismatches(oracle.fill_gaps(lab, px, mean, cats, repeat=2), want2) == 0
