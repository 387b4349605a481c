"""CPU: the segment-prune oracle (`_clean`, /root/reference/src/clean.jl:59-93 over
fast_scanning!, src/scan.jl:56-157) against an independent formulation.

Without conflicting known categories the sequential union-find segmentation is exactly
4-connected component labelling of equal labels, so scipy.ndimage.label per category plus
the per-component rule of clean.jl:65-88 must give the same image.  Hand-worked cases pin
the rule itself (sizes, the 1 x 2 rectangle convention of clean.jl:68-70, known categories)."""
import numpy as np
import pytest
from scipy import ndimage

from util import mismatches


def independent(labels, known, categories, line_threshold, prune_threshold):
    out = np.zeros_like(labels)
    four = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    for c in np.unique(labels):
        comp, n = ndimage.label(labels == c, structure=four)
        if n == 0:
            continue
        idx = np.arange(1, n + 1)
        count = ndimage.sum_labels(np.ones_like(comp), comp, idx)
        sl = ndimage.find_objects(comp)
        for k in range(n):
            m = comp[sl[k]] == (k + 1)
            h, w = m.shape
            ratio = count[k] / (max(h, w) ** 2 / 2)
            cats = np.unique(known[sl[k]][m]) if known is not None else np.array([0])
            cats = cats[cats != 0]
            assert len(cats) <= 1, "test images must not put two known categories in one segment"
            cat = int(cats[0]) if len(cats) else 0
            new = int(c)
            if c not in categories:
                new = 0
            elif count[k] < prune_threshold or ratio < line_threshold:
                if cat == 0:
                    new = 0
                elif c not in (0, cat):
                    new = cat
            out[sl[k]][m] = new
    return out


def blobs(rng, n2, n1, K, speckle):
    """smooth label field + speckle + a few thin lines"""
    f = rng.integers(1, K + 1, size=(n2 // 16 + 2, n1 // 16 + 2)).astype(np.uint8)
    lab = np.kron(f, np.ones((16, 16), dtype=np.uint8))[:n2, :n1].copy()
    sp = rng.random((n2, n1)) < speckle
    lab[sp] = rng.integers(0, K + 1, size=int(sp.sum()))
    for _ in range(4):
        y = int(rng.integers(0, n2))
        lab[y, : n1 // 2] = int(rng.integers(1, K + 1))
        x = int(rng.integers(0, n1))
        lab[: n2 // 3, x] = int(rng.integers(1, K + 1))
    return lab


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("shape", [(64, 80), (97, 131), (1, 50), (40, 1)])
def test_oracle_vs_component_labelling(oracle, seed, shape):
    rng = np.random.default_rng(seed)
    n2, n1 = shape
    K = 6
    lab = blobs(rng, n2, n1, K, 0.04)
    for cats, lt, pt in [((1, 2, 3, 4, 5, 6), 0.01, 20), ((1, 3, 5), 0.2, 5), ((2,), 0.5, 400), ((1, 2, 3, 4, 5, 6), 0.0, 0)]:
        got, nseg = oracle.clean_segments(lab, None, cats, lt, pt)
        want = independent(lab, None, cats, lt, pt)
        assert mismatches(got, want) == 0
        assert nseg >= 1


def test_known_categories(oracle):
    rng = np.random.default_rng(7)
    lab = blobs(rng, 96, 96, 5, 0.05)
    known = np.zeros_like(lab)
    # one known point per small segment at most: put them on isolated speckle pixels
    four = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    placed = 0
    for c in range(1, 6):
        comp, n = ndimage.label(lab == c, structure=four)
        sizes = ndimage.sum_labels(np.ones_like(comp), comp, np.arange(1, n + 1))
        for k in np.nonzero(sizes < 6)[0][:10]:
            ys, xs = np.nonzero(comp == k + 1)
            known[ys[0], xs[0]] = 1 + (c + placed) % 5
            placed += 1
    assert placed > 10
    cats = (1, 2, 3, 4, 5)
    got, _ = oracle.clean_segments(lab, known, cats, 0.01, 20)
    want = independent(lab, known, cats, 0.01, 20)
    assert mismatches(got, want) == 0
    assert (got != np.where(np.isin(lab, cats), lab, 0)).any()


def test_hand_worked(oracle):
    # a 1 x 2 rectangle has ratio exactly 1 (clean.jl:68-70): 2 / (2^2 / 2)
    lab = np.zeros((5, 8), dtype=np.uint8)
    lab[1, 1:3] = 1            # 2 px, bbox 1 x 2 -> ratio 1.0
    lab[3, 0:8] = 2            # 8 px line, bbox 1 x 8 -> ratio 8 / 32 = 0.25
    out, n = oracle.clean_segments(lab, None, (1, 2), line_threshold=0.25, prune_threshold=2)
    assert out[1, 1] == 1 and out[3, 0] == 2          # 0.25 < 0.25 is false; 2 < 2 is false
    out, _ = oracle.clean_segments(lab, None, (1, 2), line_threshold=0.26, prune_threshold=2)
    assert out[1, 1] == 1 and (out[3] == 0).all()     # the line goes
    out, _ = oracle.clean_segments(lab, None, (1, 2), line_threshold=0.0, prune_threshold=3)
    assert (out[1] == 0).all() and out[3, 0] == 2     # the small segment goes
    out, _ = oracle.clean_segments(lab, None, (2,), line_threshold=0.0, prune_threshold=0)
    assert (out[1] == 0).all() and out[3, 0] == 2     # category 1 is not kept
    known = np.zeros_like(lab)
    known[1, 2] = 2                                   # clicked point of category 2 inside the small segment
    out, _ = oracle.clean_segments(lab, known, (1, 2), line_threshold=0.0, prune_threshold=3)
    assert out[1, 1] == 2 and out[1, 2] == 2          # forced into the known category (clean.jl:79-85)
    # diagonal neighbours are NOT connected (backward neighbours x - e1, x - e2 only)
    lab = np.zeros((4, 4), dtype=np.uint8)
    lab[0, 0] = lab[1, 1] = lab[2, 2] = 3
    _, n = oracle.clean_segments(lab, None, (3,), 0.0, 0)
    assert n == 3 + 1                                 # three singletons + one background segment... see below


def test_scan_order_conflict(oracle):
    # two different known categories in one run of equal labels: the reference starts a new
    # segment at the pixel where the merge is refused (scan.jl:140-146) -- order dependent
    lab = np.full((1, 6), 4, dtype=np.uint8)
    known = np.zeros_like(lab)
    known[0, 1] = 1
    known[0, 4] = 2
    out, n = oracle.clean_segments(lab, known, (1, 2, 4), line_threshold=0.0, prune_threshold=100)
    assert n == 2
    assert list(out[0]) == [1, 1, 1, 1, 2, 2]
