"""CPU: the oracle's texture features (restating /root/reference/src/selectcolors.jl:706-715 and
src/stripes.jl:1-35) and its PixelProp categoriser (src/categories.jl:76-133 with a match tuple) against an
independent pure-Python reading of the same lines with the decisions stated in the oracle (zero colour outside,
left-folded sums, two-pass std, classical RGB -> HSL).  Parity unpinned by reference fixtures (none exist)."""
import math

import numpy as np
import pytest


def hsl_sl(c):
    r, g, b = (float(v) for v in c)
    mx, mn = max(r, g, b), min(r, g, b)
    l = (mx + mn) / 2.0
    if mx == mn:
        return 0.0, l
    if l < 0.5:
        return (mx - mn) / (mx + mn), l
    return (mx - mn) / ((2.0 - mx) - mn), l


def at(px, a, b):
    n2, n1, _ = px.shape
    return px[b, a] if 0 <= a < n1 and 0 <= b < n2 else np.zeros(3)


def py_stds(px):
    n2, n1, _ = px.shape
    raw = np.zeros((n2, n1))
    for j in range(n2):
        for i in range(n1):
            sds = []
            for c in range(3):
                vals = [float(at(px, i + d1, j + d2)[c]) for d2 in range(-2, 3) for d1 in range(-2, 3)]
                s = 0.0
                for v in vals:
                    s = s + v
                m = s / 25.0
                ss = 0.0
                for v in vals:
                    ss = ss + (v - m) * (v - m)
                sds.append(math.sqrt(ss / 24.0))
            raw[j, i] = (sds[0] + sds[1]) + sds[2]
    return raw / raw.max()


def fixed_order_mean(plane):
    n2, n1 = plane.shape
    total = 0.0
    for j in range(n2):
        line = 0.0
        for i0 in range(0, n1, 256):
            chunk = 0.0
            for i in range(i0, min(n1, i0 + 256)):
                chunk = chunk + float(plane[j, i])
            line = line + chunk
        total = total + line
    return total / float(n1 * n2)


def py_stripes(px, radius):
    n2, n1, _ = px.shape
    raw = np.zeros((n2, n1, 2))
    dirs = [(1, 0), (0, 1), (1, 1), (1, -1)]
    for j in range(n2):
        for i in range(n1):
            s0, l0 = hsl_sl(px[j, i])
            val = []
            for di, dj in dirs:
                sides = []
                for xs in (range(-radius, 0), range(1, radius + 1)):
                    acc = 0.0
                    for x in xs:
                        sn, ln = hsl_sl(at(px, i + di * x, j + dj * x))
                        acc = acc + ((s0 - sn) * (s0 - sn) + (l0 - ln) * (l0 - ln))
                    sides.append(acc / float(radius))
                val.append(min(sides))
            raw[j, i] = (val[0] - val[1], val[2] - val[3])
    m = (fixed_order_mean(np.abs(raw[..., 0])) + fixed_order_mean(np.abs(raw[..., 1]))) / 2.0
    return raw / m


@pytest.mark.parametrize("shape", [(1, 1), (4, 9), (7, 300)])
def test_stds_against_python_reading(oracle, shape):
    rng = np.random.default_rng(shape[1])
    px = rng.random(shape + (3,))
    with np.errstate(invalid="ignore"):
        want = py_stds(px)
        got = oracle.stds(px)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("radius", [1, 2, 3])
@pytest.mark.parametrize("shape", [(5, 8), (3, 270)])
def test_stripes_against_python_reading(oracle, radius, shape):
    rng = np.random.default_rng(radius * 100 + shape[1])
    px = rng.random(shape + (3,))
    px[0, 0] = 0.25                      # a grey pixel: s = 0 branch
    px[1, 2] = (0.9, 0.95, 0.99)         # l >= 0.5 branch
    assert oracle.stripes(px, radius).tobytes() == py_stripes(px, radius).tobytes()


def py_categorize_props(rgb, sd, st, match, mean, lo, hi, sds, sts, known=0):
    K = mean.shape[0]
    errs = []
    for c in range(K):
        if mean[c, 0] == np.inf:
            errs.append(np.inf)
            continue
        terms = []
        if "color" in match:
            d = [float(rgb[k]) - float(mean[c, k]) for k in range(3)]
            terms.append((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
        if "std" in match:
            d = float(sd) - float(sds[0][c])
            terms.append(d * d)
        if "stripe" in match:
            d = [float(st[k]) - float(sts[0][c, k]) for k in range(2)]
            terms.append((d[0] * d[0] + d[1] * d[1]) / 2.0)
        t = terms[0]
        for x in terms[1:]:
            t = t + x
        errs.append(t / float(len(terms)))
    best = int(np.argmin(errs))           # first minimum
    if known:
        return best + 1 if best + 1 == known else 0
    ok = True
    if "color" in match:
        ok = ok and all(lo[best, k] <= rgb[k] <= hi[best, k] for k in range(3))
    if "std" in match:
        ok = ok and sds[1][best] <= sd <= sds[2][best]
    if "stripe" in match:
        ok = ok and all(sts[1][best, k] <= st[k] <= sts[2][best, k] for k in range(2))
    return best + 1 if ok else 0


@pytest.mark.parametrize("match", [("color",), ("color", "std"), ("color", "std", "stripe"), ("std", "stripe"), ("stripe",)])
def test_props_categoriser_against_python_reading(oracle, match):
    rng = np.random.default_rng(len(match) * 7 + len(match[0]))
    K, n2, n1 = 5, 6, 11
    mean = rng.random((K, 3))
    mean[3] = np.inf                     # a `missing` category
    lo, hi = mean - 0.25, mean + 0.25
    sdm = rng.random(K)
    stm = rng.normal(0, 1, (K, 2))
    sds = (sdm, sdm - 0.4, sdm + 0.4)
    sts = (stm, stm - 1.0, stm + 1.0)
    px = rng.random((n2, n1, 3))
    sd = rng.random((n2, n1))
    st = rng.normal(0, 1, (n2, n1, 2))
    kidx = np.array([3, 17, 40], dtype=np.int64)
    kcat = np.array([1, 2, 5], dtype=np.uint8)
    got = oracle.categorize_props_image(px, sd, st, match, mean, lo, hi, sds, sts, kidx, kcat)
    want = np.zeros((n2, n1), dtype=np.uint8)
    known = dict(zip(kidx.tolist(), kcat.tolist()))
    for j in range(n2):
        for i in range(n1):
            want[j, i] = py_categorize_props(px[j, i], sd[j, i], st[j, i], match, mean, lo, hi, sds, sts,
                                             known.get(i + j * n1, 0))
    assert (got == want).all()
    assert len(np.unique(got)) > 2


def test_color_only_match_equals_the_colour_categoriser(oracle):
    rng = np.random.default_rng(3)
    K = 4
    mean = rng.random((K, 3))
    lo, hi = mean - 0.3, mean + 0.3
    px = rng.random((9, 13, 3))
    z1, z2 = np.zeros(K), np.zeros((K, 2))
    a = oracle.categorize_props_image(px, None, None, ("color",), mean, lo, hi, (z1, z1, z1), (z2, z2, z2))
    b = oracle.categorize_f64_image(px, mean, lo, hi)
    assert (a == b).all()
