"""CPU: the oracle's blur (restating /root/reference/src/blur.jl:6-14) against an independent Python reading
of the same lines (sequential sum in CartesianIndices order, times 1 / count, zero colour outside) and the
oracle's image-level Float64 categoriser against the reference's own known-answer vectors
(test/runtests.jl:6-20, tests/golden/)."""
import numpy as np
import pytest

from util import golden_palette


def py_blur(img, R):
    n2, n1, _ = img.shape
    x = img.astype(np.float64) / np.float64(255.0) if img.dtype == np.uint8 else img
    out = np.zeros((n2, n1, 3))
    rc = 1.0 / float((2 * R + 1) ** 2)
    for j in range(n2):
        for i in range(n1):
            s = [0.0, 0.0, 0.0]
            for d2 in range(-R, R + 1):
                for d1 in range(-R, R + 1):
                    a, b = i + d1, j + d2
                    if 0 <= a < n1 and 0 <= b < n2:
                        for c in range(3):
                            s[c] = s[c] + float(x[b, a, c])
            for c in range(3):
                out[j, i, c] = rc * s[c]
    return out


@pytest.mark.parametrize("R", [0, 1, 2])
@pytest.mark.parametrize("shape", [(1, 1), (3, 7), (12, 9)])
def test_blur_against_python_reading(oracle, R, shape):
    rng = np.random.default_rng(R * 10 + shape[0])
    img = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    got = oracle.blur(img, R, 1)
    want = py_blur(img, R)
    assert got.tobytes() == want.tobytes()
    got2 = oracle.blur(img, R, 2)
    assert got2.tobytes() == py_blur(want, R).tobytes()
    assert oracle.blur(img, R, 0).tobytes() == (img.astype(np.float64) / 255.0).tobytes()


def test_float64_image_categoriser_on_the_reference_vectors(oracle):
    mean, mn, mx, sd, tol, vectors = golden_palette()
    lo, hi = oracle.bounds(mn, mx, sd, np.full(mean.shape[0], tol))
    px = np.array([[v["rgb"] for v in vectors]], dtype=np.float64)           # one line of the reference's colours
    want = np.array([[v["expected"] for v in vectors]], dtype=np.uint8)
    assert (oracle.categorize_f64_image(px, mean, lo, hi) == want).all()


def test_balance_against_python_reading(oracle):
    """oracle `_balance` per-pixel part against a direct Python reading of src/balance.jl:25-42 with the stated
    evaluation and summation order"""
    rng = np.random.default_rng(11)
    n2, n1 = 5, 300          # two chunks of 256 pixels per line
    img = rng.integers(0, 256, size=(n2, n1, 3), dtype=np.uint8)
    coef = rng.normal(0, 1e-6, size=(3, 7))
    coef[:, 0] = rng.normal(0, 0.05, size=3)

    def pred(b, i, j):
        fi, fj = float(i), float(j)
        i2 = fi * fi
        i3 = i2 * fi
        j2 = fj * fj
        j3 = j2 * fj
        p = b[0] + b[1] * i3
        for t in (b[2] * i2, b[3] * fi, b[4] * j3, b[5] * j2, b[6] * fj):
            p = p + t
        return p
    want = np.zeros((n2, n1, 3))
    for c in range(3):
        b = [float(v) for v in coef[c]]
        total = 0.0
        for j in range(n2):
            line = 0.0
            for i0 in range(0, n1, 256):
                chunk = 0.0
                for i in range(i0, min(n1, i0 + 256)):
                    chunk = chunk + pred(b, i + 1, j + 1)
                line = line + chunk
            total = total + line
        mean = total / float(n1 * n2)
        for j in range(n2):
            for i in range(n1):
                v = float(np.float64(img[j, i, c]) / np.float64(255.0)) - pred(b, i + 1, j + 1)
                v = v + mean
                want[j, i, c] = min(1.0, max(v, 0.0))
    assert oracle.balance(img, coef).tobytes() == want.tobytes()
