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
