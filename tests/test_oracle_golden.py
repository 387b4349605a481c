"""CPU: pin the oracle against the reference's own known-answer vectors
(/root/reference/test/runtests.jl:6-20, committed as tests/golden/runtests_vectors.json) and
against hand-derived cases of the restated arithmetic."""
import numpy as np
import pytest

from util import golden_palette


def test_reference_known_answer_vectors(oracle):
    mean, mn, mx, sd, tol, vectors = golden_palette()
    lo, hi = oracle.bounds(mn, mx, sd, [tol, tol])
    got = [oracle.categorize_f64(v["rgb"], mean, lo, hi) for v in vectors]
    assert got == [v["expected"] for v in vectors] == [0, 1, 1, 2]


def test_first_vector_pins_float64_bound_arithmetic(oracle):
    # 0.8 <= 0.7 + 0.1*1.0 is false in Float64 (0.7999999999999999): SURVEY Appendix B
    mean, mn, mx, sd, tol, vectors = golden_palette()
    lo, hi = oracle.bounds(mn, mx, sd, [tol, tol])
    assert hi[0, 2] == 0.7 + 0.1 * 1.0 < 0.8
    # with today's default tolerance 2.0 (src/selectcolors.jl:1) the first vector becomes 1
    lo2, hi2 = oracle.bounds(mn, mx, sd, [2.0, 2.0])
    assert oracle.categorize_f64(vectors[0]["rgb"], mean, lo2, hi2) == 1


def test_bounds_are_two_rounded_ops(oracle):
    rng = np.random.default_rng(1)
    mn, mx, sd = rng.random((50, 3)), rng.random((50, 3)), rng.random((50, 3))
    tol = rng.random(50) * 20 - 2
    lo, hi = oracle.bounds(mn, mx, sd, tol)
    assert np.array_equal(lo, mn - sd * tol[:, None])
    assert np.array_equal(hi, mx + sd * tol[:, None])


def test_n0f8_is_correctly_rounded(oracle):
    for i in range(256):
        assert oracle.lib().mro_n0f8(i) == i / 255.0


def test_first_minimum_wins_ties(oracle):
    # two identical categories: findmin returns the first (categories.jl:80)
    mean = np.array([[0.5, 0.5, 0.5], [0.5, 0.5, 0.5]])
    lo, hi = mean - 0.1, mean + 0.1
    assert oracle.categorize_f64([0.5, 0.5, 0.5], mean, lo, hi) == 1
    # exact symmetric tie between different means
    mean = np.array([[0.25, 0.5, 0.5], [0.75, 0.5, 0.5]])
    lo, hi = mean - 0.5, mean + 0.5
    assert oracle.categorize_f64([0.5, 0.5, 0.5], mean, lo, hi) == 1


def test_box_applies_to_winner_only(oracle):
    # winner rejects -> 0 even if the runner-up's box would accept (categories.jl:81)
    mean = np.array([[0.2, 0.2, 0.2], [0.6, 0.6, 0.6]])
    lo = np.array([[0.19, 0.19, 0.19], [0.0, 0.0, 0.0]])
    hi = np.array([[0.21, 0.21, 0.21], [1.0, 1.0, 1.0]])
    assert oracle.categorize_f64([0.3, 0.3, 0.3], mean, lo, hi) == 0
    assert oracle.categorize_f64([0.5, 0.5, 0.5], mean, lo, hi) == 2


def test_nan_bounds_never_accept_and_missing_category(oracle):
    mean = np.array([[0.2, 0.2, 0.2], [np.inf, np.inf, np.inf]])
    lo = np.array([[np.nan, 0, 0], [0, 0, 0.0]])
    hi = np.array([[1, 1, 1], [1, 1, 1.0]])
    assert oracle.categorize_f64([0.2, 0.2, 0.2], mean, lo, hi) == 0      # one-point category: sd = NaN
    lo[0, 0] = 0.0
    assert oracle.categorize_f64([0.9, 0.9, 0.9], mean, lo, hi) == 1      # missing category never wins
    st = oracle.component_stats([0.5])
    assert np.isnan(st[3]) and st[0] == 0.5


def test_known_category_override(oracle):
    mean = np.array([[0.2, 0.2, 0.2], [0.6, 0.6, 0.6]])
    lo, hi = mean - 0.01, mean + 0.01
    x = [0.3, 0.3, 0.3]                       # best = 1, outside its box
    assert oracle.categorize_f64(x, mean, lo, hi, 0) == 0
    assert oracle.categorize_f64(x, mean, lo, hi, 1) == 1   # known == best: accepted without the box
    assert oracle.categorize_f64(x, mean, lo, hi, 2) == 0   # known != best


def test_rgba_alpha_term_and_transparent_rule(oracle):
    mean = np.array([[0.5, 0.5, 0.5, 1.0], [0.5, 0.5, 0.5, 0.0]])
    lo, hi = mean - 0.6, mean + 0.6
    assert oracle.categorize_f64([0.5, 0.5, 0.5, 0.9], mean, lo, hi) == 1
    assert oracle.categorize_f64([0.5, 0.5, 0.5, 0.1], mean, lo, hi) == 2
    assert oracle.categorize_f64([0.5, 0.5, 0.5, 0.0], mean, lo, hi) == 0   # categories.jl:116


def test_majority_hand_cases(oracle):
    L = np.array([[1, 1, 2],
                  [1, 2, 2],
                  [3, 3, 3]], dtype=np.uint8)
    out = oracle.majority(L, 1)
    # centre window: 1x3, 2x3, 3x3 -> tie -> lowest label 1
    assert out[1, 1] == 1
    # corner (0,0): window {1,1,1,2} -> 1 ; corner (2,2): {2,2,3,3} -> tie -> 2
    assert out[0, 0] == 1 and out[2, 2] == 2
    # label 0 votes like any other label
    Z = np.zeros((5, 5), dtype=np.uint8)
    Z[2, 2] = 7
    assert oracle.majority(Z, 1)[2, 2] == 0
    # R = 0 is the identity; centre-if-modal switch keeps the centre on ties
    assert np.array_equal(oracle.majority(L, 0), L)
    assert oracle.majority(L, 1, oracle.TIE_CENTRE_IF_MODAL)[1, 1] == 2


@pytest.mark.parametrize("R", [1, 2, 3])
def test_majority_bands_with_halo_equal_whole(oracle, R):
    rng = np.random.default_rng(R)
    L = rng.integers(0, 5, size=(40, 33), dtype=np.uint8)
    whole = oracle.majority(L, R)
    for a, b in [(0, 13), (13, 29), (29, 40)]:
        hl, hh = min(R, a), min(R, 40 - b)
        got = oracle.majority(L[a - hl:b + hh], R, halo_lo=hl, halo_hi=hh)
        assert np.array_equal(got, whole[a:b])


def test_image_level_matches_per_pixel(oracle):
    rng = np.random.default_rng(3)
    mean, mn, mx, sd, tol = __import__("util").random_palette(rng, 6)
    lo, hi = oracle.bounds(mn, mx, sd, tol)
    px = rng.integers(0, 256, size=(7, 9, 3), dtype=np.uint8)
    lab = oracle.categorize_u8(px, mean, lo, hi)
    for j in range(7):
        for i in range(9):
            assert lab[j, i] == oracle.categorize_f64(px[j, i] / 255.0, mean, lo, hi)
    lut = oracle.build_lut(mean, lo, hi)
    idx = px[..., 0].astype(np.int64) | (px[..., 1].astype(np.int64) << 8) | (px[..., 2].astype(np.int64) << 16)
    assert np.array_equal(lut[idx], lab)


def test_lab_conversion_sanity(oracle):
    assert np.allclose(oracle.rgb8_to_lab(255, 255, 255), [100.0, 0.0, 0.0], atol=2e-3)
    assert np.allclose(oracle.rgb8_to_lab(0, 0, 0), [0.0, 0.0, 0.0], atol=1e-12)
    assert np.allclose(oracle.rgb8_to_lab(255, 0, 0), [53.2408, 80.0925, 67.2032], atol=1e-3)
    for t in (1e-3, 0.0088565, 0.2, 1.0, 1.0889):
        assert abs(oracle.lib().mro_cbrt(t) - t ** (1 / 3)) < 1e-15 * 4
