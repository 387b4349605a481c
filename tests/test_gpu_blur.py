"""GPU (-m gpu): blur (/root/reference/src/blur.jl:6-14) and the categoriser on RGB{Float64} pixels through the
C ABI.  blur: bit-identical Float64 images against the oracle.  Categoriser: the reference's own known-answer
vectors (test/runtests.jl:6-20 are Float64 colours) go through the PRODUCT's Float64 path here, plus random
images against the oracle."""
import numpy as np
import pytest
import torch

from util import golden_palette, mismatches, random_palette

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(1, 1), (1, 40), (33, 1), (37, 53), (128, 200)])
def test_blur_bit_identical(ctx, oracle, R, shape):
    rng = np.random.default_rng(31 * R + shape[1])
    img = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
    for rep in (0, 1, 2, 3):
        got = ctx.blur_host(img, R, rep)
        want = oracle.blur(img, R, rep)
        assert got.tobytes() == want.tobytes(), (R, rep)
    f = rng.random(shape + (3,))
    assert ctx.blur_host(f, R, 2).tobytes() == oracle.blur(f, R, 2).tobytes()


def test_blur_device_entry_with_pitches(ctx, oracle):
    rng = np.random.default_rng(5)
    n2, n1, R = 70, 91, 2
    img = rng.integers(0, 256, size=(n2, n1, 3), dtype=np.uint8)
    dev = torch.device("cuda", 0)
    ip, op = 320, (n1 * 24 + 64)
    d_in = torch.zeros((n2, ip), dtype=torch.uint8, device=dev)
    d_in[:, :n1 * 3] = torch.from_numpy(img.reshape(n2, n1 * 3)).to(dev)
    d_out = torch.full((n2, op // 8), -7.0, dtype=torch.float64, device=dev)
    ctx._ck(ctx._L.mr_blur_dev(ctx._h, d_in.data_ptr(), n1, n2, ip, 3, R, 3, d_out.data_ptr(), op,
                               torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = d_out[:, :n1 * 3].cpu().numpy().reshape(n2, n1, 3)
    assert got.tobytes() == oracle.blur(img, R, 3).tobytes()
    assert bool((d_out[:, n1 * 3:] == -7.0).all())        # nothing written outside the lines


def test_reference_vectors_through_the_float64_path(ctx, mr, oracle):
    mean, mn, mx, sd, tol, vectors = golden_palette()
    pal = mr.Palette(mean, mn, mx, sd, np.full(mean.shape[0], tol))
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    ctx.set_known(None)
    px = np.array([[v["rgb"] for v in vectors]], dtype=np.float64)
    want = np.array([[v["expected"] for v in vectors]], dtype=np.uint8)
    assert (ctx.categorize_clean_f64_host(px, 0) == want).all()


@pytest.mark.parametrize("R", [0, 1, 2, 4])
@pytest.mark.parametrize("seed", [0, 1])
def test_float64_categorise_clean_random(ctx, mr, oracle, R, seed):
    rng = np.random.default_rng(900 + seed)
    n2, n1, K = 97, 131, 9
    mean, mn, mx, sd, tol = random_palette(rng, K)
    pal = mr.Palette(mean, mn, mx, sd, tol)
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    # colours near the means (inside and just outside the boxes) plus uniform noise
    which = rng.integers(0, K, size=(n2, n1))
    px = mean[which] + rng.normal(0, 0.05, size=(n2, n1, 3))
    px[rng.random((n2, n1)) < 0.1] = rng.random(3)
    idx = np.unique(rng.integers(0, n1 * n2, size=60))
    cat = rng.integers(1, K + 1, size=idx.size).astype(np.uint8)
    for known in (False, True):
        ctx.set_known(idx, cat) if known else ctx.set_known(None)
        try:
            got = ctx.categorize_clean_f64_host(px, R)
        finally:
            ctx.set_known(None)
        lab = oracle.categorize_f64_image(px, pal.mean, pal.lo, pal.hi, idx if known else None, cat if known else None)
        want = oracle.majority(lab, R) if R else lab
        assert mismatches(got, want) == 0


def test_mirror_blur_then_categorize(mr, oracle):
    rng = np.random.default_rng(7)
    n2, n1, K = 60, 75, 5
    pal_rgb = rng.integers(0, 256, size=(K, 3), dtype=np.uint8)
    img = np.clip(pal_rgb[rng.integers(0, K, size=(n2 // 5, n1 // 5))].repeat(5, 0).repeat(5, 1).astype(np.int32)
                  + rng.integers(-9, 10, size=(n2, n1, 3)), 0, 255).astype(np.uint8)
    b = mr.blur(img, hood=mr.Window(1), repeat=2)
    assert b.dtype == np.float64 and b.tobytes() == oracle.blur(img, 1, 2).tobytes()
    points = [[(float(5 * x + 3), float(5 * y + 3)) for y in range(n2 // 5) for x in range(n1 // 5)
               if (img[5 * y + 2, 5 * x + 2].astype(int) - pal_rgb[c]).__abs__().sum() < 30][:8] for c in range(K)]
    lab = mr.categorize(b, points)
    pal = mr.Palette.from_points(b, points)
    assert mismatches(lab, oracle.categorize_f64_image(b, pal.mean, pal.lo, pal.hi)) == 0
    with pytest.raises(ValueError):
        mr.categorize(b.astype(np.float32), points)


@pytest.mark.parametrize("shape", [(1, 1), (3, 255), (5, 256), (7, 257), (40, 1000), (64, 2049)])
def test_balance_bit_identical(ctx, oracle, shape):
    """per-pixel part of `_balance` (src/balance.jl:25-42): predictions, their mean (fixed summation order), clamp"""
    rng = np.random.default_rng(shape[1])
    n2, n1 = shape
    img = rng.integers(0, 256, size=(n2, n1, 3), dtype=np.uint8)
    coef = np.stack([rng.normal(0, s, size=3) for s in (0.05, 1e-10, 1e-7, 1e-4, 1e-10, 1e-7, 1e-4)], axis=1)   # (3, 7)
    got = ctx.balance_host(img, coef)
    want = oracle.balance(img, coef)
    assert got.tobytes() == want.tobytes()
    assert got.min() >= 0.0 and got.max() <= 1.0
    with pytest.raises(Exception):
        ctx.balance_host(img, np.full((3, 7), np.nan))


def test_balance_mirror(mr, oracle):
    rng = np.random.default_rng(3)
    n2, n1 = 90, 120
    yy, xx = np.mgrid[0:n2, 0:n1]
    img = np.clip(120 + 0.3 * xx[..., None] - 0.2 * yy[..., None] + rng.integers(-5, 6, size=(n2, n1, 3)), 0, 255).astype(np.uint8)
    pts = [[(float(rng.integers(1, n1 + 1)), float(rng.integers(1, n2 + 1))) for _ in range(12)], []]
    out = mr.balance(img, pts)
    coef = mr.balance_fit(img, pts)
    assert out.tobytes() == oracle.balance(img, coef).tobytes()
    # fewer than 8 points: A .* 1.0
    few = mr.balance(img, [pts[0][:5], []])
    assert few.tobytes() == (img.astype(np.float64) / 255.0).tobytes()
    # the linear trend is gone: the detrended image varies much less along the axes than the input
    assert abs(np.polyfit(np.arange(n1), out[:, :, 0].mean(axis=0), 1)[0]) < 0.1 * abs(np.polyfit(np.arange(n1), img[:, :, 0].mean(axis=0) / 255.0, 1)[0])


def test_float64_categoriser_adversarial(ctx, mr, oracle):
    """adversarial palettes and colours for the Float64 categoriser: duplicate means (first of equal errors
    wins), exact ties between lattice means, a missing category (+Inf), colours far outside [0, 1], a NaN
    colour, K = 1 and K = 64.  (A variant that skipped categories which provably cannot win -- triangle
    inequality on a table of distances between the means -- passed this test too but was 1.7 x slower than
    evaluating everything: the candidate sets differ from lane to lane.)"""
    rng = np.random.default_rng(123)
    n2, n1 = 64, 96
    for case in range(6):
        if case == 0:      # duplicates and near-duplicates
            base = rng.random((5, 3))
            mean = np.concatenate([base, base, base + 1e-15, base[::-1]])
        elif case == 1:    # lattice: many exact ties
            g = np.array([0.25, 0.5, 0.75])
            mean = np.array([(a, b, c) for a in g for b in g for c in g])
        elif case == 2:    # a missing category in front and in the middle
            mean = rng.random((9, 3))
            mean[0] = np.inf
            mean[4] = np.inf
        elif case == 3:
            mean = rng.random((1, 3))
        elif case == 4:
            mean = rng.random((64, 3))
        else:              # clustered means: the guess is often wrong
            mean = 0.5 + rng.normal(0, 0.01, size=(40, 3))
        K = mean.shape[0]
        lo = np.where(np.isfinite(mean), mean - 0.2, np.nan)
        hi = np.where(np.isfinite(mean), mean + 0.2, np.nan)
        ctx.set_palette(mean, lo, hi)
        ctx.palette = None
        ctx.set_known(None)
        fin = np.where(np.isfinite(mean).all(axis=1))[0]
        px = mean[rng.choice(fin, size=(n2, n1))] + rng.normal(0, 0.02, size=(n2, n1, 3))
        px[::7, ::5] = mean[fin[0]]                                  # exactly on a mean
        if case == 1:
            px[1::3] = np.round(px[1::3] * 8) / 8                      # exactly between lattice means
        px[3, 3] = (1e6, -1e6, 3.0)
        px[5, 5] = (np.nan, 0.1, 0.2)
        got = ctx.categorize_clean_f64_host(px, 0)
        want = oracle.categorize_f64_image(px, mean, lo, hi)
        assert mismatches(got, want) == 0, case
