"""GPU (-m gpu): the texture features of the reference's default match = (:color, :std, :stripe) --
`_stds` (/root/reference/src/selectcolors.jl:706-715), `_stripes` (src/stripes.jl:20-35) -- and the categoriser on
PixelProp pixels (src/categories.jl:76-133 with a match tuple) through the C ABI: Float64 planes bit-identical
to the oracle, labels identical.  Parity unpinned by reference fixtures (none exist)."""
import numpy as np
import pytest
import torch

from util import mismatches, random_palette

pytestmark = pytest.mark.gpu


def blurred_like(rng, shape):
    """a plausible blurred scan: smooth patches + noise, some grey and some bright pixels"""
    n2, n1 = shape
    base = rng.random((n2 // 8 + 2, n1 // 8 + 2, 3))
    img = np.repeat(np.repeat(base, 8, axis=0), 8, axis=1)[:n2, :n1] * 0.8 + rng.random((n2, n1, 3)) * 0.2
    img[rng.random((n2, n1)) < 0.05] = rng.random()            # grey pixels: the s = 0 branch of HSL
    return np.ascontiguousarray(img)


@pytest.mark.parametrize("shape", [(1, 1), (1, 40), (33, 1), (37, 53), (70, 600)])
def test_stds_bit_identical(ctx, oracle, shape):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    px = blurred_like(rng, shape)
    with np.errstate(invalid="ignore"):
        want = oracle.stds(px)
    got = ctx.stds_host(px)
    assert got.tobytes() == want.tobytes()


def test_stds_of_a_flat_image_is_nan_like_julia(ctx, oracle):
    px = np.zeros((9, 12, 3))          # every std is 0: 0 / maximum = 0 / 0 = NaN
    with np.errstate(invalid="ignore"):
        want = oracle.stds(px)
    got = ctx.stds_host(px)
    assert np.isnan(got).all() and got.tobytes() == want.tobytes()


@pytest.mark.parametrize("radius", [1, 2, 3, 7])
@pytest.mark.parametrize("shape", [(1, 1), (1, 300), (33, 2), (37, 53), (40, 600)])
def test_stripes_bit_identical(ctx, oracle, radius, shape):
    rng = np.random.default_rng(radius * 1000 + shape[0] * 7 + shape[1])
    px = blurred_like(rng, shape)
    px[0, 0] = (0.9, 0.95, 0.99)       # l >= 0.5 branch
    with np.errstate(invalid="ignore", divide="ignore"):
        want = oracle.stripes(px, radius)
    got = ctx.stripes_host(px, radius)
    assert got.tobytes() == want.tobytes()


def test_texture_device_entries_with_pitches(ctx, oracle):
    rng = np.random.default_rng(17)
    n2, n1 = 45, 77
    px = blurred_like(rng, (n2, n1))
    dev = torch.device("cuda", 0)
    ip = n1 * 24 + 40
    d_in = torch.zeros((n2, ip // 8), dtype=torch.float64, device=dev)
    d_in[:, :n1 * 3] = torch.from_numpy(px.reshape(n2, n1 * 3)).to(dev)
    s = torch.cuda.current_stream().cuda_stream
    op = n1 * 8 + 24
    d_out = torch.full((n2, op // 8), -7.0, dtype=torch.float64, device=dev)
    ctx._ck(ctx._L.mr_stds_dev(ctx._h, d_in.data_ptr(), n1, n2, ip, d_out.data_ptr(), op, s))
    torch.cuda.synchronize()
    assert d_out[:, :n1].cpu().numpy().tobytes() == oracle.stds(px).tobytes()
    assert bool((d_out[:, n1:] == -7.0).all())              # nothing written outside the lines
    op = n1 * 16 + 8
    d_out = torch.full((n2, op // 8), -7.0, dtype=torch.float64, device=dev)
    ctx._ck(ctx._L.mr_stripes_dev(ctx._h, d_in.data_ptr(), n1, n2, ip, 2, d_out.data_ptr(), op, s))
    torch.cuda.synchronize()
    assert d_out[:, :2 * n1].cpu().numpy().reshape(n2, n1, 2).tobytes() == oracle.stripes(px, 2).tobytes()
    assert bool((d_out[:, 2 * n1:] == -7.0).all())


def test_props_device_entry_with_pitches(ctx, mr, oracle):
    """device-resident planes with padded lines; nothing is written outside the label lines"""
    rng = np.random.default_rng(23)
    n2, n1, K = 41, 203, 5
    mean, mn, mx, sd, tol = random_palette(rng, K)
    pal = mr.Palette(mean, mn, mx, sd, tol)
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    ctx.set_known(None)
    sdm, stm = rng.random(K), rng.normal(0, 1, (K, 2))
    sds, sts = (sdm, sdm - 0.3, sdm + 0.3), (stm, stm - 1.0, stm + 1.0)
    ctx.set_texture_stats(*sds, *sts)
    which = rng.integers(0, K, size=(n2, n1))
    px = mean[which] + rng.normal(0, 0.03, size=(n2, n1, 3))
    std = sdm[which] + rng.normal(0, 0.1, size=(n2, n1))
    stripe = stm[which] + rng.normal(0, 0.35, size=(n2, n1, 2))
    dev = torch.device("cuda", 0)

    def padded(a, extra):
        flat = a.reshape(n2, -1)
        t = torch.full((n2, flat.shape[1] + extra), 9.0, dtype=torch.float64, device=dev)
        t[:, :flat.shape[1]] = torch.from_numpy(flat).to(dev)
        return t
    d_px, d_sd, d_st = padded(px, 5), padded(std, 3), padded(stripe, 7)
    s = torch.cuda.current_stream().cuda_stream
    for R in (0, 2):
        d_out = torch.full((n2, n1 + 29), 77, dtype=torch.uint8, device=dev)
        ctx._ck(ctx._L.mr_categorize_clean_props_dev(ctx._h, d_px.data_ptr(), d_px.shape[1] * 8, d_sd.data_ptr(), d_sd.shape[1] * 8,
                                                     d_st.data_ptr(), d_st.shape[1] * 8, n1, n2, 7, R, d_out.data_ptr(),
                                                     d_out.shape[1], s))
        torch.cuda.synchronize()
        lab = oracle.categorize_props_image(px, std, stripe, ("color", "std", "stripe"), pal.mean, pal.lo, pal.hi, sds, sts)
        want = oracle.majority(lab, R) if R else lab
        assert mismatches(d_out[:, :n1].cpu().numpy(), want) == 0
        assert bool((d_out[:, n1:] == 77).all())


MATCHES = [("color",), ("color", "std"), ("color", "stripe"), ("color", "std", "stripe"), ("std", "stripe"), ("std",),
           ("stripe",)]


@pytest.mark.parametrize("match", MATCHES)
@pytest.mark.parametrize("R", [0, 2])
def test_props_categorise_clean_random(ctx, mr, oracle, match, R):
    rng = np.random.default_rng(1300 + len(match) * 10 + R + len(match[0]))
    n2, n1, K = 83, 121, 7
    mean, mn, mx, sd, tol = random_palette(rng, K)
    mean[4] = np.inf                                          # a `missing` category
    pal = mr.Palette(mean, mn, mx, sd, tol)
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    sdm = rng.random(K) * 0.6
    stm = rng.normal(0, 1, (K, 2))
    sds = (sdm, sdm - 0.1 - rng.random(K) * 0.3, sdm + 0.1 + rng.random(K) * 0.3)
    sts = (stm, stm - 0.4 - rng.random((K, 2)), stm + 0.4 + rng.random((K, 2)))
    ctx.set_texture_stats(*sds, *sts)
    which = rng.integers(0, K, size=(n2, n1))
    which[which == 4] = 0
    px = mean[which] + rng.normal(0, 0.03, size=(n2, n1, 3))
    px[rng.random((n2, n1)) < 0.1] = rng.random(3)
    std = sdm[which] + rng.normal(0, 0.1, size=(n2, n1))
    stripe = stm[which] + rng.normal(0, 0.35, size=(n2, n1, 2))
    idx = np.unique(rng.integers(0, n1 * n2, size=60))
    cat = rng.integers(1, K + 1, size=idx.size).astype(np.uint8)
    bits = sum(mr.api.MATCH_BITS[m] for m in match)
    for known in (False, True):
        ctx.set_known(idx, cat) if known else ctx.set_known(None)
        try:
            got = ctx.categorize_clean_props_host(px, std if bits & 2 else None, stripe if bits & 4 else None, bits, R)
        finally:
            ctx.set_known(None)
        lab = oracle.categorize_props_image(px, std, stripe, match, pal.mean, pal.lo, pal.hi, sds, sts,
                                            idx if known else None, cat if known else None)
        want = oracle.majority(lab, R) if R else lab
        assert mismatches(got, want) == 0
        assert R > 0 or len(np.unique(got)) > 2
    if match == ("color",):                                   # bit 1 alone = the RGB{Float64} categoriser
        assert mismatches(ctx.categorize_clean_props_host(px, None, None, 1, R), ctx.categorize_clean_f64_host(px, R)) == 0


@pytest.mark.parametrize("n1", [1023, 1300, 4100])
def test_props_wide_lines(ctx, mr, oracle, n1):
    """lines wider than one block of threads (several blocks per line, ragged last word)"""
    rng = np.random.default_rng(n1)
    n2, K = 5, 6
    mean, mn, mx, sd, tol = random_palette(rng, K)
    pal = mr.Palette(mean, mn, mx, sd, tol)
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    ctx.set_known(None)
    sdm, stm = rng.random(K), rng.normal(0, 1, (K, 2))
    sds, sts = (sdm, sdm - 0.3, sdm + 0.3), (stm, stm - 1.0, stm + 1.0)
    ctx.set_texture_stats(*sds, *sts)
    which = rng.integers(0, K, size=(n2, n1))
    px = mean[which] + rng.normal(0, 0.03, size=(n2, n1, 3))
    std = sdm[which] + rng.normal(0, 0.1, size=(n2, n1))
    stripe = stm[which] + rng.normal(0, 0.35, size=(n2, n1, 2))
    for match in (("color",), ("color", "std", "stripe")):
        bits = sum(mr.api.MATCH_BITS[m] for m in match)
        got = ctx.categorize_clean_props_host(px, std if bits & 2 else None, stripe if bits & 4 else None, bits, 0)
        want = oracle.categorize_props_image(px, std, stripe, match, pal.mean, pal.lo, pal.hi, sds, sts)
        assert mismatches(got, want) == 0
        assert (got[:, -8:] != 0).any() and (got[:, :8] != 0).any()
    assert mismatches(ctx.categorize_clean_f64_host(px, 0), oracle.categorize_f64_image(px, pal.mean, pal.lo, pal.hi)) == 0
    with np.errstate(invalid="ignore", divide="ignore"):
        assert ctx.stds_host(px).tobytes() == oracle.stds(px).tobytes()
        assert ctx.stripes_host(px, 3).tobytes() == oracle.stripes(px, 3).tobytes()


@pytest.mark.parametrize("bits", [3, 5, 7])
def test_props_near_ties_and_tiny_sums(ctx, mr, oracle, bits):
    """adversarial for the division-free loop of the PixelProp categoriser: pairs of categories whose undivided error
    sums are equal or a few ulps apart (the quotient by 2 / 3 may or may not merge them: the FIRST minimum of the
    rounded quotients decides), and sums near the subnormal range (the literal loop)"""
    rng = np.random.default_rng(bits)
    K, n2, n1 = 6, 64, 512
    base = rng.random(3) * 0.5 + 0.25
    d = 0.0625
    mean = np.stack([base + (d, 0, 0), base - (d, 0, 0), base + (0, d, 0), base - (0, d, 0), base + (0, 0, d), base - (0, 0, d)])
    lo, hi = mean - 1.0, mean + 1.0
    pal = mr.Palette(mean, lo, hi, np.zeros((K, 3)))
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    ctx.set_known(None)
    sdm, stm = np.full(K, 0.3), np.tile([0.1, -0.2], (K, 1))
    sds, sts = (sdm, sdm - 9, sdm + 9), (stm, stm - 9, stm + 9)
    ctx.set_texture_stats(*sds, *sts)
    # pixels at the centre +- a few ulps per channel: the six errors are equal up to rounding
    ulps = rng.integers(-3, 4, size=(n2, n1, 3))
    px = base + ulps * np.spacing(base)
    px[: n2 // 4] = mean[rng.integers(0, K, size=(n2 // 4, n1))] + rng.integers(-2, 3, size=(n2 // 4, n1, 3)) * 1e-155   # tiny sums
    std = 0.3 + rng.integers(-2, 3, size=(n2, n1)) * np.spacing(0.3)
    stripe = np.array([0.1, -0.2]) + rng.integers(-2, 3, size=(n2, n1, 2)) * np.spacing(0.1)
    std[: n2 // 4] = 0.3
    stripe[: n2 // 4] = (0.1, -0.2)
    match = tuple(n for k, n in enumerate(("color", "std", "stripe")) if bits >> k & 1)
    got = ctx.categorize_clean_props_host(px, std if bits & 2 else None, stripe if bits & 4 else None, bits, 0)
    want = oracle.categorize_props_image(px, std, stripe, match, pal.mean, pal.lo, pal.hi, sds, sts)
    assert mismatches(got, want) == 0
    assert len(np.unique(got)) >= 4


def test_props_quotient_rounding_decides_the_first_minimum(ctx, mr, oracle):
    """categories that differ only by two ulps of one stripe mean: their undivided sums are equal or one ulp apart,
    and dividing by 3 merges some of those pairs -- then the EARLIER category wins although its sum is larger.
    The division-free loop must reproduce that (its remembered rival); the test first checks that the data holds
    such pixels."""
    rng = np.random.default_rng(5)
    K, n2, n1 = 4, 100, 2000
    cm = np.tile(rng.random(3), (K, 1))
    pal = mr.Palette(cm, cm - 9, cm + 9, np.zeros((K, 3)))
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    ctx.set_known(None)
    sdm = np.full(K, 0.3)
    stm = np.tile([0.1, -0.2], (K, 1)).astype(np.float64)
    for c in range(K):
        stm[c, 0] = 0.1 + (K - 1 - c) * 2 * np.spacing(0.1)
    sds, sts = (sdm, sdm - 9, sdm + 9), (stm, stm - 9, stm + 9)
    ctx.set_texture_stats(*sds, *sts)
    px = cm[0] + rng.normal(0, 0.3, (n2, n1, 3))
    std = 0.3 + rng.normal(0, 0.2, (n2, n1))
    st = np.array([0.1, -0.2]) + rng.normal(0, 0.3, (n2, n1, 2))
    sums = []
    for c in range(K):
        dd = px - cm[c]
        e = (dd[..., 0] * dd[..., 0] + dd[..., 1] * dd[..., 1]) + dd[..., 2] * dd[..., 2]
        t = std - sdm[c]
        e = e + t * t
        t0, t1 = st[..., 0] - stm[c, 0], st[..., 1] - stm[c, 1]
        sums.append(e + (t0 * t0 + t1 * t1) * 0.5)
    S = np.stack(sums)
    assert int((S.argmin(0) != (S / 3).argmin(0)).sum()) > 1000         # the rounding of the quotient matters here
    got = ctx.categorize_clean_props_host(px, std, st, 7, 0)
    want = oracle.categorize_props_image(px, std, st, ("color", "std", "stripe"), pal.mean, pal.lo, pal.hi, sds, sts)
    assert mismatches(got, want) == 0
    assert mismatches(want, (S / 3).argmin(0) + 1) == 0                      # and the oracle is the literal rule


def test_props_need_texture_stats_and_planes(ctx, mr):
    mean, mn, mx, sd, tol = random_palette(np.random.default_rng(1), 3)
    pal = mr.Palette(mean, mn, mx, sd, tol)
    ctx.set_palette(pal.mean, pal.lo, pal.hi)                 # a new palette drops the texture statistics
    ctx.palette = None
    px = np.zeros((4, 5, 3))
    with pytest.raises(mr.MapRastError, match="mr_set_texture_stats"):
        ctx.categorize_clean_props_host(px, np.zeros((4, 5)), None, 3, 0)
    z1, z2 = np.zeros(3), np.zeros((3, 2))
    ctx.set_texture_stats(z1, z1, z1, z2, z2, z2)
    with pytest.raises(mr.MapRastError, match="null"):
        ctx.categorize_clean_props_host(px, None, None, 3, 0)
    with pytest.raises(mr.MapRastError, match="match"):
        ctx.categorize_clean_props_host(px, None, None, 0, 0)
    with pytest.raises(mr.MapRastError, match="radius"):
        ctx.stripes_host(px, 0)
    assert ctx.stds_host(np.zeros((0, 0, 3))).shape == (0, 0)


def test_rasterize_with_the_default_match(ctx, mr, oracle):
    """the reference's DEFAULT settings (match = (:color, :std, :stripe), blur_repeat = 1, stripe_radius = 2)
    through `rasterize`: balance (too few points: A .* 1.0) -> blur -> stds / stripes -> statistics at the clicked
    points -> PixelProp categorise, against the same chain on the oracle"""
    from maprast_b200 import synth
    seed, pal_rgb, stats, spec = synth.config_workload(1)
    n1, n2 = 240, 180
    img = oracle.synth_scan(seed, 0, pal_rgb, n1, n2)
    rng = np.random.default_rng(4)
    K = 4
    points = [[(float(rng.integers(3, n1 - 2)), float(rng.integers(3, n2 - 2))) for _ in range(6)] for _ in range(K)]
    tol = np.full(K, 3.0)
    sel = mr.MapSelection(settings={"match": ("color", "std", "stripe"), "blur_repeat": 1, "stripe_radius": 2,
                                    "category_tolerances": tol, "category_balance": [False] * K,
                                    "prune_threshold": 0, "line_threshold": 0.0, "fill_repeat": 0}, points=points)
    got = mr.rasterize(img, sel, ctx=ctx)["categorized"]
    blurred = oracle.blur(img, 1, 1)
    std, stripe = oracle.stds(blurred), oracle.stripes(blurred, 2)
    pal = mr.Palette.from_points(blurred, points, tol)
    tex = mr.TextureStats.from_points(std, stripe, points, tol)
    (slo, shi), (tlo, thi) = tex.bounds()
    kidx, kcat = mr.api._known_from_points(points, n1)
    want = oracle.categorize_props_image(blurred, std, stripe, ("color", "std", "stripe"), pal.mean, pal.lo, pal.hi,
                                         (tex.std["mean"], slo, shi), (tex.stripe["mean"], tlo, thi), kidx, kcat)
    assert mismatches(got, want) == 0
    assert (got != 0).sum() > 0
