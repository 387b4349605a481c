"""GPU (-m gpu): parity of libmaprast.so (called through its C ABI) against the CPU oracle.
Bit-exact: every comparison is array equality of uint8 labels / bytes; mismatch target is 0."""
import numpy as np
import pytest
import torch

from util import golden_palette, mismatches, random_palette

pytestmark = pytest.mark.gpu


def _workload(mr, cfg, colourspace=None):
    from maprast_b200 import synth
    seed, pal, stats, spec = synth.config_workload(cfg)
    return seed, pal, stats, spec


def _set(ctx, mr, stats):
    ctx.set_known(None)
    ctx.set_palette(stats.mean, stats.lo, stats.hi, mr._lib.LAB if stats.colourspace == "lab" else mr._lib.RGB)
    ctx.palette = stats


def _ocs(oracle, stats):
    return oracle.LAB if stats.colourspace == "lab" else oracle.RGB


# ---------------------------------------------------------------- exhaustive colour sweep
@pytest.mark.parametrize("cfg", [1, 2, 4])
def test_exhaustive_2p24_colour_sweep(ctx, mr, oracle, cfg):
    """all 2^24 colours x benchmark palette, device table vs oracle."""
    _, _, stats, _ = _workload(mr, cfg)
    _set(ctx, mr, stats)
    got = ctx.get_lut()
    want = oracle.build_lut(stats.mean, stats.lo, stats.hi, _ocs(oracle, stats))
    assert mismatches(got, want) == 0
    assert 0 < ctx.palette_build_ms() < 1000


def test_exhaustive_sweep_adversarial_palettes(ctx, mr, oracle):
    from maprast_b200 import synth
    lat = synth.make_palette(11, 16, "lattice")
    st = synth.make_stats(11, lat, exact_means=True)          # exact distance ties everywhere
    _set(ctx, mr, st)
    assert mismatches(ctx.get_lut(), oracle.build_lut(st.mean, st.lo, st.hi)) == 0
    rng = np.random.default_rng(5)
    mean, mn, mx, sd, tol = random_palette(rng, 37, spread=0.3)
    mean[3] = np.inf                                          # a missing category
    sd[5] = np.nan                                            # a one-point category
    pal = mr.Palette(mean, mn, mx, sd, tol)
    _set(ctx, mr, pal)
    assert mismatches(ctx.get_lut(), oracle.build_lut(pal.mean, pal.lo, pal.hi)) == 0


def test_reference_golden_vectors_on_device(ctx, mr, oracle):
    """The reference's four known-answer colours are Float64, not N0f8; the nearest N0f8 colours
    of vectors 3 and 4 (exactly representable cases aside) must agree with the oracle, and the
    whole golden palette table must match."""
    mean, mn, mx, sd, tol, vectors = golden_palette()
    pal = mr.Palette(mean, mn, mx, sd, [tol, tol])
    _set(ctx, mr, pal)
    lut = ctx.get_lut()
    assert mismatches(lut, oracle.build_lut(pal.mean, pal.lo, pal.hi)) == 0
    for v in vectors:
        b = [int(round(c * 255)) for c in v["rgb"]]
        x = [c / 255.0 for c in b]
        assert lut[b[0] | b[1] << 8 | b[2] << 16] == oracle.categorize_f64(x, pal.mean, pal.lo, pal.hi)


# ---------------------------------------------------------------- fused image parity
@pytest.mark.parametrize("cfg,n1,n2", [(1, 2000, 2000), (2, 1531, 777), (4, 1024, 300), (5, 512, 512)])
def test_fused_matches_oracle(ctx, mr, oracle, cfg, n1, n2):
    """config 1 in full; crops (incl. all four borders) of configs 2/4/5."""
    seed, palrgb, stats, spec = _workload(mr, cfg)
    _set(ctx, mr, stats)
    scan = oracle.synth_scan(seed, 0, palrgb, n1, n2)
    got, cnt = ctx.categorize_clean_host(scan, spec["R"], counters=True)
    want = oracle.categorize_clean(scan, stats.mean, stats.lo, stats.hi, spec["R"], _ocs(oracle, stats))
    assert mismatches(got, want) == 0
    pre = oracle.categorize_u8(scan, stats.mean, stats.lo, stats.hi, _ocs(oracle, stats))
    assert cnt["n_pixels"] == n1 * n2
    assert cnt["n_nomatch"] == int((want == 0).sum())
    assert cnt["n_changed"] == int((want != pre).sum())
    assert cnt["n_fine"] <= n1 * n2


@pytest.mark.parametrize("R", [0, 1, 2, 3, 5, 7])
@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (9, 1), (3, 5), (16, 16), (17, 31), (130, 67), (257, 129)])
def test_ragged_and_tiny_shapes(ctx, mr, oracle, R, shape):
    n1, n2 = shape
    seed, palrgb, stats, _ = _workload(mr, 2)
    _set(ctx, mr, stats)
    scan = oracle.synth_scan(seed + R, 3, palrgb, n1, n2)
    got = ctx.categorize_clean_host(scan, R)
    want = oracle.categorize_clean(scan, stats.mean, stats.lo, stats.hi, R)
    assert mismatches(got, want) == 0


def test_empty_inputs(ctx, mr):
    _, _, stats, _ = _workload(mr, 1)
    _set(ctx, mr, stats)
    assert ctx.categorize_clean_host(np.zeros((0, 5, 3), np.uint8), 1).shape == (0, 5)
    assert ctx.categorize_clean_host(np.zeros((5, 0, 3), np.uint8), 1).shape == (5, 0)
    assert ctx.clean_host(np.zeros((0, 0), np.uint8), 1).shape == (0, 0)


def test_uniform_random_colours_and_lattice(ctx, mr, oracle):
    """Adversarial sets: worst case for the coarse-cell filter (every colour distinct) and the
    lattice palette (exact ties)."""
    from maprast_b200 import synth
    seed, palrgb, stats, _ = _workload(mr, 2)
    _set(ctx, mr, stats)
    scan = oracle.synth_scan(seed, 0, palrgb, 640, 480, mode=1)
    assert mismatches(ctx.categorize_clean_host(scan, 2),
                      oracle.categorize_clean(scan, stats.mean, stats.lo, stats.hi, 2)) == 0
    lat = synth.make_palette(11, 16, "lattice")
    st = synth.make_stats(11, lat, exact_means=True)
    _set(ctx, mr, st)
    scan = oracle.synth_scan(seed, 0, lat, 640, 480)
    assert mismatches(ctx.categorize_clean_host(scan, 1),
                      oracle.categorize_clean(scan, st.mean, st.lo, st.hi, 1)) == 0


@pytest.mark.parametrize("R", [1, 2, 3])
def test_random_labels_every_word_undecided(ctx, mr, oracle, R):
    """Uniform random colours under a palette whose boxes accept everything: the label of every
    pixel is random, no window has a clear majority, every vote word overflows into the round-2 /
    exact-vote lists (the capped list cuts the blocks short) and every MIXED cell is hit."""
    from maprast_b200 import synth
    seed, palrgb, _, _ = _workload(mr, 2)
    st = synth.make_stats(seed, palrgb, tol=60.0)
    _set(ctx, mr, st)
    scan = oracle.synth_scan(seed, 3, palrgb, 1936, 200, mode=1)
    pre = oracle.categorize_u8(scan, st.mean, st.lo, st.hi)
    assert (pre == 0).mean() < 0.05 and len(np.unique(pre)) >= 16
    got, cnt = ctx.categorize_clean_host(scan, R, counters=True)
    want = oracle.categorize_clean(scan, st.mean, st.lo, st.hi, R)
    assert mismatches(got, want) == 0
    assert cnt["n_changed"] == int((want != pre).sum()) and cnt["n_nomatch"] == int((want == 0).sum())


@pytest.mark.parametrize("K", [2, 14, 15, 30, 31, 62, 63, 126, 127, 200])
def test_palette_sizes_cover_every_plane_count(ctx, mr, oracle, K):
    """K decides the number of label bit planes of the fast kernel (4: K <= 14, 5: <= 30, 6: <= 62,
    7: <= 126) and K > 126 takes the generic kernel: both sides of every boundary, windows 3 / 5 / 7."""
    from maprast_b200 import synth
    seed = 0x5EED0000 + K
    pal = synth.make_palette(seed, K)
    st = synth.make_stats(seed, pal)
    _set(ctx, mr, st)
    n1, n2 = 1936, 96
    scan = oracle.synth_scan(seed, 0, pal, n1, n2)
    pre = oracle.categorize_u8(scan, st.mean, st.lo, st.hi)
    assert pre.max() >= min(K, 100)
    for R in (1, 2, 3):
        got = ctx.categorize_clean_host(scan, R)
        assert mismatches(got, oracle.majority(pre, R)) == 0


def test_strided_and_unaligned_views(ctx, mr, oracle):
    seed, palrgb, stats, _ = _workload(mr, 1)
    _set(ctx, mr, stats)
    big = oracle.synth_scan(seed, 0, palrgb, 333, 100)
    view = big[5:90, 7:300]                     # line stride != n1*3, start not 16-byte aligned
    got = ctx.categorize_clean_host(view, 1)
    want = oracle.categorize_clean(np.ascontiguousarray(view), stats.mean, stats.lo, stats.hi, 1)
    assert mismatches(got, want) == 0


@pytest.mark.parametrize("n1,pitch_px", [(1001, 1001), (1937, 2000), (333, 340)])
def test_device_entry_points_with_unaligned_lines(ctx, mr, oracle, n1, pitch_px):
    """Device pointers whose lines are not 16-byte aligned (odd n1, odd pitches, offset bases): the
    library repacks them into aligned pitches and still runs the fast kernel; bands with halos too."""
    seed, palrgb, stats, spec = _workload(mr, 2)
    _set(ctx, mr, stats)
    n2, R = 150, 2
    scan = oracle.synth_scan(seed, 5, palrgb, n1, n2)
    want = oracle.categorize_clean(scan, stats.mean, stats.lo, stats.hi, R)
    pre = oracle.categorize_u8(scan, stats.mean, stats.lo, stats.hi)
    buf = torch.zeros((n2, pitch_px, 3), dtype=torch.uint8, device="cuda")
    buf[:, :n1] = torch.from_numpy(scan).cuda()
    out = torch.full((n2, pitch_px + 3), 77, dtype=torch.uint8, device="cuda")
    l0 = ctx.launch_count()
    ctx.categorize_clean_dev(buf, n1, n2, pitch_px * 3, 3, R, out, pitch_px + 3)
    torch.cuda.synchronize()
    assert mismatches(out[:, :n1].cpu().numpy(), want) == 0
    assert bool((out[:, n1:] == 77).all())                      # nothing written outside the image
    assert ctx.launch_count() - l0 == 1
    # a band with halos out of the same buffer
    a, b = 40, 110
    band = torch.full((b - a, n1), 9, dtype=torch.uint8, device="cuda")
    ctx.categorize_clean_dev(buf[a:], n1, b - a, pitch_px * 3, 3, R, band, n1, halo_lo=R, halo_hi=R)
    torch.cuda.synchronize()
    assert mismatches(band.cpu().numpy(), want[a:b]) == 0
    # standalone clean on unaligned label lines
    lab = torch.zeros((n2, pitch_px), dtype=torch.uint8, device="cuda")
    lab[:, :n1] = torch.from_numpy(pre).cuda()
    out2 = torch.full((n2, pitch_px + 1), 55, dtype=torch.uint8, device="cuda")
    ctx.clean_dev(lab, n1, n2, pitch_px, R, out2, pitch_px + 1)
    torch.cuda.synchronize()
    assert mismatches(out2[:, :n1].cpu().numpy(), want) == 0 and bool((out2[:, n1:] == 55).all())
    assert mismatches(ctx.clean_host(pre, R), want) == 0


def test_rgba_direct_path(ctx, mr, oracle):
    rng = np.random.default_rng(8)
    mean, mn, mx, sd, tol = random_palette(rng, 5, C=4, spread=0.4)
    pal = mr.Palette(mean, mn, mx, sd, tol)
    ctx.set_known(None)
    ctx.set_palette(pal.mean, pal.lo, pal.hi, mr._lib.RGB, 4)
    ctx.palette = None
    px = rng.integers(0, 256, size=(60, 77, 4), dtype=np.uint8)
    px[::7, ::5, 3] = 0                         # transparent pixels never match (categories.jl:116)
    got = ctx.categorize_clean_host(px, 1)
    want = oracle.categorize_clean(px, pal.mean, pal.lo, pal.hi, 1, channels=4)
    assert mismatches(got, want) == 0
    with pytest.raises(mr.MapRastError):        # RGB pixels with RGBA statistics
        ctx.categorize_clean_host(px[..., :3].copy(), 1)


@pytest.mark.parametrize("transparent", [0.0, 0.02, 0.6])
def test_rgba_sheets_take_the_fast_kernels(ctx, mr, oracle, transparent):
    """RGBA{N0f8} with 4-channel statistics (alpha joins the distance, src/categories.jl:92; alpha == 0
    never matches, :115-118): alpha byte stripped, label table for opaque pixels, direct re-evaluation
    of the others, standalone majority filter.  Opaque, sprinkled and mostly transparent sheets."""
    from maprast_b200 import synth
    seed, palrgb, stats3, spec = _workload(mr, 2)
    rng = np.random.default_rng(11)
    K = spec["K"]
    mean = np.concatenate([stats3.mean, 1.0 - rng.random((K, 1)) * 0.05], axis=1)
    mn = np.concatenate([stats3.min, np.full((K, 1), 0.9)], axis=1)
    mx = np.concatenate([stats3.max, np.full((K, 1), 1.0)], axis=1)
    sd = np.concatenate([stats3.sd, np.full((K, 1), 0.01)], axis=1)
    pal = mr.Palette(mean, mn, mx, sd, np.full(K, 2.0))
    ctx.set_known(None)
    ctx.set_palette(pal.mean, pal.lo, pal.hi, mr._lib.RGB, 4)
    ctx.palette = None
    n1, n2, R = 1001, 140, 2
    rgb = oracle.synth_scan(seed, 2, palrgb, n1, n2)
    alpha = np.full((n2, n1, 1), 255, dtype=np.uint8)
    sp = rng.random((n2, n1)) < transparent
    alpha[sp, 0] = rng.choice(np.array([0, 0, 7, 128, 254], dtype=np.uint8), size=int(sp.sum()))
    scan = np.ascontiguousarray(np.concatenate([rgb, alpha], axis=2))
    want = oracle.categorize_clean(scan, pal.mean, pal.lo, pal.hi, R, channels=4)
    l0 = ctx.launch_count()
    got = ctx.categorize_clean_host(scan, R)
    assert mismatches(got, want) == 0
    assert ctx.launch_count() - l0 <= 6        # strip, categorise, alpha fix, label maximum, vote -- not the generic kernel
    pre = oracle.categorize_u8(scan, pal.mean, pal.lo, pal.hi, channels=4)
    assert mismatches(ctx.categorize_clean_host(scan, 0), pre) == 0
    if transparent > 0:
        assert (pre[sp & (alpha[..., 0] == 0)] == 0).all()
    # device entry point, band with halos
    d = torch.from_numpy(scan).cuda()
    band = torch.empty((60, n1), dtype=torch.uint8, device="cuda")
    ctx.categorize_clean_dev(d[40:], n1, 60, n1 * 4, 4, R, band, n1, halo_lo=R, halo_hi=R)
    torch.cuda.synchronize()
    assert mismatches(band.cpu().numpy(), want[40:100]) == 0


def test_known_category_override(ctx, mr, oracle):
    seed, palrgb, stats, _ = _workload(mr, 1)
    _set(ctx, mr, stats)
    n1, n2 = 200, 150
    scan = oracle.synth_scan(seed, 0, palrgb, n1, n2)
    rng = np.random.default_rng(2)
    idx = rng.choice(n1 * n2, size=300, replace=False).astype(np.int64)
    cat = rng.integers(0, 9, size=300).astype(np.uint8)
    ctx.set_known(idx, cat)
    for R in (0, 2):
        got = ctx.categorize_clean_host(scan, R)
        pre = oracle.categorize_u8(scan, stats.mean, stats.lo, stats.hi, known_idx=idx, known_cat=cat)
        want = pre if R == 0 else oracle.majority(pre, R)
        assert mismatches(got, want) == 0
    ctx.set_known(None)


@pytest.mark.parametrize("R,repeat", [(1, 1), (2, 3), (3, 2), (1, 0), (0, 4)])
def test_clean_standalone_and_repeat(ctx, mr, oracle, R, repeat):
    rng = np.random.default_rng(R * 10 + repeat)
    L = rng.integers(0, 6, size=(90, 141), dtype=np.uint8)
    L[20:60, 30:100] = 3
    want = L.copy()
    for _ in range(repeat if R > 0 else 0):
        want = oracle.majority(want, R)
    assert mismatches(ctx.clean_host(L, R, repeat), want) == 0
    assert mr.clean(L, hood=mr.Window(R), repeat=repeat, ctx=ctx).dtype == np.int64


@pytest.mark.parametrize("maxlab", [1, 14, 15, 30, 31, 62, 63, 126, 127, 254])
def test_clean_standalone_label_ranges(ctx, mr, oracle, maxlab):
    """standalone clean: the largest label picks the plane count of the fast kernel (labels <= 126);
    larger labels take the generic kernel.  Blobby label fields with speckle, windows 3 / 5 / 7."""
    rng = np.random.default_rng(maxlab)
    n2, n1 = 120, 1936
    f = rng.integers(0, maxlab + 1, size=(n2 // 8 + 1, n1 // 8 + 1)).astype(np.uint8)
    lab = np.kron(f, np.ones((8, 8), dtype=np.uint8))[:n2, :n1].copy()
    lab[0, 0] = maxlab
    sp = rng.random((n2, n1)) < 0.05
    lab[sp] = rng.integers(0, maxlab + 1, size=int(sp.sum()))
    d = torch.from_numpy(lab).cuda()
    out = torch.empty_like(d)
    for R in (1, 2, 3):
        ctx.clean_dev(d, n1, n2, n1, R, out, n1)
        torch.cuda.synchronize()
        assert mismatches(out.cpu().numpy(), oracle.majority(lab, R)) == 0
    assert mismatches(ctx.clean_host(lab, 2, repeat=3), oracle.majority(oracle.majority(oracle.majority(lab, 2), 2), 2)) == 0
    with pytest.raises(mr.MapRastError):
        ctx.clean_dev(d, n1, n2, n1, 2, d, n1)       # in place


def test_reference_shaped_api(ctx, mr, oracle):
    """categorize(img, points; tolerances, match) -> Int labels, like _categorize (categories.jl:49-74)."""
    seed, palrgb, _, _ = _workload(mr, 1)
    n1, n2 = 160, 120
    scan = oracle.synth_scan(seed, 0, palrgb, n1, n2)
    rng = np.random.default_rng(4)
    points = [[(float(rng.integers(1, n1 + 1)), float(rng.integers(1, n2 + 1))) for _ in range(12)] for _ in range(4)]
    points.append([])                                       # a category without points: `missing`
    tol = [2.0, 1.0, 3.5, 0.5, 2.0]
    pal = mr.Palette.from_points(scan, points, tol)
    want_pre = oracle.categorize_u8(scan, pal.mean, pal.lo, pal.hi)
    got = mr.categorize(scan, points, tolerances=tol, ctx=ctx)
    assert got.dtype == np.int64 and got.shape == (n2, n1) and mismatches(got, want_pre) == 0
    got2 = mr.categorize_clean(scan, points, tolerances=tol, hood=mr.Window(2), ctx=ctx)
    assert mismatches(got2, oracle.majority(want_pre, 2)) == 0
    # known-category plane from the clicked points (selectcolors.jl:670-681)
    got3 = mr.categorize_u8(scan, points, tolerances=tol, use_known=True, ctx=ctx)
    idx, cat = mr.api._known_from_points(points, n1)
    want3 = oracle.categorize_u8(scan, pal.mean, pal.lo, pal.hi, known_idx=idx, known_cat=cat)
    assert mismatches(got3, want3) == 0


# ---------------------------------------------------------------- device API: bands, batch, synth
def test_map_selection_pipeline(ctx, mr, oracle):
    """MapSelection (src/selectcolors.jl:3-8) -> categorize -> majority clean -> segment prune, the
    button sequence of selectcolors (:466-489) without the GUI, against the oracle stage by stage."""
    from maprast_b200 import synth
    seed, palrgb, stats, spec = _workload(mr, 1)
    n1, n2, K = 640, 400, spec["K"]
    scan = oracle.synth_scan(seed, 0, palrgb, n1, n2)
    rng = np.random.default_rng(3)
    # clicked points: pixels whose colour is close to each palette colour (1-based (i, j))
    points = []
    for c in range(K):
        d = np.abs(scan.astype(np.int32) - palrgb[c].astype(np.int32)).sum(axis=2)
        js, is_ = np.nonzero(d < 12)
        pick = rng.choice(len(js), size=12, replace=False)
        points.append([(float(is_[p] + 1), float(js[p] + 1)) for p in pick])
    tol = np.full(K, 2.5)
    keep = [True] * K
    keep[2] = False
    no_balance = [False] * K      # `category_balance`: no category feeds the detrend -> `_balance` is A .* 1.0
    sel = mr.MapSelection(settings=dict(category_tolerances=tol, category_keep=keep, prune_threshold=30,
                                        line_threshold=0.02, match=("color",), segment=False, blur_repeat=0,
                                        category_balance=no_balance),
                          points=points)
    # no majority filter: a clicked point then carries its own category or 0 (src/categories.jl:110-112),
    # so two different known categories can only meet in the no-match background
    res = mr.rasterize(scan, sel, hood=None, ctx=ctx)
    pal = mr.Palette.from_points(scan, points, tol)
    idx, cat = mr.api._known_from_points(points, n1)
    pre = oracle.categorize_u8(scan, pal.mean, pal.lo, pal.hi, known_idx=idx, known_cat=cat)
    want_cat = pre
    assert mismatches(res["categorized"], want_cat) == 0
    known = np.zeros((n2, n1), dtype=np.uint8)
    known.reshape(-1)[idx] = cat
    want_clean, _ = oracle.clean_segments(want_cat, known, sel.kept_categories(), 0.02, 30)
    assert mismatches(res["cleaned"], want_clean) == 0
    assert not (res["cleaned"] == 3).any() and (res["categorized"] == 3).any()
    with pytest.raises(ValueError):
        mr.rasterize(scan, mr.MapSelection(settings=dict(segment=True), points=points), ctx=ctx)
    with pytest.raises(ValueError):
        mr.rasterize(scan, mr.MapSelection(settings=dict(match=("std", "color")), points=points), ctx=ctx)   # order
    # the "fill" stage (src/selectcolors.jl:490-501) on the pruned labels
    want_fill = oracle.fill_gaps(want_clean, scan, pal.mean, list(sel.kept_categories()), known=known,
                                 repeat=1)
    assert mismatches(res["filled"], want_fill) == 0
    # blur_repeat = 1 (the reference's default): blur -> statistics from the blurred image -> categorise RGB{Float64}
    sel_b = mr.MapSelection(settings=dict(category_tolerances=tol, category_keep=keep, prune_threshold=30,
                                          line_threshold=0.02, match=("color",), segment=False, blur_repeat=1,
                                          category_balance=no_balance),
                            points=points)
    res_b = mr.rasterize(scan, sel_b, hood=None, ctx=ctx)
    blurred = oracle.blur(scan, 1, 1)
    pal_b = mr.Palette.from_points(blurred, points, tol)
    want_b = oracle.categorize_f64_image(blurred, pal_b.mean, pal_b.lo, pal_b.hi, known_idx=idx, known_cat=cat)
    assert mismatches(res_b["categorized"], want_b) == 0
    assert (res_b["categorized"] != res["categorized"]).any()
    # the reference's defaults: `category_balance` = category 1, whose 12 points (>= 8) switch the detrend on
    # (src/balance.jl:16-24), then blur_repeat = 1: balance -> blur -> statistics from that image -> categorise
    sel_d = mr.MapSelection(settings=dict(category_tolerances=tol, category_keep=keep, prune_threshold=30,
                                          line_threshold=0.02, match=("color",), segment=False, blur_repeat=1),
                            points=points)
    res_d = mr.rasterize(scan, sel_d, hood=None, ctx=ctx)
    coef = mr.balance_fit(scan, points)
    assert coef is not None and coef.shape == (3, 7)
    pixels = oracle.blur(oracle.balance(scan, coef), 1, 1)
    pal_d = mr.Palette.from_points(pixels, points, tol)
    want_d = oracle.categorize_f64_image(pixels, pal_d.mean, pal_d.lo, pal_d.hi, known_idx=idx, known_cat=cat)
    assert mismatches(res_d["categorized"], want_d) == 0


def test_device_bands_equal_whole(ctx, mr, oracle):
    """band-sharded launches (with halos) are byte-identical to one launch."""
    seed, palrgb, stats, spec = _workload(mr, 2)
    _set(ctx, mr, stats)
    n1, n2, R = 1000, 403, 2
    scan = oracle.synth_scan(seed, 0, palrgb, n1, n2)
    want = oracle.categorize_clean(scan, stats.mean, stats.lo, stats.hi, R)
    d = torch.from_numpy(scan).cuda()
    whole = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
    ctx.categorize_clean_dev(d, n1, n2, n1 * 3, 3, R, whole, n1)
    banded = torch.empty_like(whole)
    for k in range(5):
        a, b = mr.distributed.band_range(n2, 5, k)
        ctx.categorize_clean_dev(d[a:], n1, b - a, n1 * 3, 3, R, banded[a:], n1,
                                 halo_lo=min(R, a), halo_hi=min(R, n2 - b))
    torch.cuda.synchronize()
    assert mismatches(whole.cpu().numpy(), want) == 0
    assert torch.equal(whole, banded)


def test_bands_with_remote_halo_lines(ctx, mr, oracle):
    """mr_categorize_clean_dev_peer: the halo lines of a band live in OTHER buffers (on the GPU box
    with several GPUs: the neighbour's HBM over NVLink; here: separate allocations on one GPU)."""
    seed, palrgb, stats, spec = _workload(mr, 2)
    _set(ctx, mr, stats)
    n1, n2, R = 1936, 300, 2
    scan = oracle.synth_scan(seed, 1, palrgb, n1, n2)
    want = oracle.categorize_clean(scan, stats.mean, stats.lo, stats.hi, R)
    d = torch.from_numpy(scan).cuda()
    got = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
    cuts = [0, 77, 150, 151 + R, n2]
    for k in range(len(cuts) - 1):
        a, b = cuts[k], cuts[k + 1]
        band = d[a:b].clone()                                   # own lines only
        lo = d[a - R:a].clone() if a > 0 else None              # "neighbour" lines, separate allocations
        hi = d[b:b + R].clone() if b < n2 else None
        ctx.categorize_clean_dev_peer(band, n1, b - a, n1 * 3, 3, R, lo, R if a > 0 else 0, hi, R if b < n2 else 0,
                                      got[a:b], n1)
        torch.cuda.synchronize()
    assert mismatches(got.cpu().numpy(), want) == 0
    with pytest.raises(mr.MapRastError):
        ctx.categorize_clean_dev_peer(d[10:20], n1, 10, n1 * 3, 3, R, None, R, None, 0, got[10:20], n1)
    # IPC export of a device pointer (open needs a second process; exercised by bench.py --gpus N)
    h, off = mr._lib.ipc_export(d[5:])
    assert len(h) == 64 and off >= 5 * n1 * 3


def test_batch_of_sheets(ctx, mr, oracle):
    seed, palrgb, stats, spec = _workload(mr, 5)
    _set(ctx, mr, stats)
    n1 = n2 = 256
    sheets = np.stack([oracle.synth_scan(seed, s, palrgb, n1, n2) for s in range(5)])
    d = torch.from_numpy(sheets).cuda()
    out = torch.empty((5, n2, n1), dtype=torch.uint8, device="cuda")
    ctx.batch_categorize_clean_dev(5, d, n1 * n2 * 3, n1, n2, n1 * 3, 3, 2, out, n1 * n2, n1)
    torch.cuda.synchronize()
    for s in range(5):   # independent sheets: no bleeding across sheet borders
        want = oracle.categorize_clean(sheets[s], stats.mean, stats.lo, stats.hi, 2)
        assert mismatches(out[s].cpu().numpy(), want) == 0


def test_device_synth_matches_oracle_bytes(ctx, mr, oracle):
    seed, palrgb, _, _ = _workload(mr, 2)
    n1, n2 = 700, 300
    for mode in (0, 1):
        d = torch.empty((100, n1, 3), dtype=torch.uint8, device="cuda")
        ctx.synth_scan_dev(seed, 2, palrgb, n1, n2, 150, 100, d, n1 * 3, mode=mode)
        want = oracle.synth_scan(seed, 2, palrgb, n1, n2, line0=150, nlines=100, mode=mode)
        assert np.array_equal(d.cpu().numpy(), want)


def test_argument_errors(ctx, mr):
    _, _, stats, _ = _workload(mr, 1)
    _set(ctx, mr, stats)
    with pytest.raises(mr.MapRastError):
        ctx.categorize_clean_host(np.zeros((4, 4, 3), np.uint8), 9)           # R too large
    with pytest.raises(mr.MapRastError):
        ctx.set_palette(np.full((2, 3), np.inf), np.zeros((2, 3)), np.ones((2, 3)))   # all missing
    with pytest.raises(mr.MapRastError):
        ctx.set_palette(np.zeros((300, 3)), np.zeros((300, 3)), np.ones((300, 3)))    # K > 254
    fresh = mr.Context(0)
    with pytest.raises(mr.MapRastError, match="mr_set_palette"):
        fresh.categorize_clean_host(np.zeros((4, 4, 3), np.uint8), 1)
    fresh.close()
    _set(ctx, mr, stats)
