"""GPU (-m gpu): seeded random geometries for the entry points added in round 2 -- fill_gaps (masks, known
points, category orders, pitches through the host path), blur (u8 / f64, R, repeat), the Float64 categoriser
and the multi-GPU handle (band counts that do not divide the sheet, bands shorter than a block, known points).
Every result is compared with the CPU oracle."""
import numpy as np
import pytest

from util import mismatches, random_palette
from test_oracle_fill import holes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(16))
def test_fill_gaps_random_geometry(ctx, mr, oracle, seed):
    rng = np.random.default_rng(4000 + seed)
    n1 = int(rng.choice([1, 2, 5, 15, 16, 17, 31, 33, 100, 511, 512, 513, 1030]))
    n2 = int(rng.choice([1, 2, 3, 4, 5, 9, 17, 40, 77]))
    K = int(rng.integers(1, 12))
    lab = holes(rng, n2, n1, K, float(rng.choice([0.0, 0.02, 0.3, 0.8, 1.0])))
    bpp = int(rng.choice([3, 4]))
    px = rng.integers(0, 256, size=(n2, n1, bpp), dtype=np.uint8)
    mean = rng.random((K, bpp))
    pal = mr.Palette(mean, mean - 0.05, mean + 0.05, np.full((K, bpp), 0.02))
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    cats = [int(c) for c in rng.permutation(np.arange(1, K + 1))[:int(rng.integers(0, K + 1))]]
    known = np.where(rng.random((n2, n1)) < 0.05, rng.integers(1, K + 1, size=(n2, n1)), 0).astype(np.uint8) if seed % 2 else None
    mask = (rng.random((n2, n1)) < 0.1).astype(np.uint8) if seed % 3 else None
    rep = int(rng.integers(0, 4))
    if known is not None:
        idx = np.flatnonzero(known)
        ctx.set_known(idx, known.reshape(-1)[idx])
    else:
        ctx.set_known(None)
    try:
        got = ctx.fill_gaps_host(lab, px, cats, mask=mask, repeat=rep)
    finally:
        ctx.set_known(None)
    want = oracle.fill_gaps(lab, px, mean, cats, known=known, mask=mask, repeat=rep)
    assert mismatches(got, want) == 0, dict(seed=seed, n1=n1, n2=n2, K=K, cats=cats, rep=rep)


@pytest.mark.parametrize("seed", range(10))
def test_blur_and_float64_categorise_random_geometry(ctx, mr, oracle, seed):
    rng = np.random.default_rng(5000 + seed)
    n1 = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 33, 130, 515]))
    n2 = int(rng.choice([1, 2, 3, 6, 11, 40]))
    R = int(rng.integers(0, 8))
    rep = int(rng.integers(0, 4))
    img = rng.integers(0, 256, size=(n2, n1, 3), dtype=np.uint8) if seed % 2 else rng.random((n2, n1, 3))
    got = ctx.blur_host(img, R, rep)
    want = oracle.blur(img, R, rep)
    assert got.tobytes() == want.tobytes(), dict(seed=seed, n1=n1, n2=n2, R=R, rep=rep)
    K = int(rng.integers(1, 20))
    mean, mn, mx, sd, tol = random_palette(rng, K)
    pal = mr.Palette(mean, mn, mx, sd, tol)
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    ctx.set_known(None)
    Rc = int(rng.integers(0, 5))
    labels = ctx.categorize_clean_f64_host(want, Rc)
    ref = oracle.categorize_f64_image(want, pal.mean, pal.lo, pal.hi)
    assert mismatches(labels, oracle.majority(ref, Rc) if Rc else ref) == 0


@pytest.mark.parametrize("seed", range(8))
def test_multi_random_geometry(mr, oracle, seed):
    from maprast_b200 import synth
    rng = np.random.default_rng(6000 + seed)
    ndev = int(rng.integers(1, 7))
    n1 = int(rng.choice([16, 48, 500, 1000, 1024]))
    n2 = int(rng.choice([1, 7, 16, 17, 33, 100, 211]))
    R = int(rng.integers(0, 4))
    K = int(rng.integers(2, 20))
    pal = synth.make_palette(700 + seed, K)
    st = synth.make_stats(700 + seed, pal)
    scan = oracle.synth_scan(31 + seed, 0, pal, n1, n2)
    m = mr.MultiContext([0] * ndev)
    try:
        m.set_palette(st.mean, st.lo, st.hi)
        if seed % 2:
            idx = np.unique(rng.integers(0, n1 * n2, size=min(60, n1 * n2)))
            cat = rng.integers(1, K + 1, size=idx.size).astype(np.uint8)
            m.set_known(idx, cat)
            lab = oracle.categorize_u8(scan, st.mean, st.lo, st.hi, known_idx=idx, known_cat=cat)
            want = oracle.majority(lab, R) if R else lab
        else:
            want = oracle.categorize_clean(scan, st.mean, st.lo, st.hi, R)
        got = m.categorize_clean_host(scan, R)
        assert mismatches(got, want) == 0, dict(seed=seed, ndev=ndev, n1=n1, n2=n2, R=R, K=K)
    finally:
        m.close()


@pytest.mark.parametrize("seed", range(12))
def test_texture_random_geometry(ctx, mr, oracle, seed):
    """_stds / _stripes planes bit-identical and PixelProp labels identical on seeded random shapes, radii, match
    subsets, window radii and known points (host entry points)"""
    rng = np.random.default_rng(7000 + seed)
    n1 = int(rng.choice([1, 2, 3, 4, 5, 7, 16, 31, 255, 256, 257, 513, 1025]))
    n2 = int(rng.choice([1, 2, 3, 5, 8, 13, 40]))
    K = int(rng.integers(1, 10))
    radius = int(rng.choice([1, 2, 3, 4, 5, 8]))
    px = rng.random((n2, n1, 3))
    px[rng.random((n2, n1)) < 0.1] = rng.random()          # grey pixels
    with np.errstate(invalid="ignore", divide="ignore"):
        std_want, st_want = oracle.stds(px), oracle.stripes(px, radius)
    std, st = ctx.stds_host(px), ctx.stripes_host(px, radius)
    assert std.tobytes() == std_want.tobytes() and st.tobytes() == st_want.tobytes()
    if not (np.isfinite(std).all() and np.isfinite(st).all()):
        return                                              # one-pixel / flat images: NaN planes (0 / 0), as in Julia
    mean, mn, mx, sd, tol = random_palette(rng, K, spread=0.3)
    pal = mr.Palette(mean, mn, mx, sd, tol)
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    sdm, stm = rng.random(K), rng.normal(0, 1, (K, 2))
    sds, sts = (sdm, sdm - rng.random(K), sdm + rng.random(K)), (stm, stm - 2 * rng.random((K, 2)), stm + 2 * rng.random((K, 2)))
    ctx.set_texture_stats(*sds, *sts)
    names = ("color", "std", "stripe")
    bits = int(rng.integers(1, 8))
    match = tuple(n for k, n in enumerate(names) if bits >> k & 1)
    R = int(rng.choice([0, 1, 2, 3]))
    idx = np.unique(rng.integers(0, n1 * n2, size=5))
    cat = rng.integers(1, K + 1, size=idx.size).astype(np.uint8)
    ctx.set_known(idx, cat)
    try:
        got = ctx.categorize_clean_props_host(px, std if bits & 2 else None, st if bits & 4 else None, bits, R)
    finally:
        ctx.set_known(None)
    lab = oracle.categorize_props_image(px, std, st, match, pal.mean, pal.lo, pal.hi, sds, sts, idx, cat)
    assert mismatches(got, oracle.majority(lab, R) if R else lab) == 0
