"""CPU: host-side logic of the reference-interface mirror and the C-ABI surface (no compute)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(mr):
    hdr = open(os.path.join(ROOT, "include", "maprast.h")).read()
    declared = sorted(set(re.findall(r"\b(mr_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = C.CDLL(mr.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"libmaprast.so does not export {name}"
    assert sorted(mr._lib.SYMBOLS) == declared
    assert mr._lib.load().mr_version() >= 100


def test_no_cpu_fallback_without_gpu(mr):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mr.MapRastError, match="no usable CUDA device"):
        mr.Context(0)
    img = np.zeros((4, 4, 3), dtype=np.uint8)
    pal = mr.Palette([[0.5] * 3], [[0.4] * 3], [[0.6] * 3], [[0.1] * 3])
    with pytest.raises(mr.MapRastError):
        mr.categorize(img, pal)


def test_palette_bounds_match_oracle(mr, oracle):
    rng = np.random.default_rng(0)
    mean, mn, mx, sd, tol = __import__("util").random_palette(rng, 9)
    pal = mr.Palette(mean, mn, mx, sd, tol)
    lo, hi = oracle.bounds(mn, mx, sd, tol)
    assert np.array_equal(pal.lo, lo) and np.array_equal(pal.hi, hi)
    assert pal.K == 9 and pal.channels == 3
    assert np.array_equal(mr.Palette(mean, mn, mx, sd).tolerances, np.full(9, 2.0))   # selectcolors.jl:1


def test_component_stats_and_from_points(mr, oracle):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(20, 30, 3), dtype=np.uint8)   # n2=20, n1=30
    # 1-based (i, j) points; 2.5 rounds to 2, 3.5 to 4 (Julia round(Int, x): ties to even)
    pts = [[(2.5, 3.5), (10.2, 7.9), (30.0, 20.0)], [], [(1.0, 1.0)]]
    pal = mr.Palette.from_points(img, pts)
    samples = np.stack([img[4 - 1, 2 - 1], img[8 - 1, 10 - 1], img[20 - 1, 30 - 1]]).astype(np.float64) / 255.0
    for ch in range(3):
        st = oracle.component_stats(samples[:, ch])
        assert np.allclose([pal.mean[0, ch], pal.min[0, ch], pal.max[0, ch], pal.sd[0, ch]], st, rtol=1e-15, atol=0)
    assert np.all(np.isinf(pal.mean[1]))                 # `missing` category
    assert np.all(np.isnan(pal.sd[2])) and np.all(np.isnan(pal.lo[2]))   # one point -> sd NaN
    with pytest.raises(IndexError):
        mr.Palette.from_points(img, [[(31.0, 1.0)]])


def test_unsupported_options_raise(mr):
    img = np.zeros((4, 4, 3), dtype=np.uint8)
    pal = mr.Palette([[0.5] * 3], [[0.4] * 3], [[0.6] * 3], [[0.1] * 3])
    with pytest.raises(ValueError, match="match"):
        mr.categorize(img, pal, match=("color", "std", "stripe"))      # texture matching takes the RGB{Float64} image
    with pytest.raises(ValueError, match="match"):
        mr.categorize(img, pal, match=("std", "color"))                # not the order of _match_from_bools
    with pytest.raises(ValueError, match="match"):
        mr.categorize(img, pal, match=())
    with pytest.raises(ValueError, match="segmented"):
        mr.categorize(img, pal, segmented=True)
    with pytest.raises(ValueError, match="uint8"):
        mr.categorize(img.astype(np.float32), pal)        # float64 RGB is the blurred-image path, float32 is nothing
    with pytest.raises(ValueError):
        mr.clean(np.zeros((3, 3)))
    with pytest.raises(ValueError):
        mr.Window(-1)
    with pytest.raises(ValueError):
        mr.Palette([[0.5] * 3], [[0.4] * 3], [[0.6] * 3], [[0.1] * 3], tolerances=[1.0, 2.0])


def test_workload_definitions(mr):
    from maprast_b200 import synth
    for cfg, spec in synth.CONFIGS.items():
        seed, pal, stats, spec = synth.config_workload(cfg)
        assert seed == 0x4D520000 + cfg and pal.shape == (spec["K"], 3) and stats.K == spec["K"]
        d = ((pal[:, None].astype(int) - pal[None].astype(int)) ** 2).sum(-1)
        np.fill_diagonal(d, 1 << 30)
        assert np.sqrt(d.min()) >= 24            # min pairwise distance >= 24/255
        assert np.all(np.isfinite(stats.lo)) and np.all(stats.lo < stats.hi)
    lat = synth.make_palette(7, 16, "lattice")
    assert set(np.unique(lat)) <= {32, 96, 160, 224}
    st = synth.make_stats(7, lat, exact_means=True)
    assert np.array_equal(st.mean, lat / 255.0)


def test_balance_fit_host_part(mr):
    """host part of `_balance` (src/balance.jl:1-24): fewer than 8 points -> None (A .* 1.0); otherwise the
    least-squares cubic through the clicked points' colour deviations from their category mean"""
    rng = np.random.default_rng(5)
    n2, n1 = 200, 300
    img = rng.integers(0, 256, size=(n2, n1, 3), dtype=np.uint8)
    pts = [(float(rng.integers(1, n1 + 1)), float(rng.integers(1, n2 + 1))) for _ in range(7)]
    assert mr.balance_fit(img, [pts, []]) is None
    assert mr.balance_fit(img, [pts, pts], category_bitindex=[True, True]).shape == (3, 7)     # 14 points
    assert mr.balance_fit(img, [pts, pts * 2]) is None                                             # default: category 1 only
    # a sheet whose red channel is (mean + a cubic in i): the fit returns that cubic
    ii = np.arange(1, n1 + 1, dtype=np.float64)
    trend = 2e-8 * ii ** 3 - 6e-6 * ii ** 2 + 4e-4 * ii
    sheet = np.zeros((n2, n1, 3))
    sheet[:, :, 0] = 0.3 + trend[None, :]
    sheet[:, :, 1:] = 0.5
    img = np.rint(sheet * 255).astype(np.uint8)
    pts = [(float(i), float(j)) for i in range(10, n1, 29) for j in (20, 90, 160)]
    coef = mr.balance_fit(img, [pts])
    xs = np.array([p[0] for p in pts])
    got = coef[0, 0] + coef[0, 1] * xs ** 3 + coef[0, 2] * xs ** 2 + coef[0, 3] * xs
    want = img[0, (xs - 1).astype(int), 0] / 255.0
    assert np.abs((got - got.mean()) - (want - want.mean())).max() < 3e-3      # up to N0f8 quantisation
    assert np.abs(coef[1:]).max() < 1e-6                                         # flat channels: no trend


def test_texture_stats_from_points(mr, oracle):
    """`std` / `stripe` entries of `_component_stats(::Vector{<:PixelProp})` (src/categories.jl:10-17) at the clicked
    points, bounds as in :131-133; a plane that is None is the reference's initial all-zero plane"""
    rng = np.random.default_rng(2)
    n2, n1 = 12, 17
    std, stripe = rng.random((n2, n1)), rng.normal(0, 1, (n2, n1, 2))
    pts = [[(2.5, 3.5), (10.2, 7.9), (17.0, 12.0)], [], [(1.0, 1.0)]]
    tol = np.array([2.0, 1.0, 3.0])
    tex = mr.TextureStats.from_points(std, stripe, pts, tol)
    ij = [(4 - 1, 2 - 1), (8 - 1, 10 - 1), (12 - 1, 17 - 1)]                 # (j, i) 0-based; 2.5 -> 2, 3.5 -> 4
    want = oracle.component_stats(np.array([std[j, i] for j, i in ij]))
    assert np.allclose([tex.std[k][0] for k in ("mean", "min", "max", "sd")], want, rtol=1e-15, atol=0)
    for e in range(2):
        want = oracle.component_stats(np.array([stripe[j, i, e] for j, i in ij]))
        assert np.allclose([tex.stripe[k][0, e] for k in ("mean", "min", "max", "sd")], want, rtol=1e-15, atol=0)
    assert np.isnan(tex.std["mean"][1]) and np.isnan(tex.stripe["mean"][1]).all()      # `missing` category
    assert np.isnan(tex.std["sd"][2])                                                     # one point -> sd NaN
    (slo, shi), (tlo, thi) = tex.bounds()
    assert slo[0] == tex.std["min"][0] - tex.std["sd"][0] * 2.0 and shi[0] == tex.std["max"][0] + tex.std["sd"][0] * 2.0
    assert np.array_equal(tlo[0], tex.stripe["min"][0] - tex.stripe["sd"][0] * 2.0)
    zero = mr.TextureStats.from_points(None, None, pts, tol)
    assert zero.std["mean"][0] == 0.0 and (zero.stripe["max"][0] == 0.0).all()


def test_polygon_and_morph_host_logic(mr):
    """host side of the polygon / erode / dilate mirrors: ring flattening, category of a ring, argument checks
    (nothing here needs a GPU: the checks come before the context is created)"""
    polys = [[[(100.0, 100.0)]], [[(1, 1), (5, 1), (3, 4)], [(2, 2), (3, 2), (3, 3), (2, 3)]], []]
    rings, cats = mr.api._rings_of(polys)
    assert len(rings) == 3 and cats == [1, 2, 2]
    assert mr.api._rings_of(None) == ([], [])
    with pytest.raises(ValueError):
        mr.paint_polygons(np.zeros((3, 3)), polys)                 # float labels
    with pytest.raises(ValueError):
        mr.paint_polygons(np.full((3, 3), 255), polys)             # label 255
    with pytest.raises(ValueError):
        mr.shrink_grow(np.zeros((3, 3), dtype=np.int64), -1, 0)
    with pytest.raises(ValueError):
        mr.shrink_grow(np.zeros((3, 3), dtype=np.int64), 1, 1, missingval=7)
    with pytest.raises(ValueError):
        mr.recolor(np.zeros((3, 3), dtype=np.int64), np.zeros((3, 3, 3)), [])
    assert (mr.mean_color(np.full((4, 4, 3), 0.25), []) == 0).all()             # no points: the zero colour
    assert np.allclose(mr.mean_color(np.full((4, 4, 3), 0.25), [(1.0, 1.0), (2.0, 3.0)]), 0.25)
    sel = mr.MapSelection(settings={}, points=[[]], polygons=polys)
    assert sel.polygons is polys
