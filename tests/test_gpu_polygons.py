"""GPU (-m gpu): polygon masks and polygon painting (mr_polygons_*; /root/reference/src/selectcolors.jl:480-481,
491-497, 559-563) against the oracle, and the button sequence with drawn polygons through `rasterize`."""
import numpy as np
import pytest
import torch

from util import mismatches
from test_oracle_polygons import random_rings

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(10))
def test_polygons_random(ctx, oracle, seed):
    rng = np.random.default_rng(100 + seed)
    n1 = int(rng.choice([1, 7, 64, 255, 256, 257, 700]))
    n2 = int(rng.choice([1, 5, 33, 120]))
    rings = random_rings(rng, n1, n2, int(rng.integers(0, 7)))
    rings.append([(100.0, 100.0)])
    rings.append([(2.0, 2.0), (8.0, 2.0), (8.0, 6.0), (2.0, 6.0)])
    assert mismatches(ctx.polygons_host(rings, (n2, n1)), oracle.polygons(rings, (n2, n1))) == 0
    plane = rng.integers(0, 9, size=(n2, n1), dtype=np.uint8)
    fill = rng.integers(0, 9, size=len(rings)).astype(np.uint8)
    assert mismatches(ctx.polygons_host(rings, (n2, n1), fill=fill, plane=plane), oracle.polygons(rings, (n2, n1), fill, plane)) == 0
    assert ctx.polygons_host([], (n2, n1)).sum() == 0


def test_polygon_with_thousands_of_vertices(ctx, oracle):
    """more edges than the shared-memory crossing lists hold: the per-pixel kernel"""
    n1, n2 = 130, 90
    ang = np.linspace(0, 2 * np.pi, 6000, endpoint=False)
    ring = np.stack([65 + 50 * np.cos(ang) * (1 + 0.2 * np.sin(7 * ang)), 45 + 35 * np.sin(ang)], axis=1).tolist()
    rings = [ring, [(10.0, 10.0), (40.0, 12.0), (20.0, 50.0)]]
    got = ctx.polygons_host(rings, (n2, n1))
    assert mismatches(got, oracle.polygons(rings, (n2, n1))) == 0 and 1000 < got.sum() < n1 * n2


def test_polygons_device_entry_with_pitch(ctx, oracle):
    rng = np.random.default_rng(5)
    n1, n2 = 301, 77
    rings = random_rings(rng, n1, n2, 5)
    off, xy = oracle.flatten_rings(rings)
    fill = np.arange(1, 6, dtype=np.uint8)
    plane = rng.integers(0, 9, size=(n2, n1), dtype=np.uint8)
    d = torch.full((n2, n1 + 19), 77, dtype=torch.uint8, device="cuda")
    d[:, :n1] = torch.from_numpy(plane).cuda()
    ctx._ck(ctx._L.mr_polygons_dev(ctx._h, len(rings), off.ctypes.data, xy.ctypes.data, fill.ctypes.data, n1, n2, d.data_ptr(),
                                   n1 + 19, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert mismatches(d[:, :n1].cpu().numpy(), oracle.polygons(rings, (n2, n1), fill, plane)) == 0
    assert bool((d[:, n1:] == 77).all())
    with pytest.raises(Exception, match="finite"):
        ctx.polygons_host([[(1.0, 1.0), (np.inf, 2.0), (3.0, 3.0)]], (n2, n1))


def test_rasterize_with_drawn_polygons(ctx, mr, oracle):
    """"clean" masks the categorised labels with the polygons, "fill" leaves masked gaps alone and then paints every
    polygon with its category (src/selectcolors.jl:479-501)"""
    from maprast_b200 import synth
    seed, pal_rgb, stats, spec = synth.config_workload(1)
    n1, n2 = 320, 200
    img = oracle.synth_scan(seed, 0, pal_rgb, n1, n2)
    rng = np.random.default_rng(8)
    K = 4
    points = [[(float(rng.integers(3, n1 - 2)), float(rng.integers(3, n2 - 2))) for _ in range(6)] for _ in range(K)]
    polygons = [[[(100.0, 100.0)]], [[(20.0, 30.0), (120.0, 25.0), (90.0, 90.0)]], [], [[(200.5, 50.5), (300.0, 60.0), (280.0, 150.0), (210.0, 140.0)]]]
    tol = np.full(K, 3.0)
    sel = mr.MapSelection(settings={"match": ("color",), "blur_repeat": 0, "category_tolerances": tol,
                                    "category_balance": [False] * K, "prune_threshold": 5, "line_threshold": 0.0,
                                    "fill_repeat": 2}, points=points, polygons=polygons)
    res = mr.rasterize(img, sel, ctx=ctx)
    pal = mr.Palette.from_points(img, points, tol)
    kidx, kcat = mr.api._known_from_points(points, n1)
    cat = oracle.categorize_u8(img, pal.mean, pal.lo, pal.hi, known_idx=kidx, known_cat=kcat)
    assert mismatches(res["categorized"], cat) == 0
    rings, rc = mr.api._rings_of(polygons)
    mask = oracle.polygons(rings, (n2, n1))
    assert mask.sum() > 1000
    known = np.zeros((n2, n1), dtype=np.uint8)
    known.reshape(-1)[kidx] = kcat
    cleaned, _ = oracle.clean_segments(cat * (1 - mask), known, range(1, K + 1), 0.0, 5)
    assert mismatches(res["cleaned"], cleaned) == 0
    filled = oracle.fill_gaps(cleaned, img, pal.mean, list(range(1, K + 1)), known=known, mask=mask, repeat=2)
    filled = oracle.polygons(rings, (n2, n1), np.asarray(rc, dtype=np.uint8), filled)
    assert mismatches(res["filled"], filled) == 0
    assert (res["filled"][mask != 0] != 0).all()
