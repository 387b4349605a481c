"""GPU (-m gpu): BASELINE.json full-size checks through size-independent properties."""
import numpy as np
import pytest
import torch

from util import mismatches

pytestmark = pytest.mark.gpu


def _crop_check(ctx, oracle, stats, d_scan, d_out, n1, n2, R, boxes, cs):
    """oracle on crops (with R context) of a device-resident sheet vs the device labels."""
    for (i0, i1, j0, j1) in boxes:
        a0, a1, b0, b1 = max(i0 - R, 0), min(i1 + R, n1), max(j0 - R, 0), min(j1 + R, n2)
        crop = d_scan[b0:b1, a0:a1].contiguous().cpu().numpy()
        pre = oracle.categorize_u8(crop, stats.mean, stats.lo, stats.hi, cs)
        # the oracle clips at the crop edge; only compare where the window lies inside the crop or
        # the crop edge coincides with the sheet edge
        want = oracle.majority(pre, R)
        got = d_out[b0:b1, a0:a1].cpu().numpy()
        ys = slice(j0 - b0, (j1 - b0))
        xs = slice(i0 - a0, (i1 - a0))
        assert mismatches(got[ys, xs], want[ys, xs]) == 0


@pytest.mark.parametrize("cfg", [2, 3, 4])
def test_full_size_sheet(ctx, mr, oracle, cfg):
    from maprast_b200 import synth
    seed, palrgb, stats, spec = synth.config_workload(cfg)
    n1, n2, R = spec["n1"], spec["n2"], spec["R"]
    cs = oracle.LAB if stats.colourspace == "lab" else oracle.RGB
    ctx.set_known(None)
    ctx.set_palette(stats.mean, stats.lo, stats.hi, mr._lib.LAB if cs == oracle.LAB else mr._lib.RGB)
    d = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda")
    ctx.synth_scan_dev(seed, 0, palrgb, n1, n2, 0, n2, d, n1 * 3)
    out = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
    ctx.enable_counters(True)
    ctx.read_counters(reset=True)
    ctx.categorize_clean_dev(d, n1, n2, n1 * 3, 3, R, out, n1)
    cnt = ctx.read_counters(reset=True)
    ctx.enable_counters(False)
    assert cnt["n_pixels"] == n1 * n2
    # (1) fused == categorise-only followed by the standalone clean kernel (different code paths)
    pre = torch.empty_like(out)
    ctx.categorize_clean_dev(d, n1, n2, n1 * 3, 3, 0, pre, n1)
    two = torch.empty_like(out)
    ctx.clean_dev(pre, n1, n2, n1, R, two, n1)
    torch.cuda.synchronize()
    assert torch.equal(out, two)
    assert cnt["n_nomatch"] == int((out == 0).sum().item())
    assert cnt["n_changed"] == int((out != pre).sum().item())
    # (2) 8 bands with halos == one launch
    banded = torch.empty_like(out)
    for k in range(8):
        a, b = mr.distributed.band_range(n2, 8, k)
        ctx.categorize_clean_dev(d[a:], n1, b - a, n1 * 3, 3, R, banded[a:], n1, halo_lo=min(R, a),
                                 halo_hi=min(R, n2 - b))
    torch.cuda.synchronize()
    assert torch.equal(out, banded)
    del banded, two
    # (3) oracle on crops: four corners, edges, band seams, interior
    s = 192
    seam = mr.distributed.band_range(n2, 8, 3)[1]
    boxes = [(0, s, 0, s), (n1 - s, n1, 0, s), (0, s, n2 - s, n2), (n1 - s, n1, n2 - s, n2),
             (n1 // 2, n1 // 2 + s, seam - s // 2, seam + s // 2), (7777, 7777 + s, 9999, 9999 + s),
             (n1 - 3 * s, n1 - 2 * s, n2 - 3 * s, n2 - 2 * s),
             (n1 - s, n1, seam - 40, seam + 40)]
    _crop_check(ctx, oracle, stats, d, out, n1, n2, R, boxes, cs)
    # (4) clean is idempotent on constant regions and label histogram is plausible
    hist = torch.bincount(out.flatten().to(torch.int64), minlength=spec["K"] + 1)
    assert int(hist.sum().item()) == n1 * n2 and int(hist[spec["K"] + 1:].sum().item()) == 0
    assert (hist[1:spec["K"] + 1] > 0).all()


def test_widened_stages_at_sheet_width(ctx, mr, oracle):
    """the stages around the hot path at BASELINE config 2's line width (20000 pixels, a few lines): texture planes bit
    for bit, PixelProp labels, polygon mask, erode / dilate, display layer -- all against the oracle"""
    rng = np.random.default_rng(77)
    n1, n2, K = 20000, 7, 6
    px = rng.random((n2, n1, 3))
    with np.errstate(invalid="ignore", divide="ignore"):
        std_w, st_w = oracle.stds(px), oracle.stripes(px, 2)
    std, st = ctx.stds_host(px), ctx.stripes_host(px, 2)
    assert std.tobytes() == std_w.tobytes() and st.tobytes() == st_w.tobytes()
    mean = rng.random((K, 3))
    pal = mr.Palette(mean, mean - 0.3, mean + 0.3, np.zeros((K, 3)))
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    ctx.set_known(None)
    sdm, stm = np.full(K, float(std.mean())), np.zeros((K, 2))
    sds, sts = (sdm, sdm - 1, sdm + 1), (stm, stm - 50, stm + 50)
    ctx.set_texture_stats(*sds, *sts)
    got = ctx.categorize_clean_props_host(px, std, st, 7, 0)
    want = oracle.categorize_props_image(px, std, st, ("color", "std", "stripe"), pal.mean, pal.lo, pal.hi, sds, sts)
    assert int((got != want).sum()) == 0 and len(np.unique(got)) > 3
    rings = [[(10.0, 1.0), (19990.5, 2.0), (15000.0, 7.5), (300.0, 6.0)], [(5000.0, 0.0), (5100.0, 8.0), (4900.0, 8.0)]]
    assert int((ctx.polygons_host(rings, (n2, n1)) != oracle.polygons(rings, (n2, n1))).sum()) == 0
    assert int((ctx.shrink_grow_host(got, 1, 2) != oracle.shrink_grow(got, 1, 2)).sum()) == 0
    assert ctx.recolor_host(got, mean).tobytes() == oracle.recolor(got, mean).tobytes()
