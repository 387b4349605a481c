"""GPU (-m gpu): segment prune (`_clean`, /root/reference/src/clean.jl:59-93) through the C ABI
against the CPU oracle's literal restatement of the sequential algorithm.  Bit-exact labels."""
import numpy as np
import pytest
import torch

from util import mismatches
from test_oracle_segments import blobs

pytestmark = pytest.mark.gpu


def _known_lists(known):
    idx = np.flatnonzero(known)
    return idx, known.reshape(-1)[idx].astype(np.uint8)


@pytest.mark.parametrize("shape", [(1, 1), (1, 77), (90, 1), (33, 32), (64, 100), (301, 257), (512, 1000)])
@pytest.mark.parametrize("seed", [0, 1])
def test_random_fields(ctx, oracle, shape, seed):
    rng = np.random.default_rng(100 + seed)
    n2, n1 = shape
    K = 7
    lab = blobs(rng, n2, n1, K, 0.05)
    ctx.set_known(None)
    for cats, lt, pt in [(range(1, K + 1), 0.01, 20), ((1, 3, 5, 7), 0.3, 6), ((2,), 0.0, 0), (range(0, K + 1), 0.02, 50)]:
        keep = oracle.keep_table(cats)
        got, nseg = ctx.clean_segments_host(lab, keep, lt, pt)
        want, nseg_o = oracle.clean_segments(lab, None, cats, lt, pt)
        assert mismatches(got, want) == 0
        assert nseg == nseg_o


def test_worst_case_shapes(ctx, oracle):
    """checkerboard (every pixel its own segment), one giant segment, spirals / combs (long merge chains)"""
    ctx.set_known(None)
    n2, n1 = 200, 333
    yy, xx = np.mgrid[0:n2, 0:n1]
    cases = [((yy + xx) % 2 + 1).astype(np.uint8), np.full((n2, n1), 3, dtype=np.uint8)]
    comb = np.full((n2, n1), 1, dtype=np.uint8)
    comb[1:, ::2] = 2                     # teeth joined only through line 0
    cases.append(comb)
    comb2 = np.full((n2, n1), 1, dtype=np.uint8)
    comb2[:-1, ::2] = 2                   # teeth joined only through the last line
    cases.append(comb2)
    spiral = np.zeros((n2, n1), dtype=np.uint8)
    for k in range(0, min(n2, n1) // 2, 2):
        spiral[k, k:n1 - k] = 4
        spiral[k:n2 - k, n1 - k - 1] = 4
        spiral[n2 - k - 1, k:n1 - k] = 4
        spiral[k + 2:n2 - k, k] = 4
    cases.append(spiral)
    for lab in cases:
        keep = oracle.keep_table(range(1, 5))
        for lt, pt in [(0.01, 20), (0.0, 0), (0.6, 3)]:
            got, nseg = ctx.clean_segments_host(lab, keep, lt, pt)
            want, nseg_o = oracle.clean_segments(lab, None, range(1, 5), lt, pt)
            assert mismatches(got, want) == 0 and nseg == nseg_o


def test_known_categories_and_conflict(ctx, mr, oracle):
    rng = np.random.default_rng(5)
    lab = blobs(rng, 160, 200, 5, 0.06)
    known = np.zeros_like(lab)
    # one known point on each of many small segments (isolated speckle)
    ys, xs = np.nonzero((lab != np.roll(lab, 1, 0)) & (lab != np.roll(lab, -1, 0)) & (lab != np.roll(lab, 1, 1))
                        & (lab != np.roll(lab, -1, 1)))
    for y, x in list(zip(ys, xs))[:60]:
        if 0 < y < 159 and 0 < x < 199:
            known[y, x] = 1 + (int(lab[y, x]) + 1) % 5
    assert known.any()
    cats = range(1, 6)
    ctx.set_known(*_known_lists(known))
    got, _ = ctx.clean_segments_host(lab, oracle.keep_table(cats), 0.01, 20)
    want, _ = oracle.clean_segments(lab, known, cats, 0.01, 20)
    assert mismatches(got, want) == 0
    assert (got != oracle.clean_segments(lab, None, cats, 0.01, 20)[0]).any()
    # two different known categories inside one segment: scan-order dependent in the reference -> refused
    lab2 = np.full((4, 6), 4, dtype=np.uint8)
    kn2 = np.zeros_like(lab2)
    kn2[0, 1], kn2[0, 4] = 1, 2
    ctx.set_known(*_known_lists(kn2))
    with pytest.raises(mr.MapRastError) as ei:
        ctx.clean_segments_host(lab2, oracle.keep_table((1, 2, 4)), 0.0, 100)
    assert ei.value.code == -5
    # ... but not when the segment's label is not kept (e.g. the no-match background holding clicked
    # points that did not match): every part of it becomes 0 however the reference splits it
    got2, _ = ctx.clean_segments_host(lab2, oracle.keep_table((1, 2)), 0.0, 100)
    want2, _ = oracle.clean_segments(lab2, kn2, (1, 2), 0.0, 100)
    assert mismatches(got2, want2) == 0 and not got2.any()
    ctx.set_known(None)


def test_api_mirror_and_device_entry(ctx, mr, oracle):
    rng = np.random.default_rng(9)
    lab = blobs(rng, 257, 384, 6, 0.04)
    out = mr.clean_segments(lab.astype(np.int64), categories=(1, 2, 3, 4, 5, 6), line_threshold=0.05,
                            prune_threshold=12, ctx=ctx)
    want, _ = oracle.clean_segments(lab, None, (1, 2, 3, 4, 5, 6), 0.05, 12)
    assert out.dtype == np.int64 and mismatches(out, want) == 0
    with pytest.raises(ValueError):
        mr.clean_segments(lab.astype(np.float64), categories=(1,), ctx=ctx)
    # device pointers with padded strides
    d_in = torch.zeros((257, 400), dtype=torch.uint8, device="cuda")
    d_in[:, :384] = torch.from_numpy(lab).cuda()
    d_out = torch.full((257, 416), 99, dtype=torch.uint8, device="cuda")
    ctx.set_known(None)
    nseg = ctx.clean_segments_dev(d_in, 384, 257, 400, oracle.keep_table(range(1, 7)), 0.05, 12, d_out, 416)
    torch.cuda.synchronize()
    assert mismatches(d_out[:, :384].cpu().numpy(), want) == 0 and nseg > 0
    assert bool((d_out[:, 384:] == 99).all())


def test_fused_then_segments_on_synthetic_sheet(ctx, mr, oracle):
    """the reference pipeline order: categorise (+ majority clean) -> segment prune, 2048 x 1536 crop of config 2"""
    from maprast_b200 import synth
    seed, palrgb, stats, spec = synth.config_workload(2)
    n1, n2 = 2048, 1536
    ctx.set_known(None)
    ctx.set_palette(stats.mean, stats.lo, stats.hi)
    scan = oracle.synth_scan(seed, 0, palrgb, n1, n2)
    labels = ctx.categorize_clean_host(scan, spec["R"])
    keep = oracle.keep_table(range(1, spec["K"] + 1))
    got, nseg = ctx.clean_segments_host(labels, keep, 0.01, 20)
    want, nseg_o = oracle.clean_segments(labels, None, range(1, spec["K"] + 1), 0.01, 20)
    assert mismatches(got, want) == 0 and nseg == nseg_o


@pytest.mark.parametrize("R", [0, 2])
def test_prune_at_sheet_width_against_the_sequential_oracle(ctx, mr, oracle, R):
    """the segment prune at BASELINE config 2's line width (20000 pixels: five 4096-pixel chunks per line, the last one
    ragged) on raw categorise-only labels (R = 0: about one speckle segment per 40 pixels) and on filtered labels:
    labels and segment count identical to the literal sequential restatement"""
    import torch
    from maprast_b200 import synth
    seed, palrgb, stats, spec = synth.config_workload(2)
    n1, n2, K = spec["n1"], 1200, spec["K"]
    ctx.set_palette(stats.mean, stats.lo, stats.hi)
    ctx.palette = None
    ctx.set_known(None)
    px = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda")
    ctx.synth_scan_dev(seed, 0, palrgb, n1, spec["n2"], 0, n2, px, n1 * 3)
    lab = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
    ctx.categorize_clean_dev(px, n1, n2, n1 * 3, 3, R, lab, n1)
    keep = np.zeros(256, dtype=np.uint8)
    keep[1:K + 1] = 1
    out = torch.empty_like(lab)
    nseg = ctx.clean_segments_dev(lab, n1, n2, n1, keep, 0.01, 20, out, n1)
    torch.cuda.synchronize()
    want, want_seg = oracle.clean_segments(lab.cpu().numpy(), None, range(1, K + 1), 0.01, 20)
    assert nseg == want_seg and nseg > (100000 if R == 0 else 1000)
    assert mismatches(out.cpu().numpy(), want) == 0
