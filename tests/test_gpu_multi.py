"""GPU (-m gpu): the multi-GPU C ABI (mr_multi_*).  The test box has one GPU, so the
handle is built over the SAME device listed several times (one context per entry): the band split, the
halo handling, the per-band known-category translation and the host threads are exactly those of a real
multi-GPU box.  Results must be byte-identical to the single-context call and to the oracle."""
import numpy as np
import pytest
import torch

from util import mismatches, random_palette

pytestmark = pytest.mark.gpu


def _setup(mr, oracle, K, n1, n2, seed):
    from maprast_b200 import synth
    rng = np.random.default_rng(seed)
    palrgb = synth.make_palette(seed, K) if hasattr(synth, "make_palette") else rng.integers(0, 256, size=(K, 3), dtype=np.uint8)
    scan = oracle.synth_scan(seed, 0, palrgb, n1, n2)
    mean = palrgb.astype(np.float64) / 255.0
    lo, hi = mean - 0.08, mean + 0.08
    return scan, mean, lo, hi


@pytest.mark.parametrize("ndev", [1, 2, 3, 5])
@pytest.mark.parametrize("R", [0, 1, 2, 3])
def test_host_sheet_bands(mr, oracle, ndev, R):
    n1, n2, K = 1000, 333, 12
    scan, mean, lo, hi = _setup(mr, oracle, K, n1, n2, 11)
    m = mr.MultiContext([0] * ndev)
    try:
        assert m.n_dev == ndev
        ranges = [m.band_range(n2, i) for i in range(ndev)]
        assert ranges[0][0] == 0 and sum(r[1] for r in ranges) == n2
        m.set_palette(mean, lo, hi)
        got, cnt = m.categorize_clean_host(scan, R, counters=True)
        want = oracle.categorize_clean(scan, mean, lo, hi, R)
        assert mismatches(got, want) == 0
        assert cnt["n_pixels"] == n1 * n2 and cnt["n_nomatch"] == int((want == 0).sum())
    finally:
        m.close()


def test_known_points_on_band_seams(mr, oracle, ctx):
    n1, n2, K, R = 512, 160, 8, 2
    scan, mean, lo, hi = _setup(mr, oracle, K, n1, n2, 12)
    rng = np.random.default_rng(5)
    m = mr.MultiContext([0, 0, 0])
    try:
        m.set_palette(mean, lo, hi)
        seams = [m.band_range(n2, i)[0] for i in range(1, 3)]
        js = np.concatenate([np.arange(s - 3, s + 3) for s in seams] + [rng.integers(0, n2, size=40)])
        js = np.repeat(js, 6)
        ii = rng.integers(0, n1, size=js.size)
        idx = np.unique(ii + js * n1)
        cat = rng.integers(1, K + 1, size=idx.size).astype(np.uint8)
        m.set_known(idx, cat)
        got = m.categorize_clean_host(scan, R)
        # single context, whole image
        ctx.set_palette(mean, lo, hi)
        ctx.palette = None
        ctx.set_known(idx, cat)
        try:
            one = ctx.categorize_clean_host(scan, R)
        finally:
            ctx.set_known(None)
        assert mismatches(got, one) == 0
        lab = oracle.categorize_u8(scan, mean, lo, hi, known_idx=idx, known_cat=cat)
        assert mismatches(got, oracle.majority(lab, R)) == 0
        m.set_known(None)
        assert mismatches(m.categorize_clean_host(scan, R), oracle.categorize_clean(scan, mean, lo, hi, R)) == 0
    finally:
        m.close()


def test_batch_of_sheets(mr, oracle):
    n1, n2, K, R = 256, 96, 6, 1
    from maprast_b200 import synth
    palrgb = np.random.default_rng(3).integers(0, 256, size=(K, 3), dtype=np.uint8)
    sheets = np.stack([oracle.synth_scan(77, s, palrgb, n1, n2) for s in range(5)])
    mean = palrgb.astype(np.float64) / 255.0
    m = mr.MultiContext([0, 0])
    try:
        m.set_palette(mean, mean - 0.07, mean + 0.07)
        got = m.batch_categorize_clean_host(sheets, R)
        for s in range(5):
            assert mismatches(got[s], oracle.categorize_clean(sheets[s], mean, mean - 0.07, mean + 0.07, R)) == 0
    finally:
        m.close()


@pytest.mark.parametrize("ndev", [2, 3])
def test_device_resident_bands_peer_halos(mr, oracle, ndev):
    n1, n2, K, R = 1024, 208, 10, 2
    scan, mean, lo, hi = _setup(mr, oracle, K, n1, n2, 13)
    dev = torch.device("cuda", 0)
    m = mr.MultiContext([0] * ndev)
    try:
        m.set_palette(mean, lo, hi)
        bands, outs = [], []
        for i in range(ndev):
            a, n = m.band_range(n2, i)
            bands.append(torch.from_numpy(scan[a:a + n].copy()).to(dev))     # separate allocation per band
            outs.append(torch.zeros((n, n1), dtype=torch.uint8, device=dev))
        torch.cuda.synchronize()
        rng = np.random.default_rng(9)
        for known in (False, True):
            if known:
                idx = np.unique(rng.integers(0, n1 * n2, size=300))
                cat = rng.integers(1, K + 1, size=idx.size).astype(np.uint8)
                m.set_known(idx, cat)
                want = oracle.majority(oracle.categorize_u8(scan, mean, lo, hi, known_idx=idx, known_cat=cat), R)
            else:
                want = oracle.categorize_clean(scan, mean, lo, hi, R)
            m.categorize_clean_dev(bands, n1, n2, n1 * 3, 3, R, outs, n1)
            got = np.concatenate([o.cpu().numpy() for o in outs])
            assert mismatches(got, want) == 0
    finally:
        m.close()


def test_errors(mr):
    with pytest.raises(mr.MapRastError):
        mr.MultiContext([99])
    m = mr.MultiContext([0, 0])
    try:
        with pytest.raises(mr.MapRastError):     # no palette yet
            m.categorize_clean_host(np.zeros((8, 16, 3), np.uint8), 1)
    finally:
        m.close()
