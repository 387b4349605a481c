"""GPU (-m gpu): fill_gaps (/root/reference/src/clean.jl:2-54, `_fill` src/selectcolors.jl:683-699) through the
C ABI against the CPU oracle.  Bit-exact labels, including the Float64 weight sums and colour errors."""
import numpy as np
import pytest
import torch

from util import mismatches
from test_oracle_fill import holes

pytestmark = pytest.mark.gpu


def _palette(ctx, mr, mean, colourspace="rgb"):
    K, C = mean.shape
    pal = mr.Palette(mean, mean - 0.05, mean + 0.05, np.full((K, C), 0.02), colourspace=colourspace)
    ctx.set_palette(pal.mean, pal.lo, pal.hi, mr._lib.LAB if colourspace == "lab" else mr._lib.RGB)
    ctx.palette = None
    return pal


def _known_lists(known):
    idx = np.flatnonzero(known)
    return idx, known.reshape(-1)[idx].astype(np.uint8)


@pytest.mark.parametrize("shape", [(1, 1), (1, 77), (90, 1), (33, 32), (64, 100), (301, 257), (256, 1024)])
@pytest.mark.parametrize("seed", [0, 1])
def test_random_fields(ctx, mr, oracle, shape, seed):
    rng = np.random.default_rng(700 + seed)
    n2, n1 = shape
    K = 9
    lab = holes(rng, n2, n1, K, 0.2 if seed == 0 else 0.5)
    px = rng.integers(0, 256, size=(n2, n1, 3), dtype=np.uint8)
    mean = rng.random((K, 3))
    _palette(ctx, mr, mean)
    ctx.set_known(None)
    for cats in (list(range(1, K + 1)), [7, 2, 9, 4], [5], []):
        for rep in (1, 2, 3):
            got = ctx.fill_gaps_host(lab, px, cats, repeat=rep)
            want = oracle.fill_gaps(lab, px, mean, cats, repeat=rep)
            assert mismatches(got, want) == 0
    assert mismatches(ctx.fill_gaps_host(lab, px, [1, 2], repeat=0), lab) == 0


def test_known_mask_and_ties(ctx, mr, oracle):
    rng = np.random.default_rng(77)
    n2, n1, K = 200, 333, 4
    # few categories and many gaps: plenty of equal weights, so the colour rule and the first-maximum rule matter
    lab = rng.integers(0, K + 1, size=(n2, n1)).astype(np.uint8)
    lab[rng.random((n2, n1)) < 0.4] = 0
    px = rng.integers(0, 256, size=(n2, n1, 3), dtype=np.uint8)
    mean = rng.random((K, 3))
    known = np.where(rng.random((n2, n1)) < 0.02, rng.integers(1, K + 1, size=(n2, n1)), 0).astype(np.uint8)
    mask = (rng.random((n2, n1)) < 0.1).astype(np.uint8)
    _palette(ctx, mr, mean)
    for cats in ([1, 2, 3, 4], [4, 3, 2, 1], [2, 4]):
        ctx.set_known(*_known_lists(known))
        try:
            got = ctx.fill_gaps_host(lab, px, cats, mask=mask, repeat=2)
        finally:
            ctx.set_known(None)
        want = oracle.fill_gaps(lab, px, mean, cats, known=known, mask=mask, repeat=2)
        assert mismatches(got, want) == 0


def test_rgba_and_lab(ctx, mr, oracle):
    rng = np.random.default_rng(78)
    n2, n1, K = 120, 96, 5
    lab = holes(rng, n2, n1, K, 0.45)
    px4 = rng.integers(0, 256, size=(n2, n1, 4), dtype=np.uint8)
    mean4 = rng.random((K, 4))
    pal = mr.Palette(mean4, mean4 - 0.05, mean4 + 0.05, np.full((K, 4), 0.02))
    ctx.set_palette(pal.mean, pal.lo, pal.hi)
    ctx.palette = None
    ctx.set_known(None)
    got = ctx.fill_gaps_host(lab, px4, [1, 2, 3, 4, 5], repeat=2)
    assert mismatches(got, oracle.fill_gaps(lab, px4, mean4, [1, 2, 3, 4, 5], repeat=2)) == 0
    px = np.ascontiguousarray(px4[:, :, :3])
    meanl = mr.api.rgb8_to_lab(rng.integers(0, 256, size=(K, 3), dtype=np.uint8))
    _palette(ctx, mr, meanl, "lab")
    got = ctx.fill_gaps_host(lab, px, [1, 2, 3, 4, 5])
    assert mismatches(got, oracle.fill_gaps(lab, px, meanl, [1, 2, 3, 4, 5], colourspace=oracle.LAB)) == 0


def test_device_entry_with_pitches(ctx, mr, oracle):
    rng = np.random.default_rng(79)
    n2, n1, K = 150, 203, 6          # odd width: byte path; padded pitches on every plane
    lab = holes(rng, n2, n1, K, 0.3)
    px = rng.integers(0, 256, size=(n2, n1, 3), dtype=np.uint8)
    mask = (rng.random((n2, n1)) < 0.1).astype(np.uint8)
    mean = rng.random((K, 3))
    _palette(ctx, mr, mean)
    ctx.set_known(None)
    dev = torch.device("cuda", 0)
    lp, pp, mp, op = 256, 640, 224, 208
    d_lab = torch.zeros((n2, lp), dtype=torch.uint8, device=dev)
    d_lab[:, :n1] = torch.from_numpy(lab).to(dev)
    d_px = torch.zeros((n2, pp), dtype=torch.uint8, device=dev)
    d_px[:, :n1 * 3] = torch.from_numpy(px.reshape(n2, n1 * 3)).to(dev)
    d_mask = torch.zeros((n2, mp), dtype=torch.uint8, device=dev)
    d_mask[:, :n1] = torch.from_numpy(mask).to(dev)
    d_out = torch.full((n2, op), 99, dtype=torch.uint8, device=dev)
    for rep in (1, 2):
        ctx.fill_gaps_dev(d_lab, n1, n2, lp, d_px, pp, 3, [1, 2, 3, 4, 5, 6], d_out, op, mask=d_mask, mask_stride2=mp,
                          repeat=rep, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = d_out[:, :n1].cpu().numpy()
        assert mismatches(got, oracle.fill_gaps(lab, px, mean, [1, 2, 3, 4, 5, 6], mask=mask, repeat=rep)) == 0
        assert int((d_out[:, n1:] != 99).sum()) == 0     # nothing written outside the lines


def test_argument_errors(ctx, mr):
    lab = np.zeros((4, 4), np.uint8)
    px = np.zeros((4, 4, 3), np.uint8)
    _palette(ctx, mr, np.random.default_rng(0).random((3, 3)))
    with pytest.raises(mr.MapRastError):
        ctx.fill_gaps_host(lab, px, [4])            # category outside 1..K
    with pytest.raises(mr.MapRastError):
        ctx.fill_gaps_host(lab, px, [0])
    with pytest.raises(mr.MapRastError):
        ctx.fill_gaps_host(lab, px, [1], repeat=-1)


def test_mirror_api(mr, oracle):
    rng = np.random.default_rng(80)
    n2, n1, K = 64, 80, 4
    lab = holes(rng, n2, n1, K, 0.3)
    img = rng.integers(0, 256, size=(n2, n1, 3), dtype=np.uint8)
    mean = rng.random((K, 3))
    pal = mr.Palette(mean, mean - 0.05, mean + 0.05, np.full((K, 3), 0.02))
    got = mr.fill_gaps(lab.astype(np.int64), categories=(1, 2, 3), stats=pal, original=img, repeat=2)
    assert got.dtype == np.int64
    assert mismatches(got, oracle.fill_gaps(lab, img, mean, [1, 2, 3], repeat=2)) == 0
    with pytest.raises(ValueError):
        mr.fill_gaps(lab, categories=(1,), stats=pal, original=img, missingval=-1)
    with pytest.raises(ValueError):
        mr.fill_gaps(lab, categories=(1,), stats=pal, original=img, match=("color", "std"))
