"""GPU (-m gpu): `_shrink_grow(A, i, j)` = erosions then dilations (/root/reference/src/clean.jl:96-140) through the
C ABI against the oracle."""
import numpy as np
import pytest
import torch

from util import mismatches
from test_oracle_morph import random_labels

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(12))
def test_shrink_grow_random(ctx, mr, oracle, seed):
    rng = np.random.default_rng(300 + seed)
    n1 = int(rng.choice([1, 2, 3, 4, 5, 63, 64, 65, 1023, 1025, 2050]))
    n2 = int(rng.choice([1, 2, 3, 17, 60]))
    A = random_labels(rng, n2, n1, int(rng.integers(1, 9)), float(rng.choice([0.0, 0.05, 0.5, 0.95])))
    for i, j in ((1, 0), (0, 1), (2, 3), (0, 0), (3, 0)):
        assert mismatches(ctx.shrink_grow_host(A, i, j), oracle.shrink_grow(A, i, j)) == 0
    assert mismatches(mr.erode(A, ctx=ctx), oracle.shrink_grow(A, 1, 0)) == 0
    assert mismatches(mr.dilate(A, ctx=ctx), oracle.shrink_grow(A, 0, 1)) == 0


def test_shrink_grow_device_entry_with_pitches(ctx, oracle):
    rng = np.random.default_rng(9)
    n2, n1 = 77, 333
    A = random_labels(rng, n2, n1, 6, 0.2)
    d_in = torch.full((n2, n1 + 11), 9, dtype=torch.uint8, device="cuda")
    d_in[:, :n1] = torch.from_numpy(A).cuda()
    d_out = torch.full((n2, n1 + 5), 77, dtype=torch.uint8, device="cuda")
    ctx._ck(ctx._L.mr_shrink_grow_dev(ctx._h, d_in.data_ptr(), n1, n2, n1 + 11, 2, 3, d_out.data_ptr(), n1 + 5,
                                      torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert mismatches(d_out[:, :n1].cpu().numpy(), oracle.shrink_grow(A, 2, 3)) == 0
    assert bool((d_out[:, n1:] == 77).all())


def test_recolor(ctx, mr, oracle):
    """`_recolor` (src/selectcolors.jl:650-655): display layer of a label image, bit-identical copies of the mean colours"""
    rng = np.random.default_rng(4)
    n2, n1, K = 37, 301, 5
    src = rng.random((n2, n1, 3))
    points = [[(float(rng.integers(1, n1 + 1)), float(rng.integers(1, n2 + 1))) for _ in range(4)] for _ in range(K)]
    points[2] = []                                            # a category without points: the zero colour
    lab = rng.integers(0, K + 3, size=(n2, n1)).astype(np.uint8)      # labels above K: transparent
    got = mr.recolor(lab, src, points, ctx=ctx)
    colors = np.stack([mr.mean_color(src, ps) for ps in points])
    want = oracle.recolor(lab, colors)
    assert got.tobytes() == want.tobytes()
    assert (got[lab == 0] == 0).all() and (got[(lab >= 1) & (lab <= K)][:, 3] == 1).all() and (colors[2] == 0).all()
    d_lab = torch.from_numpy(lab).cuda()
    d_out = torch.full((n2, n1 * 4 + 6), -7.0, dtype=torch.float64, device="cuda")
    ctx._ck(ctx._L.mr_recolor_dev(ctx._h, d_lab.data_ptr(), n1, n2, n1, K, colors.ctypes.data, d_out.data_ptr(), (n1 * 4 + 6) * 8,
                                  torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert d_out[:, :n1 * 4].cpu().numpy().reshape(n2, n1, 4).tobytes() == want.tobytes()
    assert bool((d_out[:, n1 * 4:] == -7.0).all())
