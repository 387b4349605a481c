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
