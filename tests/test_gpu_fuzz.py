"""GPU (-m gpu): seeded random shapes / windows / palettes / bands / pitches through the device
entry points against the oracle.  Bit-exact; every case prints its parameters on failure."""
import numpy as np
import pytest
import torch

from util import mismatches

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(24))
def test_fused_random_geometry(ctx, mr, oracle, seed):
    from maprast_b200 import synth
    rng = np.random.default_rng(1000 + seed)
    K = int(rng.choice([3, 8, 16, 31, 64, 100]))
    R = int(rng.integers(0, 4))
    pal = synth.make_palette(77 + seed, K)
    st = synth.make_stats(77 + seed, pal, tol=float(rng.choice([0.5, 2.0, 6.0])))
    ctx.set_known(None)
    ctx.set_palette(st.mean, st.lo, st.hi)
    ctx.palette = None
    n1 = int(rng.choice([16, 48, 880, 896, 912, 1000, 1777, 2704, 3001]))
    n2 = int(rng.choice([1, 2, 5, 17, 64, 131]))
    pitch = n1 + int(rng.choice([0, 0, 1, 16, 37]))
    mode = int(rng.choice([0, 0, 0, 1]))
    scan = oracle.synth_scan(0x1234 + seed, seed, pal, n1, n2, mode=mode)
    want = oracle.categorize_clean(scan, st.mean, st.lo, st.hi, R)
    buf = torch.zeros((n2, pitch, 3), dtype=torch.uint8, device="cuda")
    buf[:, :n1] = torch.from_numpy(scan).cuda()
    out = torch.full((n2, pitch), 201, dtype=torch.uint8, device="cuda")
    info = dict(seed=seed, K=K, R=R, n1=n1, n2=n2, pitch=pitch, mode=mode)
    ctx.categorize_clean_dev(buf, n1, n2, pitch * 3, 3, R, out, pitch)
    torch.cuda.synchronize()
    assert mismatches(out[:, :n1].cpu().numpy(), want) == 0, info
    assert bool((out[:, n1:] == 201).all()), info
    # the same sheet as bands with halos
    if n2 >= 8 and R > 0:
        cuts = sorted(set([0, n2] + [int(c) for c in rng.integers(1, n2, size=3)]))
        got = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
        for a, b in zip(cuts[:-1], cuts[1:]):
            ctx.categorize_clean_dev(buf[a:], n1, b - a, pitch * 3, 3, R, got[a:b], n1, halo_lo=min(R, a),
                                     halo_hi=min(R, n2 - b))
        torch.cuda.synchronize()
        assert mismatches(got.cpu().numpy(), want) == 0, dict(info, cuts=cuts)
    # standalone clean of the categorise-only labels
    if R > 0:
        pre = oracle.categorize_u8(scan, st.mean, st.lo, st.hi)
        lab = torch.from_numpy(pre).cuda()
        out2 = torch.empty_like(lab)
        ctx.clean_dev(lab, n1, n2, n1, R, out2, n1)
        torch.cuda.synchronize()
        assert mismatches(out2.cpu().numpy(), want) == 0, info
        # the same with the label bound supplied by the caller (no label-maximum pass, no synchronisation)
        out2.zero_()
        ctx.clean_dev(lab, n1, n2, n1, R, out2, n1, max_label=st.K)
        torch.cuda.synchronize()
        assert mismatches(out2.cpu().numpy(), want) == 0, info


@pytest.mark.parametrize("seed", range(12))
def test_segments_random_geometry(ctx, oracle, seed):
    rng = np.random.default_rng(2000 + seed)
    n1 = int(rng.choice([1, 15, 16, 17, 4095, 4096, 4097, 5000]))
    n2 = int(rng.choice([1, 2, 9, 40]))
    K = int(rng.integers(1, 9))
    f = rng.integers(0, K + 1, size=(n2 // 4 + 1, n1 // 12 + 1)).astype(np.uint8)
    lab = np.kron(f, np.ones((4, 12), dtype=np.uint8))[:n2, :n1].copy()
    sp = rng.random((n2, n1)) < 0.03
    lab[sp] = rng.integers(0, K + 1, size=int(sp.sum()))
    cats = tuple(int(c) for c in rng.choice(np.arange(0, K + 1), size=max(1, K // 2 + 1), replace=False))
    lt, pt = float(rng.choice([0.0, 0.01, 0.3])), float(rng.choice([0, 5, 20, 200]))
    ctx.set_known(None)
    got, nseg = ctx.clean_segments_host(lab, oracle.keep_table(cats), lt, pt)
    want, nseg_o = oracle.clean_segments(lab, None, cats, lt, pt)
    assert mismatches(got, want) == 0 and nseg == nseg_o, dict(seed=seed, n1=n1, n2=n2, cats=cats, lt=lt, pt=pt)
