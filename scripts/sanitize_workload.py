"""Small workload for compute-sanitizer (scripts/sanitize.sh): every kernel family of libmaprast.so once, on
sizes that finish under the tools' 10-100x slowdown, each result checked against the CPU oracle.
Covers k_fast for R = 0..3 (fused and labels-in), strips that touch the left / right border, lines that
are not a multiple of 16, bands with in-place and remote (peer entry point) halo lines, known-category
points, RGBA, the generic kernel (K > 126, R > 3), the segment prune and fill_gaps."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maprast_b200 as mr                      # noqa: E402
from maprast_b200 import synth                 # noqa: E402
from oracle import oracle as O                 # noqa: E402


def check(name, got, want):
    bad = int((np.asarray(got) != np.asarray(want)).sum())
    print(f"{name}: {bad} mismatches")
    if bad:
        raise SystemExit(f"{name}: MISMATCH")


def main():
    dev = torch.device("cuda", 0)
    ctx = mr.Context(0)
    rng = np.random.default_rng(1)
    for K, R, n1, n2 in [(16, 2, 2048, 72), (8, 1, 1000, 40), (64, 3, 1200, 48), (16, 0, 1024, 24), (40, 2, 960, 40)]:
        pal = synth.make_palette(100 + K, K)
        st = synth.make_stats(100 + K, pal)
        ctx.set_palette(st.mean, st.lo, st.hi)
        scan = O.synth_scan(7, 0, pal, n1, n2)
        got, cnt = ctx.categorize_clean_host(scan, R, counters=True)
        want = O.categorize_clean(scan, st.mean, st.lo, st.hi, R)
        check(f"fused K={K} R={R} {n1}x{n2}", got, want)
        if R:
            lab = O.categorize_u8(scan, st.mean, st.lo, st.hi)
            check(f"clean K={K} R={R}", ctx.clean_host(lab, R), O.majority(lab, R))
            # band with in-place halos and the peer entry point with the halo lines in separate allocations
            a, b = 16, n2 - 16
            d_all = torch.from_numpy(scan).to(dev)
            d_out = torch.zeros((b - a, n1), dtype=torch.uint8, device=dev)
            if n1 % 16 == 0:
                ctx.categorize_clean_dev(d_all[a], n1, b - a, n1 * 3, 3, R, d_out, n1, halo_lo=R, halo_hi=R)
                torch.cuda.synchronize()
                check(f"band in-place halos R={R}", d_out.cpu().numpy(), want[a:b])
                lo = d_all[a - R:a].clone()
                hi = d_all[b:b + R].clone()
                body = d_all[a:b].clone()
                d_out.zero_()
                ctx.categorize_clean_dev_peer(body, n1, b - a, n1 * 3, 3, R, lo, R, hi, R, d_out, n1)
                torch.cuda.synchronize()
                check(f"band remote halos R={R}", d_out.cpu().numpy(), want[a:b])
            # known-category points
            idx = np.unique(rng.integers(0, n1 * n2, size=50))
            cat = rng.integers(1, K + 1, size=idx.size).astype(np.uint8)
            ctx.set_known(idx, cat)
            got = ctx.categorize_clean_host(scan, R)
            ctx.set_known(None)
            check(f"known K={K} R={R}", got, O.majority(O.categorize_u8(scan, st.mean, st.lo, st.hi, known_idx=idx, known_cat=cat), R))
    # generic kernel: K > 126 and R > 3
    K = 130
    pal = synth.make_palette(5, K)
    st = synth.make_stats(5, pal)
    ctx.set_palette(st.mean, st.lo, st.hi)
    scan = O.synth_scan(8, 0, pal, 300, 40)
    check("generic K=130 R=4", ctx.categorize_clean_host(scan, 4), O.categorize_clean(scan, st.mean, st.lo, st.hi, 4))
    # RGBA
    K = 6
    pal = synth.make_palette(6, K)
    mean4 = np.concatenate([pal.astype(np.float64) / 255.0, np.ones((K, 1))], axis=1)
    lo4, hi4 = mean4 - 0.1, mean4 + 0.1
    ctx.set_palette(mean4, lo4, hi4)
    rgba = np.concatenate([O.synth_scan(9, 0, pal, 512, 40), rng.integers(0, 256, size=(40, 512, 1), dtype=np.uint8)], axis=2)
    rgba[:, ::3, 3] = 255
    check("rgba R=1", ctx.categorize_clean_host(rgba, 1),
          O.categorize_clean(rgba, mean4, lo4, hi4, 1, channels=4))
    # segment prune + fill_gaps
    K = 7
    lab = rng.integers(1, K + 1, size=(12, 16)).astype(np.uint8)
    lab = np.kron(lab, np.ones((8, 8), dtype=np.uint8))
    lab[rng.random(lab.shape) < 0.06] = 0
    keep = O.keep_table(range(1, K + 1))
    got, nseg = ctx.clean_segments_host(lab, keep, 0.01, 20)
    want, nseg_o = O.clean_segments(lab, None, range(1, K + 1), 0.01, 20)
    check("segment prune", got, want)
    mean = rng.random((K, 3))
    ctx.set_palette(mean, mean - 0.05, mean + 0.05)
    px = rng.integers(0, 256, size=lab.shape + (3,), dtype=np.uint8)
    check("fill_gaps", ctx.fill_gaps_host(want, px, list(range(1, K + 1)), repeat=2),
          O.fill_gaps(want, px, mean, list(range(1, K + 1)), repeat=2))
    # multi-GPU handle over the same device twice
    m = mr.MultiContext([0, 0])
    pal = synth.make_palette(116, 16)
    st = synth.make_stats(116, pal)
    m.set_palette(st.mean, st.lo, st.hi)
    scan = O.synth_scan(10, 0, pal, 1024, 64)
    check("multi host bands", m.categorize_clean_host(scan, 2), O.categorize_clean(scan, st.mean, st.lo, st.hi, 2))
    m.close()
    ctx.close()
    print("sanitize workload ok")


if __name__ == "__main__":
    main()
