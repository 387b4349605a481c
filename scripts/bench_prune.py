"""segment prune on a config-2 sheet: on the majority-filtered labels (few, big segments) and on the raw
categorise-only labels (what the reference's own pipeline prunes: speckle = millions of tiny segments)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maprast_b200 as mr
from maprast_b200 import synth
seed, palrgb, stats, spec = synth.config_workload(2)
n1, n2, K = spec["n1"], (int(sys.argv[1]) if len(sys.argv) > 1 else spec["n2"]), spec["K"]
ctx = mr.Context(0)
ctx.set_palette(stats.mean, stats.lo, stats.hi)
px = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda")
ctx.synth_scan_dev(seed, 0, palrgb, n1, spec["n2"], 0, n2, px, n1 * 3)
keep = np.zeros(256, dtype=np.uint8); keep[1:K + 1] = 1
out = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
for R in (spec["R"], 0):
    lab = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
    ctx.categorize_clean_dev(px, n1, n2, n1 * 3, 3, R, lab, n1)
    nseg = ctx.clean_segments_dev(lab, n1, n2, n1, keep, 0.01, 20, out, n1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ctx.clean_segments_dev(lab, n1, n2, n1, keep, 0.01, 20, out, n1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"labels after window R={R}: {nseg} segments, {ms:.3f} ms = {n1 * n2 / ms / 1e6:.1f} Gpx/s, zeros out {int((out == 0).sum())}")
