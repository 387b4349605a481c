"""Full-size check of the segment prune against the sequential oracle on raw categorise-only labels
(config 2 sheet, ~10 M segments): labels and segment count."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maprast_b200 as mr
from maprast_b200 import synth
from oracle import oracle as O
seed, palrgb, stats, spec = synth.config_workload(2)
n1, n2, K = spec["n1"], (int(sys.argv[1]) if len(sys.argv) > 1 else 8000), spec["K"]
ctx = mr.Context(0)
ctx.set_palette(stats.mean, stats.lo, stats.hi)
px = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda")
ctx.synth_scan_dev(seed, 0, palrgb, n1, spec["n2"], 0, n2, px, n1 * 3)
lab = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
ctx.categorize_clean_dev(px, n1, n2, n1 * 3, 3, 0, lab, n1)
keep = np.zeros(256, dtype=np.uint8); keep[1:K + 1] = 1
out = torch.empty_like(lab)
counts = [ctx.clean_segments_dev(lab, n1, n2, n1, keep, 0.01, 20, out, n1) for _ in range(3)]
t0 = time.perf_counter()
want, nseg = O.clean_segments(lab.cpu().numpy(), None, range(1, K + 1), 0.01, 20)
print(f"oracle {time.perf_counter() - t0:.1f} s; segments oracle {nseg} gpu {counts}; label mismatches {int((out.cpu().numpy() != want).sum())}")
