"""fill_gaps on a config-2 sized label sheet (device resident): timing, and the target of an ncu capture.
usage: python scripts/bench_fill.py [gap_fraction]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maprast_b200 as mr
from maprast_b200 import synth
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
seed, palrgb, stats, spec = synth.config_workload(2)
n1, n2, R, K = spec["n1"], spec["n2"], spec["R"], spec["K"]
ctx = mr.Context(0)
ctx.set_palette(stats.mean, stats.lo, stats.hi)
px = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda")
ctx.synth_scan_dev(seed, 0, palrgb, n1, n2, 0, n2, px, n1 * 3)
lab = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
ctx.categorize_clean_dev(px, n1, n2, n1 * 3, 3, R, lab, n1)
if frac > 0:
    g = torch.Generator(device="cuda").manual_seed(1)
    lab[torch.rand((n2, n1), device="cuda", generator=g) < frac] = 0
out = torch.empty_like(lab)
cats = list(range(1, K + 1))
for _ in range(2):
    ctx.fill_gaps_dev(lab, n1, n2, n1, px, n1 * 3, 3, cats, out, n1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ctx.fill_gaps_dev(lab, n1, n2, n1, px, n1 * 3, 3, cats, out, n1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"gap fraction {frac}: {ms:.4f} ms, {n1 * n2 / ms / 1e6:.1f} Gpx/s, {2 * n1 * n2 / ms / 1e6 / 6527.5:.3f} of the 2 B/px roofline; "
      f"gaps in {int((lab == 0).sum())} out {int((out == 0).sum())}")
