import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import maprast_b200 as mr
from maprast_b200 import synth
ctx = mr.Context(0)
for cfg in (2, 4):
    seed, palrgb, stats, spec = synth.config_workload(cfg)
    n1, n2, R = spec["n1"], 4000, spec["R"]
    ctx.set_palette(stats.mean, stats.lo, stats.hi, mr._lib.LAB if stats.colourspace == "lab" else mr._lib.RGB)
    d = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda")
    ctx.synth_scan_dev(seed, 0, palrgb, n1, 20000, 0, n2, d, n1 * 3)
    out = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
    ctx.enable_counters(True); ctx.read_counters(reset=True)
    ctx.categorize_clean_dev(d, n1, n2, n1 * 3, 3, R, out, n1)
    c = ctx.read_counters(reset=True); ctx.enable_counters(False)
    print(cfg, {k: v / c["n_pixels"] for k, v in c.items()})
