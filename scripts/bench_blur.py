"""blur + Float64 categorise on a config-2 wide sheet (device resident), timing.  usage: python scripts/bench_blur.py [lines]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maprast_b200 as mr
from maprast_b200 import synth
seed, palrgb, stats, spec = synth.config_workload(2)
n1, n2, R, K = spec["n1"], (int(sys.argv[1]) if len(sys.argv) > 1 else 5000), spec["R"], spec["K"]
ctx = mr.Context(0)
ctx.set_palette(stats.mean, stats.lo, stats.hi)
px = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda")
ctx.synth_scan_dev(seed, 0, palrgb, n1, spec["n2"], 0, n2, px, n1 * 3)
f64 = torch.empty((n2, n1, 3), dtype=torch.float64, device="cuda")
lab = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
L = ctx._L
def blur():
    ctx._ck(L.mr_blur_dev(ctx._h, px.data_ptr(), n1, n2, n1 * 3, 3, 1, 1, f64.data_ptr(), n1 * 24, None))
def cat(Rr):
    ctx._ck(L.mr_categorize_clean_f64_dev(ctx._h, f64.data_ptr(), n1, n2, n1 * 24, Rr, lab.data_ptr(), n1, None))
def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
mpx = n1 * n2 / 1e6
tb = timeit(blur)
print(f"blur Window(1) u8 -> f64, {mpx:.0f} Mpx: {tb:.3f} ms = {mpx / tb:.1f} Gpx/s, {27 * n1 * n2 / tb / 1e6 / 6527.5:.3f} of the 27 B/px roofline")
tc = timeit(lambda: cat(0))
print(f"categorise f64 (K={K}): {tc:.3f} ms = {mpx / tc:.1f} Gpx/s, {25 * n1 * n2 / tc / 1e6 / 6527.5:.3f} of the 25 B/px roofline")
tcc = timeit(lambda: cat(R))
print(f"categorise f64 + clean R={R}: {tcc:.3f} ms = {mpx / tcc:.1f} Gpx/s")
