"""Time the kernels of the segment prune on the BASELINE config 2 label sheet (run under ncu for a launch list)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import maprast_b200 as mr
from maprast_b200 import synth
seed, palrgb, stats, spec = synth.config_workload(2)
n1, n2, R, K = spec["n1"], spec["n2"], spec["R"], spec["K"]
ctx = mr.Context(0)
ctx.set_palette(stats.mean, stats.lo, stats.hi)
d = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda")
ctx.synth_scan_dev(seed, 0, palrgb, n1, n2, 0, n2, d, n1 * 3)
out = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
ctx.categorize_clean_dev(d, n1, n2, n1 * 3, 3, R, out, n1)
keep = np.zeros(256, dtype=np.uint8); keep[1:K + 1] = 1
seg = torch.empty_like(out)
for _ in range(2):
    n = ctx.clean_segments_dev(out, n1, n2, n1, keep, 0.01, 20, seg, n1)
torch.cuda.synchronize()
print("segments", n)
