"""_shrink_grow(labels, 1, 1) on the fused config 2 labels, device resident."""
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
lab = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
ctx.categorize_clean_dev(px, n1, n2, n1 * 3, 3, spec["R"], lab, n1)
out = torch.empty_like(lab)
for ne, nd in ((1, 0), (0, 1), (1, 1)):
    f = lambda: ctx._ck(ctx._L.mr_shrink_grow_dev(ctx._h, lab.data_ptr(), n1, n2, n1, ne, nd, out.data_ptr(), n1, None))
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f()
    e1.record(); torch.cuda.synchronize()
    print(f"erode x{ne} dilate x{nd}: {e0.elapsed_time(e1) / 3:.3f} ms")
