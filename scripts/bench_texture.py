"""stds / stripes / PixelProp categoriser on the first lines of the config 2 sheet (blurred), device resident."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maprast_b200 as mr
from maprast_b200 import synth
seed, palrgb, stats, spec = synth.config_workload(2)
n1, n2, K = spec["n1"], (int(sys.argv[1]) if len(sys.argv) > 1 else 4096), spec["K"]
ctx = mr.Context(0)
ctx.set_palette(stats.mean, stats.lo, stats.hi)
L = ctx._L
px = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda")
ctx.synth_scan_dev(seed, 0, palrgb, n1, spec["n2"], 0, n2, px, n1 * 3)
f64 = torch.empty((n2, n1, 3), dtype=torch.float64, device="cuda")
sd = torch.empty((n2, n1), dtype=torch.float64, device="cuda")
st = torch.empty((n2, n1, 2), dtype=torch.float64, device="cuda")
lab = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
z1, z2 = np.full(K, 0.5), np.zeros((K, 2))
ctx.set_texture_stats(z1, z1 - 0.6, z1 + 0.6, z2, z2 - 50.0, z2 + 50.0)
fns = {
    "blur": lambda: ctx._ck(L.mr_blur_dev(ctx._h, px.data_ptr(), n1, n2, n1 * 3, 3, 1, 1, f64.data_ptr(), n1 * 24, None)),
    "stds": lambda: ctx._ck(L.mr_stds_dev(ctx._h, f64.data_ptr(), n1, n2, n1 * 24, sd.data_ptr(), n1 * 8, None)),
    "stripes": lambda: ctx._ck(L.mr_stripes_dev(ctx._h, f64.data_ptr(), n1, n2, n1 * 24, 2, st.data_ptr(), n1 * 16, None)),
    "props": lambda: ctx._ck(L.mr_categorize_clean_props_dev(ctx._h, f64.data_ptr(), n1 * 24, sd.data_ptr(), n1 * 8, st.data_ptr(),
                                                             n1 * 16, n1, n2, 7, 0, lab.data_ptr(), n1, None)),
}
for name, fn in fns.items():
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name}: {ms:.3f} ms = {n1 * n2 / ms / 1e6:.1f} Gpx/s")
