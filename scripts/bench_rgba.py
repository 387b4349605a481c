"""RGBA{N0f8} sheet (config 2 colours + opaque alpha, 4-channel statistics) through the device entry point."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import maprast_b200 as mr
from maprast_b200 import synth
seed, palrgb, st, spec = synth.config_workload(2)
n1, n2, R, K = spec["n1"], spec["n2"], spec["R"], spec["K"]
ctx = mr.Context(0)
mean = np.concatenate([st.mean, np.full((K, 1), 1.0)], axis=1)
mn = np.concatenate([st.min, np.full((K, 1), 0.9)], axis=1)
mx = np.concatenate([st.max, np.full((K, 1), 1.0)], axis=1)
sd = np.concatenate([st.sd, np.full((K, 1), 0.01)], axis=1)
pal = mr.Palette(mean, mn, mx, sd, np.full(K, 2.0))
ctx.set_palette(pal.mean, pal.lo, pal.hi, mr._lib.RGB, 4)
rgb = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda")
ctx.synth_scan_dev(seed, 0, palrgb, n1, n2, 0, n2, rgb, n1 * 3)
rgba = torch.full((n2, n1, 4), 255, dtype=torch.uint8, device="cuda")
rgba[:, :, :3] = rgb
del rgb
out = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
for _ in range(2):
    ctx.categorize_clean_dev(rgba, n1, n2, n1 * 4, 4, R, out, n1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = ctx.launch_count()
e0.record()
for _ in range(5):
    ctx.categorize_clean_dev(rgba, n1, n2, n1 * 4, 4, R, out, n1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("RGBA fused: %.3f ms  %.1f Gpx/s  launches/call %.1f" % (ms, n1 * n2 / ms / 1e6, (ctx.launch_count() - l0) / 5))
