"""Per-warp timeline of one fused launch (MR_TIMELINE build): start-up, first vote, finish spread.
usage (GPU box): MAPRAST_LIB=.../libmaprast_tl.so python scripts/timeline.py [cfg] [lines]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maprast_b200 as mr
from maprast_b200 import synth
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
seed, palrgb, stats, spec = synth.config_workload(cfg)
n1, n2, R = spec["n1"], (int(sys.argv[2]) if len(sys.argv) > 2 else spec["n2"]), spec["R"]
ctx = mr.Context(0)
ctx.set_palette(stats.mean, stats.lo, stats.hi)
px = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda")
ctx.synth_scan_dev(seed, 0, palrgb, n1, spec["n2"], 0, n2, px, n1 * 3)
out = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
for _ in range(3):
    ctx.categorize_clean_dev(px, n1, n2, n1 * 3, 3, R, out, n1)
torch.cuda.synchronize()
L = mr._lib.load()
nw = 148 * 16
buf = np.zeros((nw, 4), dtype=np.uint64)
rc = L.mr_debug_timeline(C.c_void_p(buf.ctypes.data), nw)
assert rc == 0, rc
t = buf.astype(np.float64)
t0 = t[:, 0].min()
start, first, end, lines = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, t[:, 3]
act = lines > 0
print(f"cfg {cfg} lines {n2}: warps {act.sum()} active; kernel span {end.max():.1f} us")
print(f"start  (us): min {start.min():.1f} med {np.median(start):.1f} max {start.max():.1f}")
print(f"first vote : min {first[act].min():.1f} med {np.median(first[act]):.1f} max {first[act].max():.1f}")
print(f"end        : min {end.min():.1f} p10 {np.percentile(end,10):.1f} med {np.median(end):.1f} p90 {np.percentile(end,90):.1f} max {end.max():.1f}")
print(f"lines/warp : min {lines.min():.0f} med {np.median(lines):.0f} max {lines.max():.0f}; us per line (busy): {np.median((end-first)[act]/np.maximum(lines[act],1)):.3f}")
idle = (end.max() - end).sum() / (end.max() * nw)
print(f"idle at the end: {100*idle:.2f} % of warp-time; before first vote: {100*first[act].sum()/(end.max()*nw):.2f} %")
sm_end = end.reshape(148, 16).max(axis=1)
print(f"per-SM finish: min {sm_end.min():.1f} med {np.median(sm_end):.1f} max {sm_end.max():.1f}")
