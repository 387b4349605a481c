import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maprast_b200 as mr
from oracle import oracle as O
from maprast_b200 import synth
O.build()
seed, palrgb, stats, spec = synth.config_workload(1)
n1, n2, K = 640, 400, spec["K"]
scan = O.synth_scan(seed, 0, palrgb, n1, n2)
ctx = mr.Context(0)
blurred = O.blur(scan, 1, 1)
rng = np.random.default_rng(3)
points = []
for c in range(K):
    d = np.abs(scan.astype(np.int32) - palrgb[c].astype(np.int32)).sum(axis=2)
    js, is_ = np.nonzero(d < 12)
    pick = rng.choice(len(js), size=12, replace=False)
    points.append([(float(is_[p] + 1), float(js[p] + 1)) for p in pick])
tol = np.full(K, 2.5)
pal = mr.Palette.from_points(blurred, points, tol)
ctx.set_palette(pal.mean, pal.lo, pal.hi)
ctx.set_known(None)
got = ctx.categorize_clean_f64_host(blurred, 0)
want = O.categorize_f64_image(blurred, pal.mean, pal.lo, pal.hi)
bad = got != want
print("mismatches", bad.sum(), "cols", np.nonzero(bad.any(axis=0))[0][:20], "rows", np.nonzero(bad.any(axis=1))[0][:20])
js, is_ = np.nonzero(bad)
for k in range(min(5, len(js))):
    j, i = js[k], is_[k]
    print(i, j, blurred[j, i], got[j, i], want[j, i], pal.lo[got[j, i] - 1] if got[j, i] else None, pal.hi[got[j, i] - 1] if got[j, i] else None)
