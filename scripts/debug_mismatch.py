import sys, numpy as np
sys.path.insert(0, '.')
import maprast_b200 as mr
from maprast_b200 import synth
from oracle import oracle as O
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n1 = int(sys.argv[2]) if len(sys.argv) > 2 else 1984
n2 = int(sys.argv[3]) if len(sys.argv) > 3 else 300
seed, pal, st, spec = synth.config_workload(cfg)
R = spec["R"]
ctx = mr.Context(0)
ctx.set_palette(st.mean, st.lo, st.hi, mr._lib.LAB if st.colourspace == "lab" else mr._lib.RGB)
scan = O.synth_scan(seed, 0, pal, n1, n2)
cs = O.LAB if st.colourspace == "lab" else O.RGB
for RR in (0, R):
    got, cnt = ctx.categorize_clean_host(scan, RR, counters=True)
    pre = O.categorize_u8(scan, st.mean, st.lo, st.hi, cs)
    want = pre if RR == 0 else O.majority(pre, RR)
    bad = np.argwhere(got != want)
    print("R", RR, "mismatches", len(bad), "of", n1 * n2, cnt)
    if len(bad):
        ys, xs = bad[:, 0], bad[:, 1]
        print(" x%32 hist", np.bincount(xs % 32, minlength=32))
        print(" x//32 %30 hist", np.bincount(((xs // 32) % 30), minlength=30))
        print(" y%16 hist", np.bincount(ys % 16, minlength=16))
        for (y, x) in bad[:10]:
            w = pre[max(0, y - RR):y + RR + 1, max(0, x - RR):x + RR + 1]
            print(" at", y, x, "got", got[y, x], "want", want[y, x], "centre", pre[y, x], "window", w.tolist())
