import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import maprast_b200 as mr
from maprast_b200 import synth
from oracle import oracle as O
ctx = mr.Context(0)
bad = 0
for seed in range(60):
    rng = np.random.default_rng(9000 + seed)
    K = int(rng.choice([4, 16, 16, 16, 40, 64]))
    R = int(rng.integers(1, 4))
    pal = synth.make_palette(500 + seed, K)
    st = synth.make_stats(500 + seed, pal)
    ctx.set_known(None); ctx.set_palette(st.mean, st.lo, st.hi)
    n1 = int(rng.choice([896, 1792, 2000, 4096, 6000, 8192]))
    n2 = int(rng.integers(300, 2200))
    scan = O.synth_scan(0x77 + seed, seed, pal, n1, n2)
    want = O.categorize_clean(scan, st.mean, st.lo, st.hi, R)
    d = torch.from_numpy(scan).cuda()
    out = torch.empty((n2, n1), dtype=torch.uint8, device="cuda")
    for rep in range(3):
        out.fill_(250)
        ctx.categorize_clean_dev(d, n1, n2, n1 * 3, 3, R, out, n1)
        torch.cuda.synchronize()
        m = int((out.cpu().numpy() != want).sum())
        if m:
            bad += 1
            print("MISMATCH", dict(seed=seed, K=K, R=R, n1=n1, n2=n2, rep=rep, m=m))
print("done, bad =", bad)
