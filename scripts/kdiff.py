"""GPU box: labels of library A vs library B on a crop of a config sheet; prints where they differ.
python scripts/kdiff.py libA.so libB.so [--cfg 2] [--lines 1024]"""
import argparse, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
def run(cfg, lines, out):
    import numpy as np, torch
    import maprast_b200 as mr
    from maprast_b200 import synth
    seed, palrgb, stats, spec = synth.config_workload(cfg)
    n1, n2, R = spec["n1"], lines, spec["R"]
    dev = torch.device("cuda", 0)
    ctx = mr.Context(0)
    ctx.set_palette(stats.mean, stats.lo, stats.hi, mr._lib.LAB if stats.colourspace == "lab" else mr._lib.RGB)
    px = torch.empty((n2, n1, 3), dtype=torch.uint8, device=dev)
    ctx.synth_scan_dev(seed, 0, palrgb, n1, spec["n2"], 0, n2, px, n1 * 3)
    o = torch.empty((n2, n1), dtype=torch.uint8, device=dev)
    ctx.categorize_clean_dev(px, n1, n2, n1 * 3, 3, R, o, n1)
    torch.cuda.synchronize()
    np.save(out, o.cpu().numpy())
if __name__ == "__main__":
    if sys.argv[1] == "--child":
        run(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]); sys.exit(0)
    ap = argparse.ArgumentParser(); ap.add_argument("a"); ap.add_argument("b"); ap.add_argument("--cfg", type=int, default=2); ap.add_argument("--lines", type=int, default=1024)
    a = ap.parse_args()
    import numpy as np
    outs = []
    for k, lib in enumerate((a.a, a.b)):
        f = f"/tmp/kdiff_{k}.npy"
        subprocess.check_call([sys.executable, __file__, "--child", str(a.cfg), str(a.lines), f], env=dict(os.environ, MAPRAST_LIB=os.path.abspath(lib)))
        outs.append(np.load(f))
    d = np.argwhere(outs[0] != outs[1])
    print("differing pixels:", len(d))
    if len(d):
        ys, xs = d[:, 0], d[:, 1]
        print("first:", d[:20].tolist())
        print("y mod 8 hist:", np.bincount(ys % 8, minlength=8).tolist())
        print("x mod 896 // 28 hist (vote word):", np.bincount((xs % 896) // 28, minlength=32).tolist())
        print("x mod 32 hist:", np.bincount(xs % 32, minlength=32).tolist())
        print("values A/B sample:", [(int(outs[0][y, x]), int(outs[1][y, x])) for y, x in d[:20]])
