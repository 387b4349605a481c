"""Kernel A/B bench (GPU box): python scripts/kbench.py [--cfg 2] [--steps 10] [--ref path/to/reference.so]
Times the device-resident fused kernel of the library selected by MAPRAST_LIB (default: the in-tree one)
and, with --ref, compares its labels on the WHOLE sheet with another build's (run in a subprocess)."""
import argparse, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--lines", type=int, default=0, help="0 = the config's own n2")
    ap.add_argument("--digest-only", action="store_true")
    ap.add_argument("--ref", default=None)
    ap.add_argument("--permute", action="store_true", help="experiment: 32x32 pixel transposition inside 1024-px groups (coherent lanes)")
    args = ap.parse_args()
    import torch
    import maprast_b200 as mr
    from maprast_b200 import synth
    seed, palrgb, stats, spec = synth.config_workload(args.cfg)
    n1, n2, R = spec["n1"], args.lines or spec["n2"], spec["R"]
    ns = spec["n_sheets"] if spec["n_sheets"] == 1 else 32
    dev = torch.device("cuda", 0)
    ctx = mr.Context(0)
    ctx.set_palette(stats.mean, stats.lo, stats.hi, mr._lib.LAB if stats.colourspace == "lab" else mr._lib.RGB)
    px = torch.empty((ns, n2, n1, 3), dtype=torch.uint8, device=dev)
    for s in range(ns):
        ctx.synth_scan_dev(seed, s, palrgb, n1, spec["n2"] if ns == 1 else n2, 0, n2, px[s], n1 * 3)
    if args.permute:
        g = n1 // 1024
        v = px[:, :, :g * 1024].reshape(ns, n2, g, 32, 32, 3).transpose(3, 4).contiguous().reshape(ns, n2, g * 1024, 3)
        px[:, :, :g * 1024] = v
        del v
    out = torch.empty((ns, n2, n1), dtype=torch.uint8, device=dev)
    def step():
        if ns > 1:
            ctx.batch_categorize_clean_dev(ns, px, n1 * n2 * 3, n1, n2, n1 * 3, 3, R, out, n1 * n2, n1)
        else:
            ctx.categorize_clean_dev(px[0], n1, n2, n1 * 3, 3, R, out[0], n1)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        step(); ev[k + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps))
    digest = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
    res = {"lib": os.path.basename(mr._lib.LIB_PATH), "cfg": args.cfg, "px": ns * n1 * n2, "ms_min": ms[0], "ms_med": ms[len(ms) // 2],
           "gpx_s": ns * n1 * n2 / ms[len(ms) // 2] / 1e6, "frac_6527": 4 * ns * n1 * n2 / (ms[len(ms) // 2] * 1e-3) / 1e9 / 6527.5,
           "digest": digest}
    if args.ref and not args.digest_only:
        env = dict(os.environ, MAPRAST_LIB=args.ref)
        r = subprocess.run([sys.executable, __file__, "--cfg", str(args.cfg), "--steps", "3", "--lines", str(args.lines), "--digest-only"],
                           env=env, capture_output=True, text=True)
        try:
            ref = json.loads(r.stdout.strip().splitlines()[-1])
            res["ref_ms_med"] = ref["ms_med"]; res["same_labels_as_ref"] = ref["digest"] == digest
        except Exception:
            res["ref_error"] = (r.stdout + r.stderr)[-400:]
    print(json.dumps(res))

if __name__ == "__main__":
    main()
