"""Multi-GPU C ABI on a real multi-GPU box (run under `gpurun --gpus N`): one process, one mr_multi handle over
all visible devices.  (1) parity: a sheet through N bands (host path and device-resident peer-halo path, with
known-category points on the seams) equals the single-GPU result; (2) timing of config 3 (60000 x 40000) through
mr_multi_categorize_clean_host (pinned host sheet, H2D / kernel / D2H per GPU) and through
mr_multi_categorize_clean_dev (bands resident, halo lines over NVLink peer access)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maprast_b200 as mr                      # noqa: E402
from maprast_b200 import synth                 # noqa: E402


def main():
    ndev = torch.cuda.device_count()
    seed, palrgb, stats, spec = synth.config_workload(3)
    R = spec["R"]
    m = mr.MultiContext(list(range(ndev)))
    m.set_palette(stats.mean, stats.lo, stats.hi)
    one = mr.Context(0)
    one.set_palette(stats.mean, stats.lo, stats.hi)
    res = {"n_dev": ndev}
    # ---- parity on a 4096 x (64 * ndev + 40) sheet
    n1, n2 = 4096, 64 * ndev + 40
    d = torch.empty((n2, n1, 3), dtype=torch.uint8, device="cuda:0")
    one.synth_scan_dev(seed, 0, palrgb, n1, n2, 0, n2, d, n1 * 3)
    scan = d.cpu().numpy()
    rng = np.random.default_rng(1)
    seams = [m.band_range(n2, i)[0] for i in range(1, ndev)]
    js = np.concatenate([np.arange(max(s - 3, 0), min(s + 3, n2)) for s in seams] + [rng.integers(0, n2, size=50)]) if seams else rng.integers(0, n2, size=50)
    idx = np.unique(rng.integers(0, n1, size=js.size * 4) + np.repeat(js, 4) * n1)
    cat = rng.integers(1, spec["K"] + 1, size=idx.size).astype(np.uint8)
    for known in (False, True):
        one.set_known(idx, cat) if known else one.set_known(None)
        m.set_known(idx, cat) if known else m.set_known(None)
        want = one.categorize_clean_host(scan, R)
        got = m.categorize_clean_host(scan, R)
        res[f"host_bands_mismatches_known{int(known)}"] = int((got != want).sum())
        bands, outs = [], []
        for i in range(ndev):
            a, n = m.band_range(n2, i)
            bands.append(torch.from_numpy(scan[a:a + n].copy()).to(f"cuda:{i}"))
            outs.append(torch.zeros((n, n1), dtype=torch.uint8, device=f"cuda:{i}"))
        for i in range(ndev):
            torch.cuda.synchronize(i)
        m.categorize_clean_dev(bands, n1, n2, n1 * 3, 3, R, outs, n1)
        got = np.concatenate([o.cpu().numpy() for o in outs])
        res[f"dev_bands_mismatches_known{int(known)}"] = int((got != want).sum())
    one.set_known(None)
    m.set_known(None)
    del bands, outs, d
    # ---- timing, config 3
    n1, n2 = spec["n1"], spec["n2"]
    px_total = n1 * n2
    bands, outs = [], []
    for i in range(ndev):
        a, n = m.band_range(n2, i)
        c = mr.Context(i)
        b = torch.empty((n, n1, 3), dtype=torch.uint8, device=f"cuda:{i}")
        c.synth_scan_dev(seed, 0, palrgb, n1, n2, a, n, b, n1 * 3)
        c.close()
        bands.append(b)
        outs.append(torch.empty((n, n1), dtype=torch.uint8, device=f"cuda:{i}"))
    for i in range(ndev):
        torch.cuda.synchronize(i)
    for _ in range(3):
        m.categorize_clean_dev(bands, n1, n2, n1 * 3, 3, R, outs, n1)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        m.categorize_clean_dev(bands, n1, n2, n1 * 3, 3, R, outs, n1)     # synchronous: returns when every band is done
        ts.append(time.perf_counter() - t0)
    ts.sort()
    res["dev_resident_cfg3"] = {"ms_median": ts[len(ts) // 2] * 1e3, "ms_min": ts[0] * 1e3,
                                "Mpx_s": px_total / ts[len(ts) // 2] / 1e6,
                                "what": "mr_multi_categorize_clean_dev, wall clock around the call (host threads + launches + sync included)"}
    # host-resident sheet in pinned memory (7.2 GB in, 2.4 GB out)
    lib = mr._lib.load()
    hin = lib.mr_host_alloc(px_total * 3)
    hout = lib.mr_host_alloc(px_total)
    if hin and hout:
        import ctypes as C
        host_in = np.ctypeslib.as_array((C.c_uint8 * (px_total * 3)).from_address(hin)).reshape(n2, n1, 3)
        host_out = np.ctypeslib.as_array((C.c_uint8 * px_total).from_address(hout)).reshape(n2, n1)
        for i in range(ndev):
            a, n = m.band_range(n2, i)
            host_in[a:a + n] = bands[i].cpu().numpy()
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            m.categorize_clean_host(host_in, R, out=host_out)
            ts.append(time.perf_counter() - t0)
        ts = sorted(ts[1:])
        same = all(bool((torch.from_numpy(host_out[m.band_range(n2, i)[0]:m.band_range(n2, i)[0] + m.band_range(n2, i)[1]]).to(f"cuda:{i}") == outs[i]).all()) for i in range(ndev))
        res["host_resident_cfg3"] = {"ms_median": ts[len(ts) // 2] * 1e3, "Mpx_s": px_total / ts[len(ts) // 2] / 1e6,
                                     "labels_equal_device_path": same,
                                     "what": "mr_multi_categorize_clean_host on a pinned 7.2 GB sheet: H2D + kernel + D2H on every GPU at once"}
        lib.mr_host_free(hin)
        lib.mr_host_free(hout)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
