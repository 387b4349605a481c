#!/usr/bin/env python
"""bench.py -- megapixels/s of the fused categorise+clean hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over the whole synthetic sheet (SURVEY.md 8(d) generator).
  N = 1 : BASELINE config 2 -- 20000 x 20000 RGB{N0f8}, 16-colour palette, clean window 5 (R = 2);
          configs 1, 3, 4 and 5 are measured beside it (`next_rows.cfg*`, each with its roofline).
  N > 1 : BASELINE config 3 -- the 60000 x 40000 (2.4 Gpx) sheet split into N row bands (STRONG
          scaling), one fused launch per rank and step.  The R halo lines of a band are read in
          place from the neighbour band's HBM (CUDA IPC peer memory over NVLink, --halo peer, the
          default) or exchanged with NCCL send/recv before every launch (--halo nccl).  Config 5
          (256 sheets of 4096 x 4096, whole sheets per rank) is measured beside it.
          --workload weak2 = one config-2 sheet per GPU (weak scaling); cfgN = that config.
`value` times device-resident inputs; `e2e` times the public host-buffer API (pinned host
memory -> H2D -> kernel -> D2H) for the same sheet.  Prints ONE JSON line on rank 0.
The CPU oracle (oracle/) is only used for the cpu_baseline leg and for --impl reference.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "megapixels/sec fused categorize+clean"
UNIT = "Mpx/s"
ALGO_BYTES_PER_PX = 4   # 3 B RGB{N0f8} in + 1 B UInt8 label out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "weak2", "cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--halo", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1 bands: read the halo lines in place from the neighbour's HBM (CUDA IPC peer "
                         "memory) or exchange them with NCCL send/recv before every launch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def host_cores():
    """Cores this process may run on (affinity mask), not what OMP_NUM_THREADS says."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
    """Per-launch DRAM bytes of the fused kernel from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except Exception:
            pass
    return None


class NvmlSampler:
    """SM clock and throttle reasons through NVML, polled in a thread about every millisecond: the timed region
    of the headline is ~11 ms, which nvidia-smi's 100 ms loop sees once or twice.  Used beside ClockSampler."""

    def __init__(self, gpu_index):
        self.ok = False
        self.sm, self.reasons = [], 0
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def start(self):
        if not self.ok:
            return
        self.run = True

        def loop():
            nv = self.nv
            while self.run:
                try:
                    self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    self.reasons |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    break
                time.sleep(0.0005)
        self.t = threading.Thread(target=loop, daemon=True)
        self.t.start()

    def stop(self):
        if not self.ok:
            return None
        self.run = False
        self.t.join(timeout=1)
        nv = self.nv
        names = []
        for bit, name in ((getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8), "hw_slowdown"),
                          (getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                          (getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                          (getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4), "sw_power_cap")):
            if self.reasons & bit:
                names.append(name)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_min_mhz": min(self.sm) if self.sm else None,
                "sm_max_mhz": self.max, "samples": len(self.sm), "reasons": sorted(names)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_spec(args, world):
    from maprast_b200 import synth
    wl = args.workload
    if wl == "auto":                 # the sheet north_star names for this GPU count
        wl = "cfg2" if world == 1 else "cfg3"
    cfg = 2 if wl == "weak2" else int(wl[3])
    seed, palrgb, stats, spec = synth.config_workload(cfg)
    spec = dict(spec)
    if wl == "weak2" and world > 1:
        spec["n2"] = spec["n2"] * world          # weak scaling: one config-2 sheet per GPU
        name = (f"weak: {spec['n1']}x{spec['n2']} sheet = {world} row bands of BASELINE config 2 "
                f"(20000x20000, K=16, window 5), R=2 halo lines per band edge (see config.halo)")
        scaling = "weak"
    else:
        names = {1: "BASELINE config 1: 2000x2000, K=8, window 3",
                 2: "BASELINE config 2: 20000x20000 sheet, K=16, clean window 5",
                 3: "BASELINE config 3: 60000x40000 sheet (2.4 Gpx), K=16, window 5, row bands + halo lines (see config.halo)",
                 4: "BASELINE config 4: 20000x20000, K=64 Lab distance + box, window 7",
                 5: "BASELINE config 5: batch of 256 sheets 4096x4096, K=16, window 5, whole sheets per GPU"}
        name = names[cfg]
        scaling = "strong"    # total work fixed: the driver's N = 1 point is config 2, N > 1 is config 3 (same Mpx/s metric)
    return cfg, seed, palrgb, stats, spec, name, scaling


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's CPU path for this hot path.  The reference itself is Julia (not installed
    in this image), so this arm times oracle/ -- the C restatement of
    src/categories.jl:76-133 + Appendix A.3 -- with all host threads, on a bounded sample of the
    same workload."""
    if rank != 0:
        return
    # every host core, whatever the launcher exported: torch.distributed.run sets OMP_NUM_THREADS=1
    # for its workers, which would time one thread against N GPUs
    cores = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(cores)
    import numpy as np
    from oracle import oracle as O
    cfg, seed, palrgb, stats, spec, name, scaling = workload_spec(args, max(args.gpus, 1))
    cs = O.LAB if stats.colourspace == "lab" else O.RGB
    n1 = spec["n1"]
    run = lambda sc: O.categorize_clean(sc, stats.mean, stats.lo, stats.hi, spec["R"], cs, nthreads=cores)  # noqa: E731
    # size the per-step sample so that warmup+steps stay within ~2 minutes (same rule at every N)
    probe_lines = max(8, min(spec["n2"], (1 << 21) // n1))
    scan = O.synth_scan(seed, 0, palrgb, n1, spec["n2"], 0, probe_lines, nthreads=cores)
    run(scan)
    t0 = time.perf_counter()
    run(scan)
    rate = n1 * probe_lines / (time.perf_counter() - t0)        # px/s
    budget = 90.0 / max(1, args.steps + args.warmup)
    lines = int(min(spec["n2"], max(probe_lines, rate * min(budget, 10.0) / n1)))
    scan = O.synth_scan(seed, 0, palrgb, n1, spec["n2"], 0, lines, nthreads=cores)
    for _ in range(args.warmup):
        run(scan)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(scan)
    dt = (time.perf_counter() - t0) / args.steps
    mpx = n1 * lines / dt / 1e6
    sample = f"first {lines} lines ({n1 * lines / 1e6:.1f} Mpx) of the sheet per step, {cores} OpenMP threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": mpx, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "K": spec["K"], "R": spec["R"], "colourspace": stats.colourspace},
        "cpu_baseline": {"value": mpx, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mpx, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Julia (no toolchain here): oracle/ C restatement timed on host cores",
    }))


# ------------------------------------------------------------------------------------ our arm
def measure_config(mr, cfg, dev, peak, reps=5, sheets=None):
    """Fused kernel on BASELINE config `cfg`, one GPU, device-resident input: dict with value + roofline.
    Its own context (own palette); buffers are freed on return.  sheets: config-5 share of this GPU."""
    import torch
    from maprast_b200 import synth
    seed, palrgb, stats, spec = synth.config_workload(cfg)
    n1, n2, R, K = spec["n1"], spec["n2"], spec["R"], spec["K"]
    s0, s1 = sheets if sheets else (0, spec["n_sheets"])
    ns = s1 - s0
    ctx = mr.Context(dev.index)
    try:
        ctx.set_palette(stats.mean, stats.lo, stats.hi, mr._lib.LAB if stats.colourspace == "lab" else mr._lib.RGB)
        px = torch.empty((ns, n2, n1, 3), dtype=torch.uint8, device=dev)
        for s in range(ns):
            ctx.synth_scan_dev(seed, s0 + s, palrgb, n1, n2, 0, n2, px[s], n1 * 3)
        out = torch.empty((ns, n2, n1), dtype=torch.uint8, device=dev)

        def step():
            if ns > 1:
                ctx.batch_categorize_clean_dev(ns, px, n1 * n2 * 3, n1, n2, n1 * 3, 3, R, out, n1 * n2, n1)
            else:
                ctx.categorize_clean_dev(px[0], n1, n2, n1 * 3, 3, R, out[0], n1)
        npx = ns * n1 * n2
        small = npx * ALGO_BYTES_PER_PX < (512 << 20)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            if small:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        med = statistics.median(ms)
        ach = ALGO_BYTES_PER_PX * npx / (med * 1e-3) / 1e9
        return {"what": f"BASELINE config {cfg}: {ns} x {n1}x{n2}, K={K}, window {2 * R + 1}, {stats.colourspace}, device resident, one launch per step",
                "value": npx / (med * 1e-3) / 1e6, "unit": UNIT, "ms_median": med, "ms_min": min(ms), "pixels": npx,
                "palette_table_build_ms": ctx.palette_build_ms(),
                "cache": "L2 flushed before every step" if small else "input larger than L2",
                "roofline": {"bound": "hbm", "algorithmic_bytes_per_px": ALGO_BYTES_PER_PX, "achieved": ach, "peak": peak,
                             "unit": "GB/s", "frac": ach / peak}}
    finally:
        ctx.close()


def copy_ceiling(host_in, host_out, dev, reps=3):
    """Pure pinned H2D || D2H of the e2e step's bytes on two streams of this rank (no kernel): seconds per step."""
    import torch
    d_in = torch.empty_like(host_in, device=dev)
    d_out = torch.empty_like(host_out, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    best = None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_in.copy_(host_in, non_blocking=True)
        with torch.cuda.stream(s2):
            host_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import maprast_b200 as mr
    D = mr.distributed

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = mr.Context(local_rank)
    cfg, seed, palrgb, stats, spec, name, scaling = workload_spec(args, world)
    n1, n2, R, K = spec["n1"], spec["n2"], spec["R"], spec["K"]
    cs_id = mr._lib.LAB if stats.colourspace == "lab" else mr._lib.RGB
    ctx.set_palette(stats.mean, stats.lo, stats.hi, cs_id)
    build_ms_first = ctx.palette_build_ms()          # includes the one-time module load of the build kernels
    ctx.set_palette(stats.mean, stats.lo, stats.hi, cs_id)
    build_ms = ctx.palette_build_ms()                # what a palette change costs (slider move -> mr_set_palette)
    batch = spec["n_sheets"] > 1

    # ---- synthetic input, generated in place on the device (untimed)
    if batch:
        s0, s1 = D.sheet_range(spec["n_sheets"], world, rank)
        ns = s1 - s0
        px = torch.empty((ns, n2, n1, 3), dtype=torch.uint8, device=dev)
        for s in range(ns):
            ctx.synth_scan_dev(seed, s0 + s, palrgb, n1, n2, 0, n2, px[s], n1 * 3)
        out = torch.empty((ns, n2, n1), dtype=torch.uint8, device=dev)
        my_px = ns * n1 * n2
        total_px = spec["n_sheets"] * n1 * n2
    else:
        a, b = D.band_range(n2, world, rank)
        nb = b - a
        buf, own = D.alloc_band(n1, nb, R, device=dev)
        ctx.synth_scan_dev(seed, 0, palrgb, n1, n2, a, nb, own, n1 * 3)
        out = torch.empty((nb, n1), dtype=torch.uint8, device=dev)
        my_px = nb * n1
        total_px = n1 * n2
    torch.cuda.synchronize()
    peer = None
    halo_mode = "none"
    if world > 1 and not batch:
        halo_mode = "nccl send/recv of R lines per neighbour before every launch"
        if args.halo in ("auto", "peer"):
            try:
                peer = D.PeerHalos(ctx, buf, R, rank, world)
                halo_mode = "in place: TMA line copies read the neighbour band's HBM over NVLink (CUDA IPC peer memory)"
            except Exception as e:   # noqa: BLE001
                if args.halo == "peer":
                    raise
                peer = None
                halo_mode += f" (peer mapping unavailable: {type(e).__name__})"
        # all ranks must take the same path
        flag = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and peer is not None:
            peer.close()
            peer = None
            halo_mode = "nccl send/recv of R lines per neighbour before every launch"
        D.exchange_halos(buf, R, rank, world)   # halo slots valid for the host-buffer (e2e) leg either way

    def step():
        if batch:
            ctx.batch_categorize_clean_dev(ns, px, n1 * n2 * 3, n1, n2, n1 * 3, 3, R, out, n1 * n2, n1)
        elif peer is not None:
            peer.categorize_clean(buf, out)
        elif world > 1:
            D.categorize_clean_band(ctx, buf, R, out, rank, world)
        else:
            ctx.categorize_clean_dev(own, n1, nb, n1 * 3, 3, R, out, n1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    nvml = NvmlSampler(local_rank) if rank == 0 else None
    if rank == 0:
        sampler.start()
        nvml.start()
    launches0 = ctx.launch_count()
    # small workloads would sit in the 126 MB L2 between iterations: flush it (write a 256 MiB buffer)
    # before every timed step and time the steps one by one; big ones are larger than L2 by themselves
    small = my_px * ALGO_BYTES_PER_PX < (512 << 20)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None
    barrier()
    if small:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k in range(args.steps):
            flush_buf.zero_()
            ev[k][0].record()
            step()
            ev[k][1].record()
        barrier()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        total_ms = sum(step_ms)
    else:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        ev[0].record()
        for k in range(args.steps):
            step()
            ev[k + 1].record()
        barrier()
        total_ms = ev[0].elapsed_time(ev[-1])
        step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        fast = nvml.stop()
        if fast and fast["samples"]:
            # the millisecond samples are the ones taken under load; nvidia-smi's line stays beside them
            clocks = {"sm_mhz": fast["sm_mhz"], "sm_min_mhz": fast["sm_min_mhz"], "sm_max_mhz": fast["sm_max_mhz"],
                      "samples": fast["samples"], "reasons": sorted(set(fast["reasons"]) | set(clocks.get("reasons") or [])),
                      "source": "NVML polled every ~0.5 ms over the timed region", "nvidia_smi": clocks}
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = total_px / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: public host-buffer API, pinned host memory, H2D + D2H inside the timed region
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    if batch:
        host_in = torch.empty((ns * n2, n1, 3), dtype=torch.uint8, pin_memory=True)
        host_in.copy_(px.view(ns * n2, n1, 3))
        e_n2 = n2   # per-sheet calls below
    else:
        # the host holds the sheet: each rank uploads its band plus the halo lines straight from it
        hl, hh = (min(R, a), min(R, n2 - b))
        host_in = torch.empty((hl + nb + hh, n1, 3), dtype=torch.uint8, pin_memory=True)
        host_in.copy_(buf[R - hl:R + nb + hh])     # halos are valid after the warm-up exchanges
    host_out = torch.empty((ns * n2 if batch else nb, n1), dtype=torch.uint8, pin_memory=True)
    hin, hout = host_in.numpy(), host_out.numpy()

    def e2e_step():
        if batch:
            for s in range(ns):
                ctx.categorize_clean_host(hin[s * n2:(s + 1) * n2], R, out=hout[s * n2:(s + 1) * n2])
        else:
            ctx.categorize_clean_host(hin, R, out=hout, halo_lo=hl, halo_hi=hh)
        return int(hout[-1, -1])   # the step's result is read on the host

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_px / float(t.item()) / 1e6
    # e2e labels must equal the device-resident labels (same code path, same bytes)
    same = bool(torch.equal(host_out.to(dev).view(-1), out.view(-1)))
    # what the host side alone allows: the same bytes over the same kind of streams, no kernel
    barrier()
    t = torch.tensor([copy_ceiling(host_in, host_out, dev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ceiling = total_px / float(t.item()) / 1e6
    del host_in, host_out, hin, hout
    peak, peak_src = measured_peak()
    # ---- the other BASELINE configs beside the headline one (same kernel, same call)
    extra = {}
    if world == 1:
        for c in (1, 3, 4, 5):
            if c != cfg:
                extra[f"cfg{c}"] = measure_config(mr, c, dev, peak)
    elif cfg != 5:   # batch of 256 sheets, whole sheets per rank, no halo (every rank takes part)
        sh = D.sheet_range(256, world, rank)
        r5 = measure_config(mr, 5, dev, peak, sheets=sh)
        t = torch.tensor([r5["ms_median"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        px5 = 256 * 4096 * 4096
        r5["value"] = px5 / (float(t.item()) * 1e-3) / 1e6
        r5["what"] = f"BASELINE config 5: 256 sheets 4096x4096 over {world} GPUs ({sh[1] - sh[0]} per GPU), no halo; value = all sheets / slowest rank"
        r5["roofline"]["note"] = "rank 0's share against one GPU's peak"
        extra["cfg5"] = r5

    if rank != 0:
        return
    kern_ms = statistics.mean(step_ms)      # one fused launch per step on this stream
    launches_per_step = launches / args.steps
    achieved = ALGO_BYTES_PER_PX * my_px / (kern_ms * 1e-3) / 1e9
    wl_key = args.workload if args.workload != "auto" else "cfg2"
    res = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "ms_per_step_min": min(step_ms),
        "ms_per_step_median": statistics.median(step_ms), "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": name, "K": K, "R": R, "colourspace": stats.colourspace,
                   "pixels_total": total_px, "pixels_per_gpu": my_px,
                   "cache": ("L2 flushed (256 MiB written) before every timed step; steps timed one by one"
                             if small else "inputs larger than L2 (%.2f GB per GPU per step)" % (my_px * 4 / 1e9)),
                   "palette_table_build_ms": build_ms, "palette_table_build_ms_first_call": build_ms_first,
                   "halo": halo_mode},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0,
                     "peak_source": peak_src, "kernel": "fused categorise+clean", "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX * my_px,
                     "traffic": ncu_traffic(wl_key) if world == 1 else None},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(total_px * 3),
                "d2h_bytes_per_step": int(total_px), "steps": e2e_steps, "labels_equal_device_path": same,
                "ceiling": {"value": e2e_ceiling, "unit": UNIT,
                            "what": "pinned H2D || D2H of the same bytes on two streams per rank, no kernel (max over ranks): "
                                    "the host link, not the kernel, bounds the end-to-end number",
                            "frac_of_ceiling": e2e_value / e2e_ceiling}},
        "gpu_launches": int(launches), "gpu_launches_per_step": launches_per_step,
        "clocks": clocks,
    }
    # ---- diagnostics SURVEY 8(d) asks for: one counted launch outside the timed region (MIXED-cell =
    # "near-tie" pixels resolved through the exact table, no-match labels, labels the filter changed)
    # and the throughput with the per-palette table build NOT amortised over many sheets
    ctx.enable_counters(True)
    ctx.read_counters(reset=True)
    step()
    torch.cuda.synchronize()
    cnt = ctx.read_counters(reset=True)
    ctx.enable_counters(False)
    res["diagnostics"] = {
        "counted_pixels": cnt["n_pixels"], "exact_table_pixels": cnt["n_fine"], "nomatch_pixels": cnt["n_nomatch"],
        "changed_by_clean_pixels": cnt["n_changed"],
        "value_unamortised_table_build": my_px / ((kern_ms + build_ms) * 1e-3) / 1e6,
        "note": "counters of rank 0's share; table build = 2^24 exact Float64 evaluations per palette, once per mr_set_palette"}
    res["next_rows"] = dict(extra)
    if world == 1 and not batch and n1 * nb < (1 << 31):
        # ---- SURVEY 8(f) "next" row, measured beside the hot path: the reference's segment prune
        # (_clean, src/clean.jl:59-93) on the labels the fused kernel just produced (device resident)
        keep = np.zeros(256, dtype=np.uint8)
        keep[1:K + 1] = 1
        seg_out = torch.empty_like(out)
        nseg = ctx.clean_segments_dev(out, n1, nb, n1, keep, 0.01, 20, seg_out, n1)
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            ctx.clean_segments_dev(out, n1, nb, n1, keep, 0.01, 20, seg_out, n1)
        e1.record()
        torch.cuda.synchronize()
        seg_ms = e0.elapsed_time(e1) / reps
        res["next_rows"]["clean_segments"] = {
            "what": "_clean(labels; categories=1..K, line_threshold=0.01, prune_threshold=20), device resident",
            "value": my_px / (seg_ms * 1e-3) / 1e6, "unit": UNIT, "ms": seg_ms, "segments": nseg,
            "gpu_launches_per_call": (ctx.launch_count() - l0) / reps,
            "roofline": {"bound": "hbm", "algorithmic_bytes_per_px": 2,
                         "achieved": 2 * my_px / (seg_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": 2 * my_px / (seg_ms * 1e-3) / 1e9 / peak}}
        # fill_gaps (src/clean.jl:2-54, the "fill" button) on the pruned labels, device resident, one pass
        cats = list(range(1, K + 1))
        fill_out = torch.empty_like(out)
        if not batch and stats.colourspace == "rgb":
            ctx.fill_gaps_dev(seg_out, n1, nb, n1, own, n1 * 3, 3, cats, fill_out, n1)
            torch.cuda.synchronize()
            l0 = ctx.launch_count()
            e0.record()
            for _ in range(reps):
                ctx.fill_gaps_dev(seg_out, n1, nb, n1, own, n1 * 3, 3, cats, fill_out, n1)
            e1.record()
            torch.cuda.synchronize()
            fg_ms = e0.elapsed_time(e1) / reps
            res["next_rows"]["fill_gaps"] = {
                "what": "fill_gaps(cleaned; categories=1..K, VonNeumann{2}, original = the scan), one pass, device resident",
                "value": my_px / (fg_ms * 1e-3) / 1e6, "unit": UNIT, "ms": fg_ms,
                "gap_pixels_in": int((seg_out == 0).sum().item()), "gap_pixels_out": int((fill_out == 0).sum().item()),
                "gpu_launches_per_call": (ctx.launch_count() - l0) / reps,
                "roofline": {"bound": "hbm", "algorithmic_bytes_per_px": 2,
                             "achieved": 2 * my_px / (fg_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": 2 * my_px / (fg_ms * 1e-3) / 1e9 / peak,
                             "note": "label in + label out; the original pixel (3 B) is only read where two categories tie"}}
            # the same sheet with 5 % of the labels knocked out at random (a pruned real map has gaps; the pruned
            # synthetic sheet has almost none): the stencil path of the kernel
            gappy = seg_out.clone()
            gen = torch.Generator(device=dev).manual_seed(5)
            gappy[torch.rand(gappy.shape, device=dev, generator=gen) < 0.05] = 0
            ctx.fill_gaps_dev(gappy, n1, nb, n1, own, n1 * 3, 3, cats, fill_out, n1)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                ctx.fill_gaps_dev(gappy, n1, nb, n1, own, n1 * 3, 3, cats, fill_out, n1)
            e1.record()
            torch.cuda.synchronize()
            fg5_ms = e0.elapsed_time(e1) / reps
            res["next_rows"]["fill_gaps"]["with_5pct_gaps"] = {
                "value": my_px / (fg5_ms * 1e-3) / 1e6, "ms": fg5_ms, "gap_pixels_in": int((gappy == 0).sum().item()),
                "gap_pixels_out": int((fill_out == 0).sum().item()),
                "roofline_frac": 2 * my_px / (fg5_ms * 1e-3) / 1e9 / peak}
            del gappy
            if not args.no_cpu_baseline:
                from oracle import oracle as O
                lines = max(64, min(nb, (1 << 23) // n1))
                t0 = time.perf_counter()
                want_f = O.fill_gaps(seg_out[:lines].cpu().numpy(), own[:lines].cpu().numpy(), stats.mean, cats)
                dt = time.perf_counter() - t0
                got_f = torch.empty((lines, n1), dtype=torch.uint8, device=dev)
                ctx.fill_gaps_dev(seg_out[:lines], n1, lines, n1, own[:lines], n1 * 3, 3, cats, got_f, n1)
                res["next_rows"]["fill_gaps"]["cpu_baseline"] = {
                    "value": n1 * lines / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"first {lines} lines, sequential C oracle",
                    "label_mismatches_vs_gpu": int((got_f.cpu().numpy() != want_f).sum())}
        del fill_out
        # the reference's own button sequence (it has no majority filter): categorize -> "clean" (segment prune on
        # the RAW labels: millions of speckle segments) -> "fill" (one pass), device resident, whole sheet
        if not batch and stats.colourspace == "rgb":
            raw = torch.empty_like(out)
            pruned = torch.empty_like(out)
            filled = torch.empty_like(out)
            cats = list(range(1, K + 1))
            def _pipeline():
                ctx.categorize_clean_dev(own, n1, nb, n1 * 3, 3, 0, raw, n1)
                ns = ctx.clean_segments_dev(raw, n1, nb, n1, keep, 0.01, 20, pruned, n1)
                ctx.fill_gaps_dev(pruned, n1, nb, n1, own, n1 * 3, 3, cats, filled, n1)
                return ns
            nseg_raw = _pipeline()
            torch.cuda.synchronize()
            l0 = ctx.launch_count()
            e0.record()
            for _ in range(reps):
                _pipeline()
            e1.record()
            torch.cuda.synchronize()
            pl_ms = e0.elapsed_time(e1) / reps
            res["next_rows"]["reference_pipeline"] = {
                "what": "categorize (no filter) -> _clean segment prune on the raw labels -> fill_gaps, one pass each, device resident",
                "value": my_px / (pl_ms * 1e-3) / 1e6, "unit": UNIT, "ms": pl_ms, "segments": nseg_raw,
                "gap_pixels_after_prune": int((pruned == 0).sum().item()), "gap_pixels_after_fill": int((filled == 0).sum().item()),
                "gpu_launches_per_call": (ctx.launch_count() - l0) / reps}
            del raw, pruned, filled
        # blur (src/blur.jl:6-14) and the categoriser on its RGB{Float64} output, on the first lines of the sheet
        # (24 B per pixel of Float64 colour: 4096 lines = 2 GB)
        if not batch and stats.colourspace == "rgb":
            bl = min(nb, 4096)
            f64 = torch.empty((bl, n1, 3), dtype=torch.float64, device=dev)
            lab64 = torch.empty((bl, n1), dtype=torch.uint8, device=dev)
            L = ctx._L
            def _blur():
                ctx._ck(L.mr_blur_dev(ctx._h, own.data_ptr(), n1, bl, n1 * 3, 3, 1, 1, f64.data_ptr(), n1 * 24, None))
            def _cat():
                ctx._ck(L.mr_categorize_clean_f64_dev(ctx._h, f64.data_ptr(), n1, bl, n1 * 24, 0, lab64.data_ptr(), n1, None))
            for fn, key, what, bpp_alg in ((_blur, "blur", "blur(A; hood=Window{1}(), repeat=1): RGB{N0f8} in, RGB{Float64} out", 27),
                                           (_cat, "categorize_f64", "categorise RGB{Float64} pixels (direct Float64 evaluation)", 25)):
                fn()
                torch.cuda.synchronize()
                l0 = ctx.launch_count()
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                res["next_rows"][key] = {
                    "what": what + f", first {bl} lines, device resident", "value": n1 * bl / (ms * 1e-3) / 1e6, "unit": UNIT,
                    "ms": ms, "gpu_launches_per_call": (ctx.launch_count() - l0) / reps,
                    "roofline": {"bound": "hbm" if key == "blur" else "fp64 issue (HBM figure for reference)",
                                 "algorithmic_bytes_per_px": bpp_alg, "achieved": bpp_alg * n1 * bl / (ms * 1e-3) / 1e9,
                                 "peak": peak, "unit": "GB/s", "frac": bpp_alg * n1 * bl / (ms * 1e-3) / 1e9 / peak}}
            # texture features of the default match = (:color, :std, :stripe) on the blurred image
            # (_stds src/selectcolors.jl:706-715, _stripes src/stripes.jl:20-35, PixelProp categoriser)
            sdp = torch.empty((bl, n1), dtype=torch.float64, device=dev)
            stp = torch.empty((bl, n1, 2), dtype=torch.float64, device=dev)
            K_ = stats.K
            z1, z2 = np.full(K_, 0.5), np.zeros((K_, 2))
            ctx.set_texture_stats(z1, z1 - 0.6, z1 + 0.6, z2, z2 - 50.0, z2 + 50.0)
            def _stds():
                ctx._ck(L.mr_stds_dev(ctx._h, f64.data_ptr(), n1, bl, n1 * 24, sdp.data_ptr(), n1 * 8, None))
            def _stripes():
                ctx._ck(L.mr_stripes_dev(ctx._h, f64.data_ptr(), n1, bl, n1 * 24, 2, stp.data_ptr(), n1 * 16, None))
            def _props():
                ctx._ck(L.mr_categorize_clean_props_dev(ctx._h, f64.data_ptr(), n1 * 24, sdp.data_ptr(), n1 * 8, stp.data_ptr(),
                                                        n1 * 16, n1, bl, 7, 0, lab64.data_ptr(), n1, None))
            for fn, key, what, bpp_alg in ((_stds, "stds", "_stds(blurred): RGB{Float64} in, Float64 plane out", 32),
                                           (_stripes, "stripes", "_stripes(blurred; radius=2): RGB{Float64} in, 2 x Float64 out", 40),
                                           (_props, "categorize_props", "categorise PixelProp pixels, match = (:color, :std, :stripe)", 49)):
                fn()
                torch.cuda.synchronize()
                l0 = ctx.launch_count()
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                res["next_rows"][key] = {
                    "what": what + f", first {bl} lines, device resident", "value": n1 * bl / (ms * 1e-3) / 1e6, "unit": UNIT,
                    "ms": ms, "gpu_launches_per_call": (ctx.launch_count() - l0) / reps,
                    "roofline": {"bound": "fp64 issue (HBM figure for reference)",
                                 "algorithmic_bytes_per_px": bpp_alg, "achieved": bpp_alg * n1 * bl / (ms * 1e-3) / 1e9,
                                 "peak": peak, "unit": "GB/s", "frac": bpp_alg * n1 * bl / (ms * 1e-3) / 1e9 / peak}}
            # the reference's DEFAULT settings end to end on those lines (blur_repeat = 1, match = (:color, :std, :stripe),
            # stripe_radius = 2): blur -> _stds -> _stripes -> PixelProp categorise -> _clean -> fill_gaps, device resident
            pr64 = torch.empty((bl, n1), dtype=torch.uint8, device=dev)
            fl64 = torch.empty((bl, n1), dtype=torch.uint8, device=dev)
            cats64 = list(range(1, K + 1))
            def _default_pipeline():
                _blur(); _stds(); _stripes(); _props()
                ctx.clean_segments_dev(lab64, n1, bl, n1, keep, 0.01, 20, pr64, n1)
                ctx.fill_gaps_dev(pr64, n1, bl, n1, own, n1 * 3, 3, cats64, fl64, n1)
            _default_pipeline()
            torch.cuda.synchronize()
            l0 = ctx.launch_count()
            e0.record()
            for _ in range(reps):
                _default_pipeline()
            e1.record()
            torch.cuda.synchronize()
            dp_ms = e0.elapsed_time(e1) / reps
            res["next_rows"]["default_pipeline"] = {
                "what": "the reference's default settings: blur -> _stds -> _stripes -> categorise (:color, :std, :stripe) -> _clean -> "
                        f"fill_gaps, first {bl} lines, device resident",
                "value": n1 * bl / (dp_ms * 1e-3) / 1e6, "unit": UNIT, "ms": dp_ms,
                "gpu_launches_per_call": (ctx.launch_count() - l0) / reps}
            del f64, lab64, sdp, stp, pr64, fl64
        # polygon mask of the "clean" / "fill" buttons (src/selectcolors.jl:559-563): 16 rings of 12 vertices over the sheet
        if not batch:
            prng = np.random.default_rng(17)
            rings_xy, offs = [], [0]
            for _ in range(16):
                c = prng.random(2) * (n1, nb)
                ang = np.sort(prng.random(12)) * 2 * np.pi
                rad = (0.05 + 0.15 * prng.random(12)) * min(n1, nb)
                rings_xy.append(np.stack([c[0] + rad * np.cos(ang), c[1] + rad * np.sin(ang)], axis=1))
                offs.append(offs[-1] + 12)
            pxy = np.ascontiguousarray(np.concatenate(rings_xy).astype(np.float32).astype(np.float64))
            poff = np.asarray(offs, dtype=np.int32)
            pmask = torch.empty((nb, n1), dtype=torch.uint8, device=dev)
            def _poly():
                ctx._ck(ctx._L.mr_polygons_dev(ctx._h, 16, poff.ctypes.data, pxy.ctypes.data, None, n1, nb, pmask.data_ptr(), n1, None))
            _poly()
            torch.cuda.synchronize()
            l0 = ctx.launch_count()
            e0.record()
            for _ in range(reps):
                _poly()
            e1.record()
            torch.cuda.synchronize()
            pm_ms = e0.elapsed_time(e1) / reps
            res["next_rows"]["polygons"] = {
                "what": "_polymask: 16 rings of 12 vertices over the sheet, device resident (per call: rings packed and uploaded, one kernel)",
                "value": my_px / (pm_ms * 1e-3) / 1e6, "unit": UNIT, "ms": pm_ms, "inside_fraction": float(pmask.float().mean().item()),
                "gpu_launches_per_call": (ctx.launch_count() - l0) / reps,
                "roofline": {"bound": "hbm", "algorithmic_bytes_per_px": 1, "achieved": my_px / (pm_ms * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": my_px / (pm_ms * 1e-3) / 1e9 / peak}}
            del pmask
        # _shrink_grow (src/clean.jl:132-140) on the fused labels: one _erode pass, one _dilate pass, and _shrink_grow(A, 1, 1)
        # (the dilation then meets the plane the erosion has emptied: the counting rule runs for every rim pixel)
        if not batch:
            sg = torch.empty_like(out)
            sg_ms = {}
            for key, (ne, nd) in (("erode", (1, 0)), ("dilate", (0, 1)), ("erode_then_dilate", (1, 1))):
                def _sg():
                    ctx._ck(ctx._L.mr_shrink_grow_dev(ctx._h, out.data_ptr(), n1, nb, n1, ne, nd, sg.data_ptr(), n1, None))
                _sg()
                torch.cuda.synchronize()
                l0 = ctx.launch_count()
                e0.record()
                for _ in range(reps):
                    _sg()
                e1.record()
                torch.cuda.synchronize()
                sg_ms[key] = e0.elapsed_time(e1) / reps
            res["next_rows"]["shrink_grow"] = {
                "what": "_erode / _dilate / _shrink_grow(labels, 1, 1) over the fused labels, device resident; value and roofline: one _erode pass",
                "value": my_px / (sg_ms["erode"] * 1e-3) / 1e6, "unit": UNIT, "ms": sg_ms["erode"], "ms_dilate": sg_ms["dilate"],
                "ms_erode_then_dilate": sg_ms["erode_then_dilate"],
                "roofline": {"bound": "hbm", "algorithmic_bytes_per_px": 2, "achieved": 2 * my_px / (sg_ms["erode"] * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": 2 * my_px / (sg_ms["erode"] * 1e-3) / 1e9 / peak}}
            del sg
        # the standalone Window majority filter, clean(labels; hood): categorise-only labels in, device resident
        pre = torch.empty_like(out)
        ctx.categorize_clean_dev(own, n1, nb, n1 * 3, 3, 0, pre, n1)
        two = torch.empty_like(out)
        ctx.clean_dev(pre, n1, nb, n1, R, two, n1)
        torch.cuda.synchronize()
        l0 = ctx.launch_count()
        e0.record()
        for _ in range(reps):
            ctx.clean_dev(pre, n1, nb, n1, R, two, n1)
        e1.record()
        torch.cuda.synchronize()
        cl_ms = e0.elapsed_time(e1) / reps
        res["next_rows"]["clean_standalone"] = {
            "what": "clean(labels; hood=Window{R}) on categorise-only labels, device resident (label-maximum pass + vote kernel)",
            "value": my_px / (cl_ms * 1e-3) / 1e6, "unit": UNIT, "ms": cl_ms,
            "gpu_launches_per_call": (ctx.launch_count() - l0) / reps,
            "equal_to_fused": bool(torch.equal(two, out)),
            "roofline": {"bound": "hbm", "algorithmic_bytes_per_px": 2,
                         "achieved": 2 * my_px / (cl_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": 2 * my_px / (cl_ms * 1e-3) / 1e9 / peak}}
        e0.record()
        for _ in range(reps):
            ctx.clean_dev(pre, n1, nb, n1, R, two, n1, max_label=K)
        e1.record()
        torch.cuda.synchronize()
        clb_ms = e0.elapsed_time(e1) / reps
        res["next_rows"]["clean_standalone"]["bounded"] = {
            "what": "same through mr_clean_dev_bounded (label bound = K from the caller: one launch, no synchronisation)",
            "value": my_px / (clb_ms * 1e-3) / 1e6, "ms": clb_ms, "equal_to_fused": bool(torch.equal(two, out)),
            "roofline_frac": 2 * my_px / (clb_ms * 1e-3) / 1e9 / peak}
        del pre, two
        if not args.no_cpu_baseline:
            from oracle import oracle as O
            lines = max(64, min(nb, (1 << 24) // n1))
            crop = out[:lines].cpu().numpy()
            t0 = time.perf_counter()
            want_seg, _ = O.clean_segments(crop, None, range(1, K + 1), 0.01, 20)
            dt = time.perf_counter() - t0
            got_seg = torch.empty((lines, n1), dtype=torch.uint8, device=dev)
            ctx.clean_segments_dev(out[:lines], n1, lines, n1, keep, 0.01, 20, got_seg, n1)
            res["next_rows"]["clean_segments"]["cpu_baseline"] = {
                "value": n1 * lines / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"first {lines} lines of the label sheet, sequential C oracle (the reference algorithm is sequential)",
                "label_mismatches_vs_gpu": int((got_seg.cpu().numpy() != want_seg).sum())}
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle as O
        cs = O.LAB if stats.colourspace == "lab" else O.RGB
        cores = O.max_threads()
        src = px[0] if batch else own
        probe = max(8, (1 << 21) // n1)
        scan = src[:probe].contiguous().cpu().numpy()
        t0 = time.perf_counter()
        O.categorize_clean(scan, stats.mean, stats.lo, stats.hi, R, cs)
        rate = scan.shape[0] * n1 / (time.perf_counter() - t0)
        lines = int(min(src.shape[0], max(probe, rate * args.cpu_seconds / 2 / n1)))
        scan = src[:lines].contiguous().cpu().numpy()
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            want = O.categorize_clean(scan, stats.mean, stats.lo, stats.hi, R, cs)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        # while the oracle output is at hand: the device labels of those lines must match it
        # (lines whose window reaches past the sample are excluded)
        got = (out[0] if batch else out)[:lines].cpu().numpy()
        cut = lines if lines == src.shape[0] else lines - R
        res["cpu_baseline"] = {"value": n1 * lines / best / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"first {lines} lines ({n1 * lines / 1e6:.1f} Mpx) of the same sheet, "
                                         f"best of 2, C oracle + OpenMP ({cores} threads)",
                               "label_mismatches_vs_gpu": int((got[:cut] != want[:cut]).sum())}
    print(json.dumps(res))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
