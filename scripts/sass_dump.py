"""Dump the executed SASS of a kernel from an ncu source page csv with per-instruction counts, in address order.
usage: python scripts/sass_dump.py gpurun_out/TAG.sass.csv [min_exec]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; H = {h: i for i, h in enumerate(hdr)}
mn = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for r in rows[2:]:
    try: e = int(r[H['Instructions Executed']]); t = int(r[H['Thread Instructions Executed']])
    except Exception: continue
    if e < mn: continue
    wf = r[H['L1 Wavefronts Shared']]; 
    print(f"{r[H['Address']][-5:]} {e/1e6:7.2f}M thr={t/max(e,1):4.1f} wf={wf:>9s} {r[H['# Samples']]:>6s}  {r[H['Source']]}")
