"""Dynamic opcode histogram of one kernel from an ncu SASS page csv (ncu -i x.ncu-rep --page source --csv --print-source sass)."""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]
isrc, ie = hdr.index("Source"), hdr.index("Instructions Executed")
NL = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ops = collections.Counter(); tot = 0
for r in rows[2:]:
    try: e = int(r[ie])
    except Exception: continue
    t = r[isrc].split()
    if not t: continue
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += e; tot += e
print("total", tot, "per unit", tot / NL)
for k, v in ops.most_common(40):
    print(f"{k:14s} {v / tot * 100:5.1f}%  {v / NL:8.1f}")
