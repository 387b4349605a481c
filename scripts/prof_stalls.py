"""Warp-stall samples per source line (ncu SASS page csv + nvdisasm line info)."""
import csv, re, sys, collections
sass_csv, disasm, kern, srcfile = sys.argv[1:5]
cur_fn = None; cur_line = None; addr2line = {}; addr2op = {}
fn_re = re.compile(r"^\s*\.text\.(\S+):")
line_re = re.compile(r'//## File "([^"]+)", line (\d+)')
ins_re = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")
for ln in open(disasm):
    m = fn_re.match(ln)
    if m: cur_fn = m.group(1); continue
    m = line_re.search(ln)
    if m: cur_line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = ins_re.match(ln)
    if m and cur_fn and kern in cur_fn:
        addr2line[int(m.group(1), 16)] = cur_line; addr2op[int(m.group(1), 16)] = m.group(2).split()[0]
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia = hdr.index("Address"); isamp = hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: collections.Counter()); tot = collections.Counter()
for r in rows[2:]:
    try: a = int(r[ia], 16)
    except Exception: continue
    if base is None: base = a
    key = addr2line.get(a - base, ("?", 0))
    for i, h in stall_cols:
        try: v = int(r[i])
        except Exception: v = 0
        if v:
            agg[key][h] += v; tot[h] += v
src = open(srcfile).read().split("\n")
T = sum(tot.values())
print("stall samples by reason:", {k: f"{v / T * 100:.1f}%" for k, v in tot.most_common(10)})
for key, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:int(sys.argv[5]) if len(sys.argv) > 5 else 30]:
    f, l = key
    text = src[l - 1].strip()[:70] if f.endswith(srcfile.split("/")[-1]) and l > 0 else f
    top = ", ".join(f"{k[6:]}={v}" for k, v in c.most_common(3))
    print(f"{sum(c.values()) / T * 100:5.1f}%  {f}:{l}: {text}   [{top}]")
