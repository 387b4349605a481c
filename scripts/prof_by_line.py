"""Join an ncu SASS page (csv) with nvdisasm --print-line-info output: instructions executed per source line."""
import csv, re, sys, collections
sass_csv, disasm, kern, srcfile = sys.argv[1:5]
# parse nvdisasm: track current function and line
cur_fn = None; cur_line = None; addr2line = {}
fn_re = re.compile(r"^\s*\.text\.(\S+):")
line_re = re.compile(r'//## File "([^"]+)", line (\d+)(.*)')
ins_re = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")
inl_re = re.compile(r'inlined at "([^"]+)", line (\d+)')
for ln in open(disasm):
    m = fn_re.match(ln)
    if m: cur_fn = m.group(1); continue
    m = line_re.search(ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = ins_re.match(ln)
    if m and cur_fn and kern in cur_fn:
        addr2line[int(m.group(1), 16)] = cur_line
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, ie, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
isrc = hdr.index("Source")
base = None
agg = collections.defaultdict(lambda: [0, 0]); tot = 0
for r in rows[2:]:
    try:
        a = int(r[ia], 16); e = int(r[ie]); t = int(r[it])
    except Exception:
        continue
    if base is None: base = a
    key = addr2line.get(a - base, ("?", 0))
    agg[key][0] += e; agg[key][1] += t; tot += e
src = open(srcfile).read().split("\n")
print("total warp-instructions", tot)
for key, (e, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 50]:
    f, l = key
    text = src[l - 1].strip()[:90] if f.endswith(srcfile.split("/")[-1]) and l > 0 else f
    print(f"{e / tot * 100:5.1f}% {e:>11d} thr={t / max(e, 1):5.1f}  {f}:{l}: {text}")
