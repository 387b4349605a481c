"""Instructions per enclosing function and per pipe class (ncu SASS csv + nvdisasm line info)."""
import csv, re, sys, collections
sass_csv, disasm, kern, srcfile = sys.argv[1:5]
src = open(srcfile).read().split("\n")
# function start lines
fstart = []
for i, l in enumerate(src):
    m = re.match(r"^(?:__device__|__global__|template).*", l)
    m2 = re.search(r"\b(\w+)\s*\(", l)
    if re.match(r"^__device__ __forceinline__ .*?(\w+)\(", l) or re.match(r"^__global__ .*", l):
        name = re.findall(r"(\w+)\(", l)[0] if "__launch_bounds__" not in l else "k_fast(main)"
        fstart.append((i + 1, name))
def func_of(line):
    name = "?"
    for s, n in fstart:
        if s <= line: name = n
        else: break
    return name
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
        a = int(m.group(1), 16); addr2line[a] = cur_line
        op = m.group(2).split()
        op = op[1] if op[0].startswith("@") else op[0]
        addr2op[a] = op
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "IDP", "HFMA")
def pipe(op):
    if op.startswith(FMA): return "fma"
    if op.startswith(("LDS", "STS", "LDG", "STG", "ATOM", "LD.", "ST.", "RED", "LDC", "SHFL", "UBLKCP", "SYNCS", "LDSM")): return "lsu"
    if op.startswith(("POPC", "FLO", "BREV", "MUFU")): return "xu"
    if op.startswith(("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "BAR", "NOP", "CALL", "RET", "YIELD")): return "ctl"
    return "alu"
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
base = None
byf = collections.defaultdict(collections.Counter); tot = collections.Counter()
for r in rows[2:]:
    try: a = int(r[ia], 16); e = int(r[ie])
    except Exception: continue
    if base is None: base = a
    k = addr2line.get(a - base, ("?", 0)); op = addr2op.get(a - base, "?")
    f = func_of(k[1]) if k[0] == srcfile.split("/")[-1] else k[0]
    byf[f][pipe(op)] += e; tot[pipe(op)] += e
T = sum(tot.values())
print("total", T, {k: f"{v / T * 100:.1f}%" for k, v in tot.items()})
for f, c in sorted(byf.items(), key=lambda kv: -sum(kv[1].values())):
    s = sum(c.values())
    if s / T < 0.004: continue
    print(f"{s / T * 100:5.1f}%  {f:28s} " + " ".join(f"{k}={v / T * 100:.1f}" for k, v in c.most_common()))
