"""ALU-pipe warp instructions per source function and opcode (ncu SASS csv + nvdisasm line info), per strip line.
usage: prof_alu.py sass.csv fast.sass kernel-substr src.cu n_lines"""
import csv, re, sys, collections
sass_csv, disasm, kern, srcfile = sys.argv[1:5]
NL = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
src = open(srcfile).read().split("\n")
fstart = []
for i, l in enumerate(src):
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
line_re = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
ins_re = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")
for ln in open(disasm):
    m = fn_re.match(ln)
    if m: cur_fn = m.group(1); continue
    m = line_re.search(ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = ins_re.match(ln)
    if m and cur_fn and kern in cur_fn:
        a = int(m.group(1), 16); addr2line[a] = cur_line
        op = m.group(2).split()
        op = op[1] if op[0].startswith("@") else op[0]
        addr2op[a] = op
ALU = ("LOP3", "SHF.", "SHF ", "ISETP", "SEL", "VIADD", "IADD3", "PRMT", "PLOP3", "LEA", "R2P", "MOV", "SGXT", "BMSK", "IABS", "VIMNMX", "IMNMX", "CS2R", "P2R", "FSEL", "FSETP", "IADD")
def cls(op):
    if op.startswith("U"): return "uni"
    if op.startswith(("IMAD", "IDP", "FFMA", "FMUL", "FADD", "HFMA")): return "fma"
    if op.startswith(ALU): return "alu"
    if op.startswith(("LDS", "STS", "LDG", "STG", "ATOM", "RED", "LDC", "SHFL", "SYNCS", "LD.", "ST.")): return "lsu"
    if op.startswith(("POPC", "FLO", "BREV", "MUFU")): return "xu"
    return "ctl"
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
base = None
byf = collections.defaultdict(collections.Counter); tot = collections.Counter(); byop = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    try: a = int(r[ia], 16); e = int(r[ie])
    except Exception: continue
    if base is None: base = a
    k = addr2line.get(a - base, ("?", 0)); op = addr2op.get(a - base, "?")
    f = func_of(k[1]) if k[0] == srcfile.split("/")[-1] else k[0]
    c = cls(op)
    byf[f][c] += e; tot[c] += e
    if c == "alu": byop[f][op.split(".")[0]] += e
print("per line:", {k: round(v / NL, 1) for k, v in tot.items()}, "total", round(sum(tot.values()) / NL, 1))
for f, c in sorted(byf.items(), key=lambda kv: -kv[1]["alu"]):
    if sum(c.values()) / NL < 3: continue
    print(f"{f:24s} " + " ".join(f"{k}={c[k] / NL:6.1f}" for k in ("alu", "fma", "lsu", "uni", "xu", "ctl")) + "   | " + " ".join(f"{o}={v / NL:.1f}" for o, v in byop[f].most_common(6)))
