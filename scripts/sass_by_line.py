"""Static SASS per source line of one kernel: python scripts/sass_by_line.py <sass> <kernel-substr> <file> <lo> <hi> [-v]"""
import re, sys, collections
sass, kern, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
verbose = len(sys.argv) > 6
cur = None; line = None; cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter); seen = False
for ln in open(sass):
    m = re.match(r"^\s*\.text\.(\S+):", ln)
    if m:
        if seen and kern in (cur or ""): break
        cur = m.group(1); seen = kern in cur; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur and kern in cur and line and line[0] == fname and lo <= line[1] <= hi:
        op = m.group(2).split(); op = op[1] if op[0].startswith("@") else op[0]
        cnt[line[1]] += 1; ops[line[1]][op.split(".")[0]] += 1
        if verbose: print(line[1], m.group(2)[:90])
tot = 0
for l in sorted(cnt):
    tot += cnt[l]; print(f"{l:5d} {cnt[l]:4d}  " + " ".join(f"{k}:{v}" for k, v in ops[l].most_common()))
print("total", tot)
