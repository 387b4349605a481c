#!/bin/bash
# usage: scripts/sass_stats.sh [extra nvcc flags]   (no GPU needed)
# Compiles mr_fast.cu (only the BASELINE config 2 instantiation, -DMR_QUICK) and prints registers,
# SASS instruction count and an opcode histogram of k_fast<2,5,false,false>.
set -e
cd "$(dirname "$0")/../maprasterization.jl_b200/csrc"
OUT=/tmp/sass_stats; mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xptxas -v -DMR_QUICK "$@" \
     -cubin -o $OUT/fast.cubin mr_fast.cu 2> $OUT/ptxas.log
grep -A2 "k_fastILi2ELi5ELb0ELb0" $OUT/ptxas.log | grep -E "registers|spill" | head -3
nvdisasm --print-line-info $OUT/fast.cubin > $OUT/fast.sass 2>/dev/null
python3 - <<PY
import re, collections
cur=None; ops=collections.Counter(); n=0
for ln in open('$OUT/fast.sass'):
    m=re.match(r"^\s*\.text\.(\S+):", ln)
    if m: cur=m.group(1); continue
    m=re.match(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur and 'k_fastILi2ELi5ELb0ELb0' in cur:
        op=m.group(2).split(); op=op[1] if op[0].startswith('@') else op[0]
        ops[op.split('.')[0]]+=1; n+=1
print('SASS instructions:', n, '(%.1f KB)' % (n*16/1024))
print(' '.join(f'{k}:{v}' for k,v in ops.most_common(30)))
PY
