#!/bin/bash
# usage: scripts/analyze.sh <tag> [kernel-substr] [nlines]   (here, no GPU): per-source-line instruction counts
TAG=$1; K=${2:-k_fastILi2ELi5ELb0}; N=${3:-45}
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv --print-source sass 2>/dev/null > gpurun_out/$TAG.sass.csv
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/$TAG.raw.csv
mkdir -p /tmp/cub && (cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all "/root/repo/maprasterization.jl_b200/libmaprast.so" > /dev/null 2>&1 && nvdisasm --print-line-info mr_fast.sm_100a.cubin > fast.sass 2>/dev/null)
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/$TAG.raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for h,u,v in zip(hdr,units,vals):
    if h in want: print(f"{h:95s} {u:12s} {v}")
PY
python scripts/prof_by_line.py gpurun_out/$TAG.sass.csv /tmp/cub/fast.sass $K "maprasterization.jl_b200/csrc/mr_fast.cu" $N
