"""usage: scripts/summarize_profile.py <tag> [workload-key]   (here, no GPU)
gpurun_out/<tag>.raw.csv (ncu --set full, one k_fast launch; written by scripts/analyze.sh) ->
profiles/r02_ncu_full_k_fast_cfg2.csv (key metrics) and profiles/roofline_traffic.json (bench.py reads it)."""
import csv, json, os, sys
tag = sys.argv[1]; key = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(os.path.join(root, "gpurun_out", tag + ".raw.csv"))))
hdr, units, vals = rows[0], rows[1], rows[2]
want = """Kernel Name|Block Size|Grid Size|dram__bytes_read.sum|dram__bytes_write.sum|dram__cycles_active.avg|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|gpu__time_duration.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|launch__occupancy_limit_registers|launch__occupancy_limit_shared_mem|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|lts__t_sector_hit_rate.pct|sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio|smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio|smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio|smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio|smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio|smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio|smsp__average_warps_issue_stalled_wait_per_issue_active.ratio|smsp__inst_executed.sum|smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed.ratio""".split("|")
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
out = os.path.join(root, "profiles", f"r02_ncu_full_k_fast_{key}.csv")
with open(out, "w", newline="") as f:
    w = csv.writer(f); w.writerow(["metric", "unit", "value"])
    for m in want:
        if m in d: w.writerow([m, d[m][0], d[m][1]])
def gb(m):
    u, v = d[m]; v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
tp = os.path.join(root, "profiles", "roofline_traffic.json")
t = json.load(open(tp)) if os.path.exists(tp) else {}
t[key] = int(rd + wr)
t["_source"] = (f"profiles/r02_ncu_full_k_fast_{key}.csv ({tag}): dram__bytes_read.sum {rd/1e9:.6f} GB + "
                f"dram__bytes_write.sum {wr/1e9:.6f} GB per {d['Kernel Name'][1]} launch")
json.dump(t, open(tp, "w"), indent=1)
print(out, t[key])
