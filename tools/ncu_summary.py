"""Condense `ncu -i report.ncu-rep --page raw --csv` into the table kept under profiles/: one column per kernel launch, the rows
that matter for the roofline discussion (DRAM bytes, duration, pipe utilisation, issue slots, stall reasons, launch shape)."""
import csv
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "sm__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_membar.ratio", "smsp__average_warp_latency_issue_stalled_sleeping.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
kcol = names.index("Kernel Name")
want = set(sys.argv[2:])   # optional substrings of kernel names
data = [r for r in data if len(r) == len(names) and (not want or any(w in r[kcol] for w in want))]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [r[kcol].split("(")[0].split("::")[-1] for r in data])
for m in KEEP:
    if m in names:
        c = names.index(m)
        w.writerow([m, units[c]] + [r[c] for r in data])
