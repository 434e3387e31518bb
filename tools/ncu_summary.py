#!/usr/bin/env python
"""tools/ncu_summary.py <report.ncu-rep> <out.csv> -- the handful of raw metrics the profiles/ summaries keep."""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum.per_second", "lts__t_sectors_op_write.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
        "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "inst_executed", "sass__inst_executed_global_stores",
        "sass__inst_executed_shared_loads",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio"]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
lines = ["metric,unit,value"] + [f"{k},{d[k][0]},{d[k][1]}" for k in KEYS if k in d]
open(sys.argv[2], "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
