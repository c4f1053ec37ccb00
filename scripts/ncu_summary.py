#!/usr/bin/env python
"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) into the handful of numbers DESIGN.md quotes."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = [k for k in hdr if any(t in k for t in (
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
    "gpu__dram_throughput", "sm__throughput.avg.pct", "launch__registers_per_thread", "launch__occupancy_limit",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
    "warp_issue_stalled", "lts__t_bytes.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg ",
    "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "launch__shared_mem_per_block", "sm__maximum_warps_per_active_cycle_pct",
    "smsp__cycles_active.avg", "launch__waves_per_multiprocessor", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum"))]
for r in rows[2:]:
    for k in keys:
        i = hdr.index(k)
        v = r[i]
        if "stalled" in k and "per_warp_active" not in k:
            continue
        print(f"{k:86s} {v:>22s} {units[i]}")
    print("---")
