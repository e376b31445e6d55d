#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics the K1 notes quote.  usage: ncu_key.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum ",
        "dram__bytes_write.sum ", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__average_warps_issue_stalled",
        "smsp__inst_executed_op_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "lts__t_sectors_op_read.sum",
        "lts__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed_op_global_ld",
        "l1tex__lsu_writeback_active", "sm__pipe", "smsp__inst_executed_pipe_lsu", "l1tex__t_sector_hit_rate"]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if any(w in h for w in want) and "pct_of_peak_sustained_elapsed" not in h.replace("sm__warps", "") and ".per_second" not in h:
            print("  %-90s %-14s %s" % (h, u, v))
