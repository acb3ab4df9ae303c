#!/usr/bin/env python
"""Text summary of every kernel in an .ncu-rep (the numbers profiles/*.txt quote).  usage: ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.min.pct_of_peak_sustained_elapsed", "sm__issue_active.max.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]
units = dict(zip(rows[0], rows[1]))
for r in rows[2:]:
    d = dict(zip(rows[0], r))
    print("kernel:", d.get("Kernel Name"))
    for k in want:
        if k in d and d[k] != "":
            print("  %-68s %s %s" % (k, d[k], units.get(k, "")))
    for k, v in sorted(d.items()):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                if float(v) > 0.15:
                    print("  %-68s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall per issue: ").replace("_per_issue_active.ratio", ""), v))
            except ValueError:
                pass
    print()
