#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys

def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__issue_active.min.pct_of_peak_sustained_elapsed", "sm__issue_active.max.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "smsp__average_warp_latency_per_inst_issued.ratio"]
    for r in rows[2:]:
        d = dict(zip(rows[0], r))
        for k in want:
            if k in d:
                print("%-70s %s" % (k, d[k]))
        for k, v in d.items():
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                try:
                    if float(v) > 0.15:
                        print("%-70s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall "), v))
                except ValueError:
                    pass
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    cur, hdr, agg = None, None, {}
    def I(x):
        try:
            return int(x)
        except ValueError:
            return 0
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
        elif len(r) > 10 and r[0]:
            try:
                ln = int(r[0])
            except ValueError:
                continue
            a = agg.get((cur, ln), [0, 0, r[1]])
            a[0] += I(r[hdr.index("Instructions Executed")])
            a[1] += I(r[hdr.index("# Samples")])
            agg[(cur, ln)] = a
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print("total warp instructions %d, samples %d" % (ti, ts))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-14s %4d  inst %5.1f%%  samples %5.1f%%  %s" % (k[0][:14], k[1], 100 * v[0] / ti, 100 * v[1] / ts, v[2].strip()[:95]))

if __name__ == "__main__":
    main()
