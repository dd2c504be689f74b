#!/usr/bin/env python
"""Summarise ncu output into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv           # per-kernel time shares
    python tools/ncu_summary.py report   gpurun_out/prof_eval.ncu-rep      # key metrics of a --set full capture
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "sm__sass_inst_executed_op_local_ld.sum", "sm__sass_inst_executed_op_local_st.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[row["Metric Unit"]]
        k = row["Kernel Name"].split("(")[0]
        agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("%-64s %8s %14s %10s %8s" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
        print("%-64s %8d %14.1f %10.1f %8.4f" % (k[:64], v[0], v[1], v[1] / v[0], v[1] / tot))


def report(path, longest_only=False):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none : %s" % path)
    body = rows[2:]
    if longest_only:                       # the launch with every instance active (first tick of a step)
        i_t = hdr.index("gpu__time_duration.sum")
        body = [max(body, key=lambda r: float(r[i_t].replace(",", "")))]
        print("# longest of %d captured launches" % (len(rows) - 2))
    for r in body:
        name = r[hdr.index("Kernel Name")]
        print("\n== %s" % name)
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print("%-82s %16s %s" % (key, r[i], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "longest":
        report(sys.argv[2], longest_only=True)
    else:
        {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
