#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): headline metrics, stall reasons and the hottest
source lines.  usage: ncu_summary.py report.ncu-rep [n_lines]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
        "local_load", "local_store", "smsp__inst_executed_op_local"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("== kernel:", d.get("Kernel Name"))
        for k in hdr:
            if any(s in k for s in KEYS) and d[k] not in ("", "n/a"):
                if "TriageCompute" in k:
                    continue
                print("  %-90s %s %s" % (k, d[k], u[k]))
        st = [(float(d[k]), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and
              k.endswith("per_issue_active.ratio") and "not_issued" not in k and d[k] not in ("", "n/a")]
        print("  -- stall cycles per issued instruction")
        for v, k in sorted(st, reverse=True)[:9]:
            print("     %-30s %.3f" % (k.split("stalled_")[1].split("_per_issue")[0], v))
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur, hd, line = None, None, None
    agg = defaultdict(lambda: [0, 0, ""])
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hd = r
            i_s, i_i = hd.index("# Samples"), hd.index("Instructions Executed")
            continue
        if hd is None:
            continue
        if r[0] != "":
            line = (cur, int(r[0]))
            agg[line][2] = r[1]
        try:
            s, ins = int(r[i_s] or 0), int(r[i_i] or 0)
        except (ValueError, IndexError):
            continue
        if len(r) > 2 and r[2] != "":
            agg[line][0] += s
            agg[line][1] += ins
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print("== hottest source lines (samples %d, warp instructions %d)" % (ts, ti))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:nlines]:
        print("%-12s %4d  samp %5.1f%%  inst %5.1f%%  %s" % (k[0], k[1], 100 * v[0] / ts, 100 * v[1] / ti, v[2].strip()[:100]))


if __name__ == "__main__":
    main()
