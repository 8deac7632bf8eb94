#!/usr/bin/env python3
"""Key metrics of an `ncu --set full` report as compact JSON.

usage: ncu -i prof.ncu-rep --page raw --csv > raw.csv ; extract_ncu.py raw.csv > metrics.json"""
import csv
import json
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__cycles_active.avg", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
] + ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s for s in (
    "barrier", "long_scoreboard", "short_scoreboard", "wait", "not_selected", "math_pipe_throttle", "mio_throttle", "lg_throttle",
    "no_instruction", "branch_resolving")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out, seen = {}, {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0]
        seen[name] = seen.get(name, 0) + 1
        key = name if seen[name] == 1 else "%s#%d" % (name, seen[name])
        out[key] = {k: (r[ix[k]] + " " + units[ix[k]]).strip() for k in KEYS if k in ix}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
