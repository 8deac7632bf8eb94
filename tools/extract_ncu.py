#!/usr/bin/env python3
"""Key metrics per kernel from an .ncu-rep (`ncu -i rep --page raw --csv`), as JSON: the last captured launch of each kernel.
usage: extract_ncu.py file.ncu-rep [more.ncu-rep ...] > metrics.json"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "local_load_requests", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def main():
    out = {}
    for rep in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        names, units = rows[hdr], rows[hdr + 1]
        kcol = names.index("Kernel Name")
        names = [n.split("TriageCompute.")[-1] for n in names]          # ncu prefixes some tensor metrics with 'TPC.TriageCompute.'
        for r in rows[hdr + 2:]:
            if len(r) != len(names):
                continue
            k = re.sub(r"\(.*", "", re.sub(r"<.*", "", r[kcol].replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))).replace("void ", "")
            out[k] = {m: (r[names.index(m)] + " " + units[names.index(m)]).strip() for m in KEEP if m in names}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
