#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table per pipeline run.
usage: summarize_launches.py launches.csv "title" [first-kernel-of-a-run, default k_bucket_heads] > summary.md"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    first = sys.argv[3] if len(sys.argv) > 3 else "k_bucket_heads"
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r["Kernel Name"]))
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"^(cub::)(CUB_[0-9_A-Za-z]+::)?(detail::)?([a-z_]+::)*", r"\1", name)
        rows.append((name, float(r["Metric Value"].replace(",", "")) / 1e3))
    runs, cur = [], []
    for name, us in rows:
        if name == first and cur:
            runs.append(cur); cur = []
        cur.append((name, us))
    if cur:
        runs.append(cur)
    print(f"# {title}\n")
    print(f"{len(rows)} launches captured, {len(runs)} runs starting at `{first}`: " + ", ".join(f"{len(r)} launches / {sum(u for _, u in r) / 1e3:.2f} ms" for r in runs) + "\n")
    last = runs[-1]
    tot = sum(u for _, u in last)
    agg = OrderedDict()
    for name, us in last:
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
    print(f"Last run ({len(last)} launches, {tot / 1e3:.3f} ms of kernel time), in first-launch order:\n")
    print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for name, (c, us) in agg.items():
        print(f"| `{name}` | {c} | {us:.1f} | {100 * us / tot:.1f}% |")


if __name__ == "__main__":
    main()
