#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py.

usage: summarize_launches.py launches.csv [title] > summary.md
Splits the list into pipeline runs at every k_bucket_heads launch and tabulates the LAST run
(a device-resident step: bench.py times those after the end-to-end steps).  Times under ncu are
cold-cache and serialised: compare shares, not absolutes."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr, L = None, []
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
            L.append((d["Kernel Name"].split("(")[0].replace("void ", ""), v))
    starts = [i for i, (k, _) in enumerate(L) if k == "k_bucket_heads"]
    runs = [(a, b) for a, b in zip(starts, starts[1:] + [len(L)])]
    print(f"# {title}\n")
    print(f"{len(L)} launches captured, {len(runs)} pipeline runs: " +
          ", ".join(f"{b - a} launches / {sum(v for _, v in L[a:b]) / 1e3:.2f} ms" for a, b in runs) + "\n")
    a, b = runs[-1]
    agg = collections.OrderedDict()
    for k, v in L[a:b]:
        k = "cub::" + k.split("::")[1].split("<")[0] if k.startswith("cub::") else k
        e = agg.setdefault(k, [0, 0.0])
        e[0] += 1
        e[1] += v
    tot = sum(e[1] for e in agg.values())
    print(f"Last run ({b - a} launches, {tot / 1e3:.3f} ms of kernel time), in first-launch order:\n")
    print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for k, (n, v) in agg.items():
        print(f"| `{k}` | {n} | {v:.1f} | {100 * v / tot:.1f}% |")
    print("\nTop by time: " + ", ".join(f"`{k}` {100 * v / tot:.1f}%" for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:8]))


if __name__ == "__main__":
    main()
