#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: launch_summary.py <launches.csv> [steps]   (steps: divide totals to get per-step figures)"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for k, (c, t) in agg.items() if k.startswith("mrtm::") and "fp64_peak" not in k)
    print("| kernel | launches | total us | avg us | share of mrtm kernels |")
    print("|---|---|---|---|---|")
    for k, (c, t) in agg.items():
        share = "%.1f%%" % (100 * t / tot) if k.startswith("mrtm::") and "fp64_peak" not in k else ""
        print("| %s | %d | %.1f | %.1f | %s |" % (k, c, t, t / c, share))
    if steps:
        print("\nper step (%g steps captured): %.1f us of mrtm kernels" % (steps, tot / steps))


if __name__ == "__main__":
    main()
