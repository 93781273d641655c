#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the captured window).

  python tools/ncu_summary.py gpurun_out/r1_launches.csv > profiles/r1_launches_summary.md
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
        n += 1
    print(f"# ncu launch list `{path}`: {n} launches, {tot / 1e3:.2f} ms of device time (cold-cache, serialised)\n")
    print("| share | time (us) | launches | kernel |\n|---:|---:|---:|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * t / tot:.1f}% | {t:.1f} | {c} | `{k[:120]}` |")


if __name__ == "__main__":
    main(sys.argv[1])
