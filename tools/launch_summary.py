#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of total device time)."""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)      # -> us
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {tot / 1e3:.2f} ms total device time")
    print(f"{'ms':>10} {'share':>6} {'launches':>8}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{t / 1e3:10.3f} {100 * t / tot:5.1f}% {n:8d}  {k[:120]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
