#!/usr/bin/env python
"""Per-kernel roofline table from an `ncu --metrics ... --csv` log (tools/gpu_round.sh kmetrics): per kernel name the
launches, total time, DRAM bytes, achieved DRAM GB/s (bytes / duration) against the measured copy peak, ncu's DRAM and
tensor-pipe utilisation."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (row["ID"], re.sub(r"\(.*", "", row["Kernel Name"]))
        v, unit = to_float(row["Metric Value"]), row["Metric Unit"]
        name = row["Metric Name"]
        if name == "gpu__time_duration.sum":
            v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)   # -> us
        if name.startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            v *= mult
        per.setdefault(key, {})[name] = v
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    for (_, kname), m in per.items():
        a = agg[kname]
        a["n"] += 1
        a["us"] += m.get("gpu__time_duration.sum", 0.0)
        a["bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        t = m.get("gpu__time_duration.sum", 0.0)
        a["dram_pct_t"] += m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0) * t
        a["tensor_pct_t"] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t
        a["regs"] = m.get("launch__registers_per_thread", 0.0)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except OSError:
        peak = 6650.0
    tot = sum(a["us"] for a in agg.values())
    print(f"# {path}: {int(sum(a['n'] for a in agg.values()))} launches, {tot / 1e3:.2f} ms; DRAM peak {peak} GB/s (measured copy)")
    print(f"{'ms':>9} {'share':>6} {'n':>4} {'GB moved':>9} {'GB/s':>7} {'of peak':>7} {'ncu dram%':>9} {'tensor%':>7} {'regs':>4}  kernel")
    for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"])[:top]:
        gbs = a["bytes"] / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
        print(f"{a['us'] / 1e3:9.3f} {100 * a['us'] / tot:5.1f}% {int(a['n']):4d} {a['bytes'] / 1e9:9.2f} {gbs:7.0f} {gbs / peak:7.2f} "
              f"{a['dram_pct_t'] / a['us'] if a['us'] else 0:9.1f} {a['tensor_pct_t'] / a['us'] if a['us'] else 0:7.1f} {int(a['regs']):4d}  {k[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
