#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total ms, share, average.
usage: ncu_launches.py launches.csv [--md]"""
import collections
import csv
import re
import sys

path = sys.argv[1]; md = "--md" in sys.argv
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]; unit = r["Metric Unit"]; v = float(r["Metric Value"].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*$", "", name).replace("void ", "")
    agg[name][0] += 1; agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"Total device time in the list: {tot:.1f} ms\n")
if md:
    print("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if md:
        print(f"| `{k}` | {n} | {ms:.2f} | {100 * ms / tot:.1f}% | {ms / n:.3f} |")
    else:
        print(f"{ms:10.2f} ms {100 * ms / tot:5.1f}%  n={n:5d} avg={ms / n:8.3f}  {k}")
