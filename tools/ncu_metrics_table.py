#!/usr/bin/env python
"""Per-kernel table from an `ncu --metrics ... --csv` log (one row per launch and metric): averages per kernel name.
usage: ncu_metrics_table.py file.csv [--md]"""
import collections, csv, re, sys
path = sys.argv[1]; md = "--md" in sys.argv
lines = [ln for ln in open(path, newline="") if not ln.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.defaultdict(lambda: collections.defaultdict(list)); order = []
for r in rd:
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "")
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else float("nan")
    u = r["Metric Unit"]; m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    if m.startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    agg[name][m].append(v)
    if name not in order: order.append(name)
cols = [("gpu__time_duration.sum", "avg ms"), ("sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "DMMA inst"), ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "DMMA pipe %"),
        ("sm__inst_executed_pipe_fp64.sum", "FP64 inst"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"), ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
hdr = ["kernel", "launches"] + [c[1] for c in cols]
print(("| " + " | ".join(hdr) + " |\n|" + "---|" * len(hdr)) if md else " ; ".join(hdr))
for name in sorted(order, key=lambda n: -sum(agg[n]["gpu__time_duration.sum"])):
    a = agg[name]; n = len(a["gpu__time_duration.sum"])
    vals = []
    for m, _ in cols:
        x = a.get(m, [])
        vals.append("-" if not x else (f"{sum(x) / len(x):.3f}" if "inst" not in m else f"{sum(x) / len(x):.3e}"))
    row = [f"`{name[:60]}`", str(n)] + vals
    print(("| " + " | ".join(row) + " |") if md else " ; ".join(row))
