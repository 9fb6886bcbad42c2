#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line:
top lines by warp-stall samples with their dominant stall reasons.  usage: ncu_lines.py file.csv [topN] [kernel-index]"""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path, newline='')))
cur_file = None; hdr = None; kernel = -1
agg = collections.defaultdict(lambda: collections.Counter()); src = {}
want_kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": kernel += 1; continue
    if r[0] == "File Name": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try: ln = int(r[0])
    except ValueError: continue
    key = (cur_file, ln); src[key] = r[1]
    for i, h in enumerate(hdr):
        if i < 4: continue
        if h == "# Samples" or h == "Instructions Executed" or (h.startswith("stall_") and "Not Issued" not in h) or h in ("L1 Wavefronts Shared Excessive", "L2 Theoretical Sectors Global"):
            try: agg[key][h] += float(r[i].replace(',', '') or 0)
            except ValueError: pass
tot = sum(v["# Samples"] for v in agg.values()) or 1
print(f"total samples {tot:.0f}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:topn]:
    st = sorted(((k, x) for k, x in v.items() if k.startswith("stall_")), key=lambda t: -t[1])[:3]
    print(f"{100*v['# Samples']/tot:5.1f}% {key[0]}:{key[1]:<4d} inst={v['Instructions Executed']:.0f} " + " ".join(f"{k[6:]}={100*x/max(v['# Samples'],1):.0f}%" for k, x in st) + f"  | {src[key].strip()[:110]}")
