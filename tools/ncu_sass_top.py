#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source sass`, with a few instructions of context,
and the opcode histogram of the kernel.  usage: ncu_sass_top.py file.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], newline=''))); topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None; ins = []
for r in rows:
    if len(r) > 2 and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].startswith("0x"): continue
    d = dict(zip(hdr, r))
    stalls = {k[6:]: float(v or 0) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0")}
    ins.append((r[0], d["Source"].strip(), float(d["# Samples"] or 0), float(d["Instructions Executed"] or 0), stalls))
tot = sum(x[2] for x in ins) or 1
print(f"{len(ins)} SASS instructions, {tot:.0f} samples")
order = sorted(range(len(ins)), key=lambda i: -ins[i][2])[:topn]
for i in order:
    a, s, n, ex, st = ins[i]
    top = " ".join(f"{k}={100 * v / max(n, 1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    prev = ins[i - 1][1][:40] if i else ""
    print(f"{100 * n / tot:5.1f}% #{i:5d} exec={ex:9.0f} {s[:70]:70s} | {top} | prev: {prev}")
hist = collections.Counter(x[1].split()[0 if not x[1].startswith('@') else 1].split('.')[0] for x in ins)
print("opcodes:", dict(hist.most_common(25)))
