#!/usr/bin/env python
"""Opcode histogram per kernel of the built library (cuobjdump -sass): the evidence of which instruction families carry the hot path
(DMMA = FP64 tensor-core mma.sync m8n8k4, DFMA/DADD/DMUL = FP64 FMA pipe, LDGSTS = cp.async, REDUX = redux.sync, LDG/STG/LDS/STS, BAR, SHFL).
usage: sass_hist.py [lib.so] [--md]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = next((a for a in sys.argv[1:] if not a.startswith("--")), os.path.join(ROOT, "alf_b200", "libalf_b200.so")); md = "--md" in sys.argv
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fam = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "LDG", "STG", "LDGSTS", "LDS", "STS", "ATOMS", "RED", "ATOMG", "SHFL", "REDUX", "CREDUX", "BAR", "UTMALDG", "UBLKCP", "HMMA", "UTCHMMA"]
kern = collections.OrderedDict(); cur = None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1); kern[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        kern[cur][m.group(1)] += 1; kern[cur]["__total"] += 1
dem = subprocess.run(["c++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
rows = []
for (k, c), name in zip(kern.items(), dem):
    name = re.sub(r"\(.*$", "", name).replace("void ", "")
    rows.append((name, c))
agg = collections.OrderedDict()
for name, c in rows:      # merge instantiations that differ only in template arguments of the same kernel family
    base = re.sub(r"<.*$", "", name)
    agg.setdefault(base, [0, collections.Counter()]); agg[base][0] += 1; agg[base][1].update(c)
hdr = ["kernel (instantiations)", "SASS instr"] + fam
print(("| " + " | ".join(hdr) + " |\n|" + "---|" * len(hdr)) if md else " ; ".join(hdr))
tot = collections.Counter()
for base, (n, c) in sorted(agg.items(), key=lambda kv: -kv[1][1]["__total"]):
    tot.update(c)
    vals = [f"`{base}` ({n})", str(c["__total"])] + [str(c[f]) if c[f] else "" for f in fam]
    print(("| " + " | ".join(vals) + " |") if md else " ; ".join(vals))
print(("| **whole library** | " if md else "TOTAL ; ") + str(tot["__total"]) + (" | " if md else " ; ") + (" | " if md else " ; ").join(str(tot[f]) if tot[f] else "" for f in fam) + (" |" if md else ""))
