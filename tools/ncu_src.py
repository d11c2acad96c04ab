#!/usr/bin/env python3
"""Summarise `ncu --page source --csv --print-source sass`: hottest instructions by stall samples, and stall-reason
totals. usage: ncu_src.py src.csv [kernel-index] [top-n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
# split per kernel
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
b = blocks[k]; hdr = b["rows"][0]; data = [r for r in b["rows"][1:] if len(r) == len(hdr)]
print(b["name"], "instructions:", len(data))
ci = {h: i for i, h in enumerate(hdr)}
samp = ci["# Samples"]; ex = ci["Instructions Executed"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
tot = sum(int(r[samp]) for r in data)
print("total samples", tot)
sums = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
for h, v in sorted(sums.items(), key=lambda kv: -kv[1])[:14]:
    print("  %-28s %7d %.3f" % (h, v, v / max(tot, 1)))
print("hot instructions:")
order = sorted(range(len(data)), key=lambda i: -int(data[i][samp]))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:3]
    print("%5d %-58s smp %5s exec %9s  %s" % (i, r[1].strip()[:58], r[samp], r[ex], " ".join("%s:%d" % (n, v) for v, n in st if v)))
