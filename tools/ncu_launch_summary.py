#!/usr/bin/env python3
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list: one short line per launch (index, kernel, grid,
block, microseconds) followed by a per-kernel table (launches, total, share).  usage: ncu_launch_summary.py in.csv > out.csv"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|mac::|void |at::native::|at::", "", name)
    return re.sub(r"\(.*", "", name)[:90]


print("# launch,kernel,grid,block,us")
agg = collections.OrderedDict()
total = 0.0
for i, r in enumerate(rows):
    us = float(r["Metric Value"].replace(",", "")) / 1e3
    k = short(r["Kernel Name"])
    print("%d,%s,%s,%s,%.2f" % (i, k, r["Grid Size"].replace(",", " "), r["Block Size"].replace(",", " "), us))
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
print("# per kernel: launches,total_us,share")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("# %-90s %5d %12.1f %6.2f%%" % (k, n, t, 100 * t / total))
