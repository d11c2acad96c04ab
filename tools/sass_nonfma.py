#!/usr/bin/env python3
"""List the non-FFMA instructions of an address range with the count of FFMAs issued before each.
usage: sass_nonfma.py sass.txt <function substring> <lo hex> <hi hex>"""
import re, sys
txt = open(sys.argv[1]).read().split("Function :")
fn = [t for t in txt if sys.argv[2] in t.split("\n")[0]][0]
lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
n = 0
for line in fn.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if not m: continue
    a, s = int(m.group(1), 16), m.group(2).strip()
    if lo <= a <= hi:
        if s.startswith("FFMA"): n += 1; continue
        print("%04x [%3d] %s" % (a, n, s))
