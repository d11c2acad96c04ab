#!/usr/bin/env python3
"""Print instruction-mix statistics of the innermost loops of one kernel in a cuobjdump -sass dump.
usage: sass_loop.py <sass.txt> <substring of function name>"""
import re, sys, collections
txt = open(sys.argv[1]).read().split("Function :")
fn = [t for t in txt if sys.argv[2] in t.split("\n")[0]][0]
ins = []
for line in fn.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2i = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, s) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?`\(\.L_x_\d+\)|BRA(?:\.\w+)*\s+.*0x([0-9a-f]+)", s)
    m2 = re.search(r"0x([0-9a-f]+)", s) if "BRA" in s else None
    if m2:
        tgt = int(m2.group(1), 16)
        if tgt <= a and tgt in addr2i:
            loops.append((addr2i[tgt], i))
print("total instructions", len(ins), "backward branches", len(loops))
for (s, e) in sorted(loops, key=lambda t: t[1] - t[0]):
    body = ins[s:e + 1]
    c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in body)
    print("loop %04x..%04x  n=%d  " % (ins[s][0], ins[e][0], len(body)), dict(c.most_common(14)))
