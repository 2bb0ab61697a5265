#!/usr/bin/env python
"""Opcode histogram (weighted by executed warp instructions and by stall samples) of one kernel of an .ncu-rep:
    ncu -i rep.ncu-rep --page source --csv --kernel-id :::N > k.csv ; python tools/sass_mix.py k.csv"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data)
c, s = Counter(), Counter()
for r in data:
    t = r[isrc].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = ".".join(op.split(".")[:2])
    c[op] += int(r[ia]); s[op] += int(r[isamp])
ts = sum(s.values()) or 1
print(rows[0][1][:80] if len(rows[0]) > 1 else "", "| warp instructions:", tot)
for op, n in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:18s} {n:10d} {100 * n / tot:5.1f}%   stall samples {100 * s[op] / ts:5.1f}%")
