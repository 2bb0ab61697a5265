#!/usr/bin/env python
"""Static SASS opcode mix of the hot kernels of libpinc_b200.so (cuobjdump -sass): instruction counts per opcode class,
proof of what the kernels are made of (fp64 arithmetic, 64/128-bit global accesses, shuffles, integer REDs; no tensor-core
or TMA instructions on this path: no step is a dense contraction and the grids are L2-resident).
    python tools/sass_static.py [kernel-name-substring ...] > profiles/rN_sass_mix.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "pinc_b200", "libpinc_b200.so")
want = sys.argv[1:] or ["k_acc", "k_distr_cells", "k_distr_tail", "k_scatter", "k_keys", "k_move", "k_findiff1st", "k_mg_solve", "k_mg_cluster"]
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
name, mix = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        mix[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        mix[name][".".join(op.split(".")[:3])] += 1
demangle = subprocess.run(["c++filt"] + list(mix), capture_output=True, text=True).stdout.splitlines()
for (mangled, c), pretty in zip(mix.items(), demangle):
    if not any(w in pretty for w in want):
        continue
    tot = sum(c.values())
    print(f"== {pretty[:150]}  [{tot} SASS instructions]")
    groups = collections.Counter()
    for op, n in c.items():
        b = op.split(".")[0]
        g = ("fp64" if b in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX") else "global ld/st" if b in ("LDG", "STG", "LD", "ST") else
             "atomics/RED" if b in ("RED", "ATOM", "ATOMG", "ATOMS") else "shared ld/st" if b in ("LDS", "STS", "LDSM") else
             "shuffle/vote" if b in ("SHFL", "VOTE", "MATCH", "REDUX") else "barrier/fence" if b in ("BAR", "MEMBAR", "ERRBAR", "CCTL", "WARPSYNC", "BSYNC", "BSSY") else
             "tensor/TMA" if b.startswith(("UTMA", "UTC", "HMMA", "IMMA", "DMMA", "QMMA", "UBLK", "SYNCS")) else "local (spill)" if b in ("LDL", "STL") else
             "branch" if b in ("BRA", "BRX", "EXIT", "RET", "CALL", "JMP") else "int/other")
        groups[g] += n
    print("   " + ", ".join(f"{g} {n} ({100 * n / tot:.0f}%)" for g, n in groups.most_common()))
    print("   top: " + ", ".join(f"{op} {n}" for op, n in c.most_common(14)))
