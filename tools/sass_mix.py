#!/usr/bin/env python
"""Development aid: static opcode histogram of the longest loop of one function in a `cuobjdump -sass`
dump (the iteration loop of a fully unrolled compile-time-CG kernel is straight-line code, so the static
count IS the per-warp-iteration count).  usage: sass_mix.py all.sass <function substring> [--full]"""
import collections
import re
import sys

path, key = sys.argv[1], sys.argv[2]
ops, on = [], False
for l in open(path):
    if "Function :" in l:
        on = key in l
        continue
    if not on:
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ops.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ops)}
loops = []
for i, (a, o) in enumerate(ops):
    if "BRA" in o:
        t = re.search(r"0x([0-9a-f]+)", o)
        if t and int(t.group(1), 16) < a:
            loops.append((addr.get(int(t.group(1), 16)), i))
# the iteration loop is the longest loop that holds packed FMAs
def n_ffma2(lp):
    return sum(1 for _, o in ops[lp[0]:lp[1] + 1] if "FFMA2" in o)


a, b = max(loops, key=lambda x: (n_ffma2(x) > 0, x[1] - x[0]))


def mn(s):
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    t = s.split()[0].split(".")
    return t[0] + ("." + t[1] if t[0] in ("IMAD", "MUFU", "LDS", "STS", "BAR", "FMNMX") and len(t) > 1 else "")


c = collections.Counter(mn(o) for _, o in ops[a:b + 1])
tot = sum(c.values())
print(f"{len(ops)} instructions; longest loop [{a}, {b}] = {tot}")
for k, v in c.most_common():
    print(f"  {k:14s} {v:5d} {100 * v / tot:6.2f}%")
if "--full" in sys.argv:
    for _, o in ops[a:b + 1]:
        print(o)
