#!/usr/bin/env python
"""Development aid: print the run-length-compressed mnemonic stream of the longest loop of one
function in a `cuobjdump -sass` dump.  usage: sass_loop.py all.sass <function substring>"""
import re, sys
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
a, b = max(loops, key=lambda x: x[1] - x[0])
def mn(s):
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    t = s.split()[0].split(".")
    return t[0] + ("." + t[1] if t[0] in ("IMAD", "MUFU", "LDS", "STS", "BRA") and len(t) > 1 else "")
seq = [mn(o) for _, o in ops[a:b + 1]]
out, prev, n = [], None, 0
for s in seq + [None]:
    if s == prev:
        n += 1
    else:
        if prev:
            out.append(prev + (f"x{n}" if n > 1 else ""))
        prev, n = s, 1
print(len(ops), "instructions; longest loop", a, b, "=", len(seq))
print(" ".join(out))
