#!/usr/bin/env python
"""Print the mnemonic stream of one function from `cuobjdump -sass` output (development aid).
usage: sass_stream.py all.sass <function substring> [start_pattern] [count]"""
import re, sys
path, key = sys.argv[1], sys.argv[2]
pat = sys.argv[3] if len(sys.argv) > 3 else None
count = int(sys.argv[4]) if len(sys.argv) > 4 else 400
ops, on = [], False
for l in open(path):
    if "Function :" in l:
        on = key in l
        continue
    if not on:
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ops.append((m.group(1), m.group(2).strip()))
def mn(s):
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    return s.split()[0].split(".")[0] + ("." + s.split()[0].split(".")[1] if s.startswith(("IMAD.", "MUFU", "LDS", "STS", "BAR", "BRA")) and "." in s.split()[0] else "")
start = 0
if pat:
    for i, o in enumerate(ops):
        if re.search(pat, o[1]):
            start = max(0, i - 10)
            break
print(len(ops), "instructions; from", start)
print(" ".join(mn(o[1]) for o in ops[start:start + count]))
