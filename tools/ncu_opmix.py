#!/usr/bin/env python
"""Development aid: opcode mix (warp instructions per warp-iteration) and stall samples of the
first kernel in an `ncu --page source --csv --print-source sass` dump.
usage: ncu_opmix.py dump.csv <warps*iterations>"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
unit = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
# keep the first kernel only
end = len(rows)
for i, r in enumerate(rows[2:], 2):
    if r and r[0] == "Kernel Name":
        end = i
        break
print(rows[0][1][:90])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samp = collections.Counter(); tot = 0; ts = 0
for r in rows[2:end]:
    if len(r) < len(hdr): continue
    src = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
    t = src.split()[0].split(".")
    base = t[0] + ("." + t[1] if t[0] in ("IMAD", "MUFU", "LDS", "STS", "BAR") and len(t) > 1 else "")
    n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    ops[base] += n; samp[base] += s; tot += n; ts += s
print(f"total warp instructions {tot} = {tot/unit:.1f} per unit; samples {ts}")
for k, v in ops.most_common(40):
    print(f"{k:14s} {v/unit:9.1f} {100*v/tot:6.2f}%   samples {100*samp[k]/ts:5.1f}%")
