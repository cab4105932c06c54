#!/usr/bin/env python
"""Development aid (CPU only): registers / stack / static shared memory (`cuobjdump -res-usage`) and the counts of the
tcgen05 / TMA / packed-FP32 SASS mnemonics (`cuobjdump -sass`) of the named kernels in the object files of a build.
usage: static_resources.py <obj dir> > profiles/<tag>_static_resources.txt"""
import collections
import os
import re
import subprocess
import sys

OBJ = sys.argv[1] if len(sys.argv) > 1 else "/tmp/ccvm_b200_obj"
# (object, mangled-name substring, label)
KERNELS = [
    ("sde_mma_s0a1.o", "sde_mma_kernelILi0ELb1ELi4ELi1E", "sde_mma_kernel<DL, adam, 4 items, 1 M tile>  (bench kernel, n = 70)"),
    ("sde_mma_s0a0.o", "sde_mma_kernelILi0ELb0ELi4ELi1E", "sde_mma_kernel<DL, -, 4 items, 1 M tile>"),
    ("sde_mma_s2a0.o", "sde_mma_kernelILi2ELb0ELi4ELi1E", "sde_mma_kernel<Langevin, -, 4 items, 1 M tile>"),
    ("sde_mma_s1a1.o", "sde_mma_kernelILi1ELb1ELi4ELi1E", "sde_mma_kernel<MF, adam, 4 items, 1 M tile>"),
    ("sde_mma_s0a0.o", "sde_mma_kernelILi0ELb0ELi9ELi2E", "sde_mma_kernel<DL, -, 9 items, 2 M tiles>  (n = 160)"),
    ("sde_mma_s0a1.o", "sde_mma_kernelILi0ELb1ELi8ELi2E", "sde_mma_kernel<DL, adam, 8 items, 2 M tiles>"),
    ("sde_tmem_s0a1.o", "sde_tmem_kernelILi0ELb1ELi0ELb1ELi18E", "sde_tmem_kernel<DL, adam, TMEM, PIPE, CG 18>  (tiled kernel, n = 70)"),
    ("sde_tmem_s2a0.o", "sde_tmem_kernelILi2ELb0ELi0ELb1ELi18E", "sde_tmem_kernel<Langevin, -, TMEM, PIPE, CG 18>"),
    ("sde_tmem_s2a0.o", "sde_tmem_kernelILi2ELb0ELi0ELb1ELi5E", "sde_tmem_kernel<Langevin, -, TMEM, PIPE, CG 5>  (n = 20)"),
    ("sde_tc_s0.o", "sde_tc2_kernelILi0ELb0E", "sde_tc2_kernel<DL, ->  (n >= 256, tcgen05 3xTF32, CTA pairs)"),
    ("sde_tc_s2.o", "sde_tc2_kernelILi2ELb0E", "sde_tc2_kernel<Langevin, ->"),
]
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA",
         "MUFU", "LDS", "STS", "LDL", "STL", "LDG", "STG", "IMAD.WIDE", "ELECT")


def run(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


print("Static resources and SASS mnemonic counts (whole function: prologue + loop + fused tail), nvcc 12.9 -O3 sm_100a.")
print("LDL / STL = local-memory (stack) traffic; UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA,")
print("UTCBAR = tcgen05.commit, SYNCS = mbarrier operations.\n")
for obj, key, label in KERNELS:
    path = os.path.join(OBJ, obj)
    res = run("cuobjdump", "-res-usage", path).splitlines()
    name, usage = None, ""
    for i, l in enumerate(res):
        if "Function" in l and key in l:
            name = l.split("Function")[1].strip().rstrip(":")
            usage = res[i + 1].strip()
            break
    if name is None:
        print(f"{label}\n  (not in {obj})\n")
        continue
    sass = run("cuobjdump", "-sass", "-fun", name, path)
    c = collections.Counter()
    total = 0
    for l in sass.splitlines():
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", l)
        if not m:
            continue
        total += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w == "MUFU" and op.startswith("MUFU")):
                c[w] += 1
                break
    print(label)
    print("  " + usage)
    print(f"  {total} instructions; " + ", ".join(f"{w} {c[w]}" for w in WATCH if c[w]))
    print()
