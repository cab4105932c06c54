#!/bin/bash
# small-n tensor-core kernel with two M tiles (128 < n <= 192): parity against the oracle, timings against the hybrid kernel
tag=${1:-m26}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python tools/mma_check.py --n ${CHECK_SIZES:-129 160 192 70 128} > $out/check.jsonl 2>$out/check.err; echo "check rc=$?" | tee -a $out/rc.txt
grep -c '"ok": true' $out/check.jsonl; grep '"ok": false' $out/check.jsonl | cut -c1-300 | head -8; tail -3 $out/check.err
for n in ${SIZES:-130 160 192}; do
  timeout 300 python tools/quick_bench.py --n $n --reps 5 > $out/quick_n${n}_mma.jsonl 2>>$out/quick.err; echo "mma n=$n rc=$?" >> $out/rc.txt
  CCVM_MMA=0 timeout 300 python tools/quick_bench.py --n $n --reps 3 > $out/quick_n${n}_hyb.jsonl 2>>$out/quick.err
done
python - <<PY
import json, glob, os
rows = {}
for f in sorted(glob.glob("$out/quick_n*_*.jsonl")):
    b = os.path.basename(f)[:-6].split("_"); n, v = b[1], b[2]
    for l in open(f):
        try: d = json.loads(l)
        except Exception: continue
        if "solver" in d: rows.setdefault((n, d["solver"]), {})[v] = d["ms"]
for k in sorted(rows): print(k[0].ljust(6), k[1].ljust(22), "  ".join(f"{v} {ms:.4f}" for v, ms in sorted(rows[k].items())))
PY
tail -3 $out/quick.err
