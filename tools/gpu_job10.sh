#!/bin/bash
# small-n tensor-core kernel: issuer pacing / warpgroup stagger on and off over sizes, parity
tag=${1:-m20}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python tools/mma_check.py --n 70 100 > $out/check.jsonl 2>$out/check.err; echo "check rc=$?" | tee -a $out/rc.txt
grep -c '"ok": true' $out/check.jsonl; grep '"ok": false' $out/check.jsonl | head -3; tail -2 $out/check.err
for mode in ${MODES:-11 00 10}; do
  s=${mode:0:1}; p=${mode:1:1}
  for n in ${SIZES:-40 60 70 80 100 128}; do
    CCVM_MMA=1 CCVM_MMA_STAGGER=$s CCVM_MMA_PACE=$p timeout 300 python tools/quick_bench.py --n $n --reps 7 > $out/quick_n${n}_m$mode.jsonl 2>>$out/quick.err; echo "mode $mode quick n=$n rc=$?" >> $out/rc.txt
  done
  CCVM_MMA=1 CCVM_MMA_STAGGER=$s CCVM_MMA_PACE=$p timeout 300 python tools/quick_bench.py --n 70 --batch 8192 --reps 5 > $out/quick_n70b8192_m$mode.jsonl 2>>$out/quick.err
  CCVM_MMA=1 CCVM_MMA_STAGGER=$s CCVM_MMA_PACE=$p timeout 300 python tools/quick_bench.py --n 70 --iters 15000 --reps 3 > $out/quick_n70t15000_m$mode.jsonl 2>>$out/quick.err
done
python - <<PY
import json, glob, os
rows = {}
for f in sorted(glob.glob("$out/quick_n*_m*.jsonl")):
    b = os.path.basename(f)[:-6].split("_"); n, v = b[1], b[2]
    for l in open(f):
        try: d = json.loads(l)
        except Exception: continue
        if "solver" in d: rows.setdefault((n, d["solver"]), {})[v] = d["ms"]
for k in sorted(rows): print(k[0].ljust(10), k[1].ljust(22), "  ".join(f"{v} {ms:.4f}" for v, ms in sorted(rows[k].items())))
PY
