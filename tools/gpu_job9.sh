#!/bin/bash
# small-n tensor-core kernel: warpgroup stagger off / on (CCVM_MMA_STAGGER) over sizes, parity of both
tag=${1:-m19}
out=gpurun_out/$tag
mkdir -p $out
for s in 0 1; do
  CCVM_MMA_STAGGER=$s timeout 300 python tools/mma_check.py --n 70 > $out/check_s$s.jsonl 2>$out/check_s$s.err; echo "stagger $s check rc=$?" | tee -a $out/rc.txt
  grep -c '"ok": true' $out/check_s$s.jsonl; grep '"ok": false' $out/check_s$s.jsonl | head -3; tail -2 $out/check_s$s.err
  for n in ${SIZES:-40 50 60 70 80 100 128}; do
    CCVM_MMA=1 CCVM_MMA_STAGGER=$s timeout 300 python tools/quick_bench.py --n $n --reps 7 > $out/quick_n${n}_s$s.jsonl 2>>$out/quick_s$s.err; echo "stagger $s quick n=$n rc=$?" >> $out/rc.txt
  done
  CCVM_MMA=1 CCVM_MMA_STAGGER=$s timeout 300 python tools/quick_bench.py --n 70 --batch 8192 --reps 5 > $out/quick_n70b8192_s$s.jsonl 2>>$out/quick_s$s.err
done
python - <<PY
import json, glob, os
rows = {}
for f in sorted(glob.glob("$out/quick_n*_s*.jsonl")):
    b = os.path.basename(f)[:-6].split("_"); n, v = b[1], b[2]
    for l in open(f):
        try: d = json.loads(l)
        except Exception: continue
        if "solver" in d: rows.setdefault((n, d["solver"]), {})[v] = d["ms"]
for k in sorted(rows): print(k[0].ljust(10), k[1].ljust(22), "  ".join(f"{v} {ms:.4f}" for v, ms in sorted(rows[k].items())), " <- stagger" if rows[k].get("s1", 9) < 0.985 * rows[k].get("s0", 0) else "")
PY
