#!/bin/bash
# issuer / update-warp variants of the small-n tensor-core kernel (build/alt/lib*.so): parity, then timings
tag=${1:-m17}
out=gpurun_out/$tag
mkdir -p $out
for v in ${VARIANTS:-main alt1 wpq2 wpq2alt}; do
  if [ $v = main ]; then unset CCVM_B200_LIB; else export CCVM_B200_LIB=$PWD/build/alt/lib$v.so; fi
  timeout 300 python tools/mma_check.py --n 70 > $out/check_$v.jsonl 2>$out/check_$v.err; echo "$v check rc=$?" | tee -a $out/rc.txt
  grep -c '"ok": true' $out/check_$v.jsonl; grep '"ok": false' $out/check_$v.jsonl | head -3; tail -2 $out/check_$v.err
  for n in ${SIZES:-70 128 40}; do
    CCVM_MMA=1 timeout 300 python tools/quick_bench.py --n $n --reps 5 > $out/quick_n${n}_$v.jsonl 2>>$out/quick_$v.err; echo "$v quick n=$n rc=$?" | tee -a $out/rc.txt
  done
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
for k in sorted(rows): print(k[0], k[1].ljust(22), "  ".join(f"{v} {ms:.4f}" for v, ms in rows[k].items()))
PY
