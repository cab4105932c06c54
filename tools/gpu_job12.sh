#!/bin/bash
# A/B of two builds of the library on one box: build/alt/lib$ALT.so against the in-tree one (parity of both first)
tag=${1:-m22}
out=gpurun_out/$tag
mkdir -p $out
ALT=${ALT:-prev}
for v in main $ALT; do
  if [ $v = main ]; then unset CCVM_B200_LIB; else export CCVM_B200_LIB=$PWD/build/alt/lib$v.so; fi
  timeout 300 python tools/mma_check.py --n ${CHECK_SIZES:-40 70 100 128} > $out/check_$v.jsonl 2>$out/check_$v.err; echo "$v check rc=$?" | tee -a $out/rc.txt
  grep -c '"ok": true' $out/check_$v.jsonl; grep '"ok": false' $out/check_$v.jsonl | head -3; tail -2 $out/check_$v.err
done
for v in $ALT main $ALT main; do
  if [ $v = main ]; then unset CCVM_B200_LIB; else export CCVM_B200_LIB=$PWD/build/alt/lib$v.so; fi
  for n in ${SIZES:-40 70 100 128}; do
    CCVM_MMA=1 timeout 300 python tools/quick_bench.py --n $n --reps 7 >> $out/quick_n${n}_$v.jsonl 2>>$out/quick.err
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
        if "solver" in d: rows.setdefault((n, d["solver"]), {}).setdefault(v, []).append(d["ms"])
for k in sorted(rows): print(k[0].ljust(6), k[1].ljust(22), "  ".join(f"{v} " + "/".join(f"{m:.4f}" for m in ms) for v, ms in sorted(rows[k].items())))
PY
