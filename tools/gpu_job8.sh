#!/bin/bash
# issuer modes of the small-n tensor-core kernel (CCVM_MMA_ALT = 0 ... 3; build/alt/lib*.so): iteration traces of both
# warpgroups for DL-adam, parity, timings
tag=${1:-m18}
out=gpurun_out/$tag
mkdir -p $out
for v in ${TRACES:-trace0 trace1 trace2 trace3}; do
  CCVM_B200_LIB=$PWD/build/alt/lib$v.so timeout 120 python tools/mma_trace.py dl_adam 70 > $out/$v.txt 2>$out/$v.err; echo "$v rc=$?" | tee -a $out/rc.txt
  sed -n 10,16p $out/$v.txt
done
for v in ${VARIANTS:-main alt1 alt2 alt3}; do
  if [ $v = main ]; then unset CCVM_B200_LIB; else export CCVM_B200_LIB=$PWD/build/alt/lib$v.so; fi
  if [ $v != main ] && [ $v != alt1 ]; then
    timeout 300 python tools/mma_check.py --n 70 > $out/check_$v.jsonl 2>$out/check_$v.err; echo "$v check rc=$?" | tee -a $out/rc.txt
    grep -c '"ok": true' $out/check_$v.jsonl; grep '"ok": false' $out/check_$v.jsonl | head -3; tail -2 $out/check_$v.err
  fi
  for n in ${SIZES:-70 128 40}; do
    CCVM_MMA=1 timeout 300 python tools/quick_bench.py --n $n --reps 7 > $out/quick_n${n}_$v.jsonl 2>>$out/quick_$v.err; echo "$v quick n=$n rc=$?" | tee -a $out/rc.txt
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
for k in sorted(rows): print(k[0], k[1].ljust(22), "  ".join(f"{v} {ms:.4f}" for v, ms in sorted(rows[k].items())))
PY
