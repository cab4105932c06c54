#!/bin/bash
# development pass for the small-n tensor-core kernel: parity against the oracle, then timings with and without it
tag=${1:-mma}
out=gpurun_out/$tag
mkdir -p $out
timeout 200 python tools/mma_check.py --only langevin --batch 60 --iters 20 > $out/check_first.jsonl 2>$out/check_first.err; echo "first rc=$?" | tee -a $out/rc.txt
cat $out/check_first.jsonl; tail -5 $out/check_first.err
timeout 600 python tools/mma_check.py --n ${CHECK_N:-70} > $out/check.jsonl 2>$out/check.err; echo "check rc=$?" | tee -a $out/rc.txt
cat $out/check.jsonl; tail -5 $out/check.err
for n in ${SIZES:-70}; do
  CCVM_MMA=1 timeout 300 python tools/quick_bench.py --n $n --reps 5 > $out/quick_n${n}_mma.jsonl 2>$out/quick_mma.err; echo "quick mma rc=$?" | tee -a $out/rc.txt
  CCVM_MMA=0 timeout 300 python tools/quick_bench.py --n $n --reps 5 > $out/quick_n${n}_tiled.jsonl 2>&1
  python - <<PY
import json
for f in ("$out/quick_n${n}_mma.jsonl", "$out/quick_n${n}_tiled.jsonl"):
    print(f)
    for l in open(f):
        try: d = json.loads(l)
        except Exception: continue
        if "solver" in d: print("  %-22s %8.4f ms  frac %.4f" % (d["solver"], d.get("ms", d.get("ms_median", 0)), d["frac_of_ffma2_peak"]))
PY
done
tail -3 $out/quick_mma.err
