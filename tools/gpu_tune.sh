#!/bin/bash
# kernel tuning pass: every library under build/alt (and the main one) through tools/quick_bench.py
tag=${1:-tune}
out=gpurun_out/$tag
mkdir -p $out
sizes=${SIZES:-70}
for n in $sizes; do
  timeout 300 python tools/quick_bench.py --n $n --reps 9 > $out/quick_n${n}_main.jsonl 2>&1
  for alt in build/alt/*.so; do
    [ -f "$alt" ] || continue
    name=$(basename $alt .so)
    CCVM_B200_LIB=$PWD/$alt timeout 300 python tools/quick_bench.py --n $n --reps 9 > $out/quick_n${n}_$name.jsonl 2>&1
  done
done
python - <<PY
import json,glob,os
rows={}
for f in sorted(glob.glob("$out/quick_n*_*.jsonl")):
    name=os.path.basename(f)[:-6]
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        if "solver" in d: rows.setdefault(d["solver"],{})[name]=d["frac_of_ffma2_peak"]
names=sorted({k for v in rows.values() for k in v})
print("solver".ljust(22)+" ".join(n.replace("quick_","").replace("libccvm_","")[:14].rjust(15) for n in names))
for s,v in rows.items():
    print(s.ljust(22)+" ".join((f"{v.get(n,0):.4f}").rjust(15) for n in names))
PY
for sv in ${NCU_SOLVERS:-}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sde_tmem -s 4 -c 1 -f -o $out/ncu_${sv}_n${NCU_N:-70} \
    python tools/quick_bench.py --n ${NCU_N:-70} --only $sv --reps 1 > $out/ncu_$sv.log 2>&1; echo "ncu $sv rc=$?"
done
