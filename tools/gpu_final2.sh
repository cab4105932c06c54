#!/bin/bash
# the one-tile kernels with their own layout (main) against the two-tile layout / the late fence (build/alt), all eight loops
# at n = 70 and two more sizes; then GPU tests, smoke and the bench on the main library
tag=${1:-final2}
out=gpurun_out/$tag
mkdir -p $out
for n in 70 100 128; do
  timeout 120 python tools/quick_bench.py --n $n --reps 7 > $out/quick_n${n}_main.jsonl 2>>$out/quick.err
done
for alt in build/alt/*.so; do
  [ -f "$alt" ] || continue
  name=$(basename $alt .so)
  CCVM_B200_LIB=$PWD/$alt timeout 120 python tools/quick_bench.py --n 70 --reps 7 > $out/quick_n70_$name.jsonl 2>>$out/quick.err
done
python - <<PY
import json,glob,os
rows={}
for f in sorted(glob.glob("$out/quick_n*_*.jsonl")):
    name=os.path.basename(f)[:-6].replace("quick_","").replace("libccvm_","")
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        if "solver" in d: rows.setdefault(d["solver"],{})[name]=d["ms"]
names=sorted({k for v in rows.values() for k in v})
print("solver".ljust(22)+" ".join(n[:16].rjust(17) for n in names))
for s,v in rows.items(): print(s.ljust(22)+" ".join((f"{v.get(n,0):.4f}").rjust(17) for n in names))
PY
timeout 240 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
grep -E "^FAILED|^ERROR|passed|failed" $out/pytest_gpu.log | tail -20
timeout 60 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/rc.txt
timeout 150 python bench.py > $out/bench_n1.json 2>$out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
python -c "
import json; d=json.load(open('$out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['launch'], d['e2e']['value'], d['sweep']['wall_s'])"
