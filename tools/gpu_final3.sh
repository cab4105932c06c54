#!/bin/bash
# per-tile arrival point of the one-tile tensor-core kernels: main = DL late (CCVM_MMA_LATE_MASK 0x01), build/alt = + MF-adam, Langevin-adam
tag=${1:-final3}
out=gpurun_out/$tag
mkdir -p $out
timeout 120 python -m pytest tests/test_gpu_mma.py -m gpu -q -x -p no:cacheprovider > $out/pytest_mma.log 2>&1; echo "pytest mma rc=$?" | tee -a $out/rc.txt
tail -1 $out/pytest_mma.log
for n in 70 100 128; do
  timeout 60 python tools/quick_bench.py --n $n --reps 7 --only dl,dl_adam > $out/quick_n${n}_main.jsonl 2>>$out/quick.err
done
for alt in build/alt/*.so; do
  [ -f "$alt" ] || continue
  name=$(basename $alt .so)
  for n in 70 100; do
    CCVM_B200_LIB=$PWD/$alt timeout 60 python tools/quick_bench.py --n $n --reps 7 --only mf_adam,langevin_adam > $out/quick_n${n}_$name.jsonl 2>>$out/quick.err
  done
done
python - <<PY
import json,glob,os
for f in sorted(glob.glob("$out/quick_n*_*.jsonl")):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        if "solver" in d: print(os.path.basename(f)[6:-6].ljust(28), d["solver"].ljust(16), d["ms"], d["frac_of_ffma2_peak"])
PY
timeout 100 python bench.py > $out/bench_n1.json 2>$out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
python -c "
import json; d=json.load(open('$out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['sweep']['wall_s'], d['oracle_check']['ok'])"
timeout 150 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
grep -E "^FAILED|^ERROR|passed|failed" $out/pytest_gpu.log | tail -5
