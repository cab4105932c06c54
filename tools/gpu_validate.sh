#!/bin/bash
# validation pass at HEAD: smoke, GPU tests, bench (both arms), launch list + full capture of the bench kernel, size sweep
tag=${1:-job}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/rc.txt
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
grep -E "^FAILED|passed|failed" $out/pytest_gpu.log | tail
timeout 600 python bench.py > $out/bench_n1.json 2>$out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
timeout 600 python bench.py --impl reference > $out/bench_reference_arm.json 2>>$out/bench.err; echo "bench ref rc=$?" | tee -a $out/rc.txt
for n in 20 30 40 50 60 70 100 128 160 200 250; do
  timeout 300 python tools/quick_bench.py --n $n --reps 7 >> $out/quick_sizes.jsonl 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-extras > $out/ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a $out/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sde_mma -s 3 -c 1 -f -o $out/ncu_bench_kernel \
  python bench.py --steps 2 --warmup 3 --no-extras > $out/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $out/rc.txt
python - <<PY
import json
d=json.load(open("$out/bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"])
for k in ("sweep","config4","gpu_eager_baseline","tts","oracle_check"): print(k, json.dumps(d.get(k))[:600])
PY
grep -h "solver" $out/quick_sizes.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n'], d['solver'].ljust(22), d['ms'], round(d['traj_steps_per_s']/1e9,3), d['frac_of_ffma2_peak'])
"
