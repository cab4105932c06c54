#!/bin/bash
# One gpurun call of a development round: smoke, GPU tests, per-solver kernel timings (main and alternative
# builds), the bench line.  Everything lands in gpurun_out/$1/.
tag=${1:-job}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/rc.txt
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
grep -E "^FAILED|passed|failed" $out/pytest_gpu.log | tail -40
timeout 300 python tools/quick_bench.py --n 70 > $out/quick_n70.jsonl 2>$out/quick_n70.err; echo "quick rc=$?" | tee -a $out/rc.txt
for alt in build/alt/*.so; do
  [ -f "$alt" ] || continue
  name=$(basename $alt .so)
  CCVM_B200_LIB=$PWD/$alt timeout 300 python tools/quick_bench.py --n 70 > $out/quick_n70_$name.jsonl 2>&1
done
timeout 300 python tools/quick_bench.py --n 20 > $out/quick_n20.jsonl 2>&1
timeout 300 python tools/quick_bench.py --n 128 > $out/quick_n128.jsonl 2>&1
timeout 600 python bench.py > $out/bench.json 2>$out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
cat $out/quick_n70.jsonl
for f in $out/quick_n70_*.jsonl; do echo "== $f"; cat $f; done
tail -c 3000 $out/bench.json
# ncu --set full of one launch of the kernels named in $NCU_SOLVERS (after the plain runs above exited)
for sv in ${NCU_SOLVERS:-}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sde_tmem -s 4 -c 1 -f -o $out/ncu_$sv \
    python tools/quick_bench.py --n 70 --only $sv --reps 1 > $out/ncu_$sv.log 2>&1; echo "ncu $sv rc=$?" | tee -a $out/rc.txt
done
