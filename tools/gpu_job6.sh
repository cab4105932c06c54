#!/bin/bash
# sweep timing: chunk sizes, forked vs single-stream buckets
tag=${1:-job}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -2
for chunk in 16 32 64 128; do
  timeout 300 python tools/sweep_bench.py --sizes 20,30,40,50,60,70,80,90,100,110,120,130,140,150,160,170,180,190,200,210,220,230,240,250 --count 1024 --solvers langevin --chunk $chunk --warm 48 --on-device >> $out/sweep_fork.jsonl 2>&1
  CCVM_NO_FORK=1 timeout 300 python tools/sweep_bench.py --sizes 20,30,40,50,60,70,80,90,100,110,120,130,140,150,160,170,180,190,200,210,220,230,240,250 --count 1024 --solvers langevin --chunk $chunk --warm 48 --on-device >> $out/sweep_nofork.jsonl 2>&1
done
grep -h wall_s $out/sweep_fork.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('fork  ', d['chunk'], round(d['wall_s'],3), round(d['drift_tflops'],1), round(d['sum_kernel_solve_time_s'],3))"
grep -h wall_s $out/sweep_nofork.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('nofork', d['chunk'], round(d['wall_s'],3), round(d['drift_tflops'],1), round(d['sum_kernel_solve_time_s'],3))"
