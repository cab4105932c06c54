#!/bin/bash
# last validation pass of the round at HEAD (short GPU budget): GPU tests, smoke, bench
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
timeout 240 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
grep -E "^FAILED|^ERROR|passed|failed" $out/pytest_gpu.log | tail -20
timeout 60 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/rc.txt
timeout 150 python bench.py > $out/bench_n1.json 2>$out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
head -c 600 $out/bench_n1.json
