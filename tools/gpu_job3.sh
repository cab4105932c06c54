#!/bin/bash
tag=${1:-job}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
grep -E "^FAILED|passed|failed" $out/pytest_gpu.log | tail
for route in many single; do
  timeout 1500 python tools/equivalence_gpu.py --seeds 0,1,2,3 --route $route --out $out/equivalence_$route.json > $out/equivalence_$route.log 2>&1
  grep -E "majority|EQUIV" $out/equivalence_$route.log
done
