#!/bin/bash
# equivalence detail + noise-bias isolation
tag=${1:-job}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_equivalence.py > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
grep -E "^FAILED|passed|failed" $out/pytest_gpu.log | tail
for seed in 0 1; do
  timeout 900 python tools/equivalence_gpu.py --seed $seed --out $out/equivalence_seed$seed.json > $out/equivalence_seed$seed.log 2>&1
  tail -10 $out/equivalence_seed$seed.log
done
timeout 1500 python tools/noise_bias.py --out $out/noise_bias.json > $out/noise_bias.log 2>&1; echo "bias rc=$?" | tee -a $out/rc.txt
grep -v "^{" $out/noise_bias.log | tail -12
