#!/bin/bash
# bench at 1 GPU (sweep sub-record after the sync-free planning change) + GPU tests + tcgen05 ncu summary
tag=${1:-job}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
grep -E "^FAILED|passed|failed" $out/pytest_gpu.log | tail
timeout 600 python bench.py > $out/bench_n1.json 2>$out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
python - <<PY
import json
d=json.load(open("$out/bench_n1.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"])
print("sweep", json.dumps(d.get("sweep"))[:500])
PY
timeout 600 python tools/tensor_peak.py > $out/tensor_peak.json 2>&1; cat $out/tensor_peak.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sde_tc2 -s 2 -c 1 -f -o $out/ncu_tc2_dl \
  python tools/tensor_peak.py > $out/ncu_tc2.log 2>&1; echo "ncu tc2 rc=$?" | tee -a $out/rc.txt
