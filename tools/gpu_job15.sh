#!/bin/bash
# two-M-tile tensor-core kernel against the hybrid kernel over batch sizes (selection rule for 128 < n <= 192)
tag=${1:-m28}
out=gpurun_out/$tag
mkdir -p $out
for n in ${SIZES:-144 192}; do
  for b in ${BATCHES:-128 512 1024 2048}; do
    CCVM_MMA=1 timeout 300 python tools/quick_bench.py --n $n --batch $b --reps 3 --only dl,dl_adam,mf_adam,langevin > $out/q_n${n}_b${b}_mma.jsonl 2>>$out/quick.err
    CCVM_MMA=0 timeout 300 python tools/quick_bench.py --n $n --batch $b --reps 3 --only dl,dl_adam,mf_adam,langevin > $out/q_n${n}_b${b}_hyb.jsonl 2>>$out/quick.err
  done
done
python - <<PY
import json, glob, os
rows = {}
for f in sorted(glob.glob("$out/q_n*_*.jsonl")):
    b = os.path.basename(f)[:-6].split("_"); n, bb, v = b[1], b[2], b[3]
    for l in open(f):
        try: d = json.loads(l)
        except Exception: continue
        if "solver" in d: rows.setdefault((n, int(bb[1:]), d["solver"]), {})[v] = d["ms"]
for k in sorted(rows): print(k[0].ljust(5), str(k[1]).ljust(5), k[2].ljust(12), "  ".join(f"{v} {ms:.4f}" for v, ms in sorted(rows[k].items())))
PY
tail -3 $out/quick.err
