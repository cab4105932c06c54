#!/bin/bash
# two-M-tile tensor-core kernel: items-per-lane cap 8 against 11 for the tiles with a large per-item state (one or two waves of CTAs)
tag=${1:-m27}
out=gpurun_out/$tag
mkdir -p $out
CCVM_MMA_IPL_MAX=11 timeout 600 python tools/mma_check.py --n 160 192 --only dl_adam,mf,mf_adam > $out/check.jsonl 2>$out/check.err; echo "check rc=$?" | tee -a $out/rc.txt
grep -c '"ok": true' $out/check.jsonl; grep '"ok": false' $out/check.jsonl | cut -c1-300 | head -8; tail -3 $out/check.err
for n in ${SIZES:-130 144 160 176 192}; do
  for c in 8 9 10 11; do
    CCVM_MMA_IPL_MAX=$c timeout 300 python tools/quick_bench.py --n $n --reps 3 --only dl_adam,mf,mf_adam > $out/quick_n${n}_c$c.jsonl 2>>$out/quick.err
  done
done
python - <<PY
import json, glob, os
rows = {}
for f in sorted(glob.glob("$out/quick_n*_*.jsonl")):
    b = os.path.basename(f)[:-6].split("_"); n, v = b[1], b[2]
    for l in open(f):
        try: d = json.loads(l)
        except Exception: continue
        if "solver" in d: rows.setdefault((n, d["solver"]), {})[v] = d["ms"]
for k in sorted(rows): print(k[0].ljust(6), k[1].ljust(22), "  ".join(f"{v} {ms:.4f}" for v, ms in sorted(rows[k].items(), key=lambda x: int(x[0][1:]))))
PY
tail -3 $out/quick.err
