"""Development aid: bench.py's `sweep` sub-record several times in one process (wall seconds and the host seconds
spent enqueueing): how stable is the software pipeline of a sweep?  usage: sweep_repeat.py [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    d = bench.sweep_record(0, 1, dev)
    print(json.dumps({"rep": r, "wall_s": round(d["wall_s"], 4), "host": d["host_seconds_rank0"], "drift_tflops": round(d["drift_tflops"], 2)}), flush=True)
