"""Device time of the fused epilogue (change of variables + post-processor + energy) per instance size."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvm_b200 import engine as E  # noqa: E402
from tools.quick_bench import synth  # noqa: E402

dev = torch.device("cuda:0")
b = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for n in (20, 70, 128, 160, 200, 250):
    q, v, f = synth(n, 0, 0.05, dev)
    state = (torch.rand(b, n, device=dev) - 0.5)
    for pp in (None, "grad-descent", "adam"):
        for _ in range(3):
            E.epilogue(state, q, v, map1=(1.0, 0.5), post_processor=pp, pp_iterations=10, scaled_by=f)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            E.epilogue(state, q, v, map1=(1.0, 0.5), post_processor=pp, pp_iterations=10, scaled_by=f)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        flops = b * (2.0 * n * n) * ((10 if pp == "grad-descent" else 1 if pp == "adam" else 0) + 1)
        print(json.dumps({"n": n, "batch": b, "post_processor": pp, "us": round(us, 1), "gflops": round(flops / us / 1e3, 1)}))
