"""Measured tensor-pipe ceiling of the tcgen05 path next to the achieved rate of the SDE kernel
(BASELINE configs[3]: n = 1024, batch 8192): dense TF32 TFLOP/s of back-to-back tcgen05.mma
(ccvm_microbench_tf32), the 3xTF32 ceiling (one third of it) and the logical drift TFLOP/s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvm_b200 import engine as E, _native as nat  # noqa: E402
from tools.quick_bench import synth  # noqa: E402

dev = torch.device("cuda:0")
out = {"tf32_dense_tflops_cta_group1": E.microbench_tf32(1), "tf32_dense_tflops_cta_group2": E.microbench_tf32(2),
       "fp32_ffma2_tflops": E.microbench_fp32(1)}
out["ceiling_3xtf32_tflops"] = out["tf32_dense_tflops_cta_group2"] / 3.0
n, b, t = 1024, 8192, 200
for name, sid, mult, kw in (
        ("dl", nat.SOLVER_DL, 0.2, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, feedback_scale=100.0, g=0.05)),
        ("langevin", nat.SOLVER_LANGEVIN, 0.05, dict(s=0.5, dt=0.002, sigma=0.5, feedback_scale=1.0))):
    q, v, _ = synth(n, 0, mult, dev)
    per = []
    for r in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        E.solve(sid, nat.ALG_ORIGINAL, q, v, b, t, seed=1, offset=r, **kw)
        e1.record()
        torch.cuda.synchronize()
        per.append(e0.elapsed_time(e1))
    ms = sorted(per[1:])[len(per[1:]) // 2]
    m = 2 if name == "dl" else 1
    tf = 2.0 * m * n * n * b * t / (ms * 1e-3) / 1e12
    # the MMAs the kernel issues also cover the padding rows of a partially filled last CTA: none at these sizes
    rows = m * b
    sms_used = min(148, (rows + 127) // 128)
    out[name] = {"ms": ms, "logical_drift_tflops": tf, "frac_of_3xtf32_ceiling": tf / out["ceiling_3xtf32_tflops"],
                 "ctas": (rows + 127) // 128,
                 "frac_of_ceiling_of_the_sms_it_occupies": tf / (out["ceiling_3xtf32_tflops"] * sms_used / 148)}
print(json.dumps(out))
