"""Development aid: per-launch device time of one solver loop, 24 launches in a row
(start-up effects, run-to-run spread, bitwise reproducibility).
usage: anom_timing.py <n> <iterations> [batch] [dl|dl_adam|mf|langevin]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvm_b200 import engine as E, _native as nat  # noqa: E402
from tools.quick_bench import synth  # noqa: E402

dev = torch.device("cuda:0")
n, T = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
name = sys.argv[4] if len(sys.argv) > 4 else "dl"
hp = dict(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=False)
sid, alg, mult, kw = {
    "dl": (nat.SOLVER_DL, nat.ALG_ORIGINAL, 0.2, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, feedback_scale=100.0, g=0.05)),
    "dl_adam": (nat.SOLVER_DL, nat.ALG_ADAM, 0.2, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, g=0.05, hyperparameters=hp)),
    "mf": (nat.SOLVER_MF, nat.ALG_ORIGINAL, 0.05, dict(s=20.0, pump=0.0, dt=0.0025, j=5.0, feedback_scale=4000.0, g=0.01)),
    "langevin": (nat.SOLVER_LANGEVIN, nat.ALG_ORIGINAL, 0.05, dict(s=0.5, dt=0.002, sigma=0.5, feedback_scale=1.0)),
}[name]
q, v, f = synth(n, 0, mult, dev)
times, ref, same = [], None, True
for r in range(24):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs, _ = E.solve(sid, alg, q, v, B, T, seed=1, offset=0, **kw)
    e1.record()
    torch.cuda.synchronize()
    times.append(round(e0.elapsed_time(e1), 3))
    if ref is None:
        ref = [o.clone() for o in outs]
    else:
        same = same and all(torch.equal(a, b) for a, b in zip(ref, outs))
print(name, n, T, B, "times", times, "bitwise same", same, "finite", bool(torch.isfinite(ref[0]).all()))
