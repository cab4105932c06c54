import sys, os, json, torch
sys.path.insert(0, os.getcwd())
from ccvm_b200 import engine as E, _native as nat
from tools.quick_bench import synth
dev = torch.device("cuda:0")
n = int(sys.argv[1]); T = int(sys.argv[2]); B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
kw = dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, feedback_scale=100.0, g=0.05)
q, v, f = synth(n, 0, 0.2, dev)
times = []; ref = None; same = True
for r in range(24):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs, _ = E.solve(nat.SOLVER_DL, nat.ALG_ORIGINAL, q, v, B, T, seed=1, offset=0, **kw)
    e1.record(); torch.cuda.synchronize()
    times.append(round(e0.elapsed_time(e1), 3))
    if ref is None: ref = [o.clone() for o in outs]
    else: same = same and all(torch.equal(a, b) for a, b in zip(ref, outs))
print(n, T, B, "times", times, "bitwise same", same, "finite", bool(torch.isfinite(ref[0]).all()))
