"""Development aid: default-plan timing of one solver loop over a range of sizes (anomaly scan)."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvm_b200 import engine as E, _native as nat  # noqa: E402
from tools.quick_bench import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="8,12,16,20,24,28,32,36,40,44,48,52,56,60,64,68,72,80,88,96,104,112,120,128")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--iters", type=int, default=300)
ap.add_argument("--solvers", default="dl,langevin")
a = ap.parse_args()
dev = torch.device("cuda:0")
hp = dict(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=False)
cases = {
    "dl": (nat.SOLVER_DL, nat.ALG_ORIGINAL, 0.2, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, feedback_scale=100.0, g=0.05)),
    "dl_adam": (nat.SOLVER_DL, nat.ALG_ADAM, 0.2, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, g=0.05, hyperparameters=hp)),
    "mf": (nat.SOLVER_MF, nat.ALG_ORIGINAL, 0.05, dict(s=20.0, pump=0.0, dt=0.0025, j=5.0, feedback_scale=4000.0, g=0.01)),
    "langevin": (nat.SOLVER_LANGEVIN, nat.ALG_ORIGINAL, 0.05, dict(s=0.5, dt=0.002, sigma=0.5, feedback_scale=1.0)),
}
for name in a.solvers.split(","):
    sid, alg, mult, kw = cases[name]
    for n in [int(x) for x in a.sizes.split(",")]:
        q, v, f = synth(n, 0, mult, dev)
        plan = E.plan_solve(sid, alg, q, v, a.batch, a.iters, seed=1, offset=0, **kw)
        info = E.query_launch(plan.desc)
        for w in range(2):
            E.solve(sid, alg, q, v, a.batch, a.iters, seed=1, offset=w, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(3):
            E.solve(sid, alg, q, v, a.batch, a.iters, seed=1, offset=10 + r, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        us_it = ms * 1e3 / a.iters
        print(json.dumps({"solver": name, "n": n, "ms": round(ms, 4), "us_per_iter": round(us_it, 3),
                          "steps_per_s": a.batch * a.iters / ms * 1e3, "launch": info}), flush=True)
