"""Development aid: per-iteration clock stamps of the small-n tensor-core kernel (library built with
-DCCVM_MMA_TRACE -DCCVM_INST_... for ONE tile; usage: CCVM_B200_LIB=build/alt/libtrace.so mma_trace.py <solver> <adam>)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CCVM_MMA"] = "1"
import torch  # noqa: E402
from ccvm_b200 import engine as E, _native as nat  # noqa: E402
from tools.quick_bench import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "langevin"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 70
hp = dict(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=False)
cases = {
    "dl": (nat.SOLVER_DL, nat.ALG_ORIGINAL, 0.2, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, feedback_scale=100.0, g=0.05)),
    "dl_adam": (nat.SOLVER_DL, nat.ALG_ADAM, 0.2, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, g=0.05, hyperparameters=hp)),
    "langevin": (nat.SOLVER_LANGEVIN, nat.ALG_ORIGINAL, 0.05, dict(s=0.5, dt=0.002, sigma=0.5, feedback_scale=1.0)),
}
sid, alg, mult, kw = cases[name]
dev = torch.device("cuda", 0)
q, v, f = synth(n, 0, mult, dev)
for w in range(3):
    E.solve(sid, alg, q, v, 4096, 300, seed=1, offset=w, **kw)
torch.cuda.synchronize()
buf = (C.c_longlong * 512)()
lib = nat.load()
lib.ccvm_debug_mma_trace.argtypes = [C.POINTER(C.c_longlong)]
rc = lib.ccvm_debug_mma_trace(buf)
rows = [[buf[i * 16 + s] for s in range(16)] for i in range(32)]
print("rc", rc, "per warpgroup g: iter start | noise done | D ready | D loaded | staged | arrived (relative to the iteration start of"
      " warpgroup 0); issuer: woke / issued relative to that warpgroup's arrival")
for i in range(1, 31):
    r = rows[i]
    base = r[0]
    line = [f"{i + 64} period {r[0] - rows[i - 1][0]:5d}"]
    for g in (0, 1):
        o = 8 * g
        line.append(f"wg{g} " + " ".join(f"{r[o + k] - base:5d}" for k in range(6)))
        line.append(f"issuer(next) +{rows[i + 1][o + 6] - r[o + 5]:4d} +{rows[i + 1][o + 7] - r[o + 5]:4d}")
    print(" | ".join(line))
