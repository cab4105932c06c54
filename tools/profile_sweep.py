"""Development aid: cProfile of CCVMSolver.solve_many on the bundled instances (host-side overhead)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvm_b200.solvers import LangevinSolver  # noqa: E402
from tools.equivalence_gpu import load_bundled, SIZES  # noqa: E402

bundled = load_bundled()
solver = LangevinSolver(device="cuda", batch_size=1000)
solver.parameter_key = {n: dict(dt=0.002, S=0.5, sigma=0.5, feedback_scale=1.0, iterations=1500) for n in SIZES}
insts = []
for n in SIZES:
    for inst in bundled[n]:
        inst.scale_coefs(solver.get_scaling_factor(inst.q_matrix))
        insts.append(inst)
solver.solve_many(insts[:50], post_processor="grad-descent")
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for lo in range(0, len(insts), 50):
    solver.solve_many(insts[lo:lo + 50], post_processor="grad-descent")
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
