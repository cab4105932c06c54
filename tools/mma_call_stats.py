"""Development aid: DLSolver.__call__ (adam + adam post-processor) on a bundled N = 70 instance through the small-n
tensor-core kernel (CCVM_MMA=1) and the tiled kernel (CCVM_MMA=0), several torch seeds each: moments, upper quantiles
and the best objective of the 4096 trajectories -- the calibration of tests/test_gpu_mma.py::test_solver_call_through_tensor_core_path."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ccvm_b200.solvers import DLSolver  # noqa: E402
from ccvm_b200.solvers.algorithms import AdamParameters  # noqa: E402
from tools.equivalence_gpu import load_bundled  # noqa: E402

inst = load_bundled()[70][3]
hp = AdamParameters(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=False)
scaled = False
for seed in (5, 6, 7, 8):
    for mode in ("1", "0"):
        os.environ["CCVM_MMA"] = mode
        torch.manual_seed(seed)
        solver = DLSolver(device="cuda", batch_size=4096)
        solver.parameter_key = {70: dict(pump=8.0, dt=0.001, iterations=1500, noise_ratio=10, feedback_scale=100)}
        if not scaled:
            inst.scale_coefs(solver.get_scaling_factor(inst.q_matrix))
            scaled = True
        r = solver(instance=inst, post_processor="adam", algorithm_parameters=hp)
        o = r.objective_values.double().cpu()
        qs = torch.quantile(o, torch.tensor([0.5, 0.9, 0.99, 0.999], dtype=torch.float64)).tolist()
        print(json.dumps({"seed": seed, "mma": mode, "mean": o.mean().item(), "std": o.std().item(), "q50_90_99_999": qs,
                          "top5": torch.topk(o, 5).values.tolist(), "best": r.best_objective_value,
                          "optimal": float(inst.optimal_sol) if hasattr(inst, "optimal_sol") else None}), flush=True)
