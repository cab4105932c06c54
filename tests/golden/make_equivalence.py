"""Reference side of the statistical-equivalence gate (BASELINE.json configs[1]).

    python tests/golden/make_equivalence.py pack            # bundled instances -> one .npz
    python tests/golden/make_equivalence.py run  [--seeds 0,1] [--solvers mf,langevin,...] [--threads 6]

Needs /root/reference (build container only).  `pack` loads the 300 bundled BoxQP instances
(examples/benchmarking_instances/Size{20..70}/tuningH0NN-100-{0..49}.in) with the reference's own
ProblemInstance and stores the UNSCALED coefficients as the reference holds them (negated on load,
problem_instance.py:183-188), the optimum and the file names in `bundled_instances.npz`.
`run` calls the UNMODIFIED reference solvers on every instance on the CPU (B = 1000, T = 1500, the
parameter keys of examples/*.py reused for every size, grad-descent post-processor for
MF / Langevin / PumpedLangevin and none for DL, as the examples do) under `torch.manual_seed(seed)`
for two seeds and stores, per (solver, seed, instance): the 7 success fractions of
solution.py:130-136 and the best objective value -> `equivalence_ref.json`.  The second seed is the
reference's own seed-to-seed spread, the calibration of the gate.
"""
import argparse
import contextlib
import glob
import io
import json
import os
import sys
import time

import numpy as np
import torch

REF = os.environ.get("CCVM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
SIZES = (20, 30, 40, 50, 60, 70)
BATCH, ITERS = 1000, 1500

KEYS = {
    "mf": dict(pump=0.0, feedback_scale=4000, j=5.0, S=20.0, dt=0.0025, iterations=ITERS),
    "langevin": dict(dt=0.002, S=0.5, sigma=0.5, feedback_scale=1.0, iterations=ITERS),
    "pumped_langevin": dict(pump=2.0, dt=0.002, S=0.5, sigma=0.5, feedback_scale=1.0, iterations=ITERS),
    "dl": dict(pump=8.0, feedback_scale=100, dt=0.001, iterations=ITERS, noise_ratio=10),
}
POST = {"mf": "grad-descent", "langevin": "grad-descent", "pumped_langevin": "grad-descent", "dl": None}
# `_solve_adam` loops (round 2): the same keys and post-processors, run with the AdamParameters of the
# reference's examples (examples/ccvm_boxqp_dl.py:45-47) on the first ADAM_PER_SIZE instances of a size
ADAM = dict(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=False)
ADAM_PER_SIZE = 10
for _name in ("mf", "langevin", "pumped_langevin", "dl"):
    KEYS[_name + "_adam"] = KEYS[_name]
    POST[_name + "_adam"] = POST[_name]


def instance_files(n):
    d = f"{REF}/examples/benchmarking_instances/Size{n}"
    return sorted(glob.glob(f"{d}/*.in"), key=lambda p: int(os.path.basename(p).rsplit("-", 1)[1][:-3]))


def pack():
    sys.path.insert(0, REF)
    from ccvm_simulators.problem_classes.boxqp import ProblemInstance
    out = {}
    for n in SIZES:
        qs, vs, opts, names = [], [], [], []
        for path in instance_files(n):
            inst = ProblemInstance(instance_type="tuning", file_path=path, device="cpu")
            assert inst.problem_size == n
            qs.append(inst.q_matrix.numpy().copy())
            vs.append(inst.v_vector.numpy().copy())
            opts.append(float(inst.optimal_sol))
            names.append(os.path.basename(path))
        out[f"q{n}"] = np.stack(qs).astype(np.float32)
        out[f"v{n}"] = np.stack(vs).astype(np.float32)
        out[f"opt{n}"] = np.asarray(opts, dtype=np.float64)
        out[f"name{n}"] = np.asarray(names)
        print(n, out[f"q{n}"].shape, names[0], names[-1])
    np.savez_compressed(os.path.join(HERE, "bundled_instances.npz"), **out)


def run(seeds, solvers, threads):
    sys.path.insert(0, REF)
    torch.set_num_threads(threads)
    from ccvm_simulators.problem_classes.boxqp import ProblemInstance
    from ccvm_simulators.solvers import DLSolver, MFSolver, LangevinSolver, PumpedLangevinSolver
    from ccvm_simulators.solvers.algorithms import AdamParameters

    class DLSolverAdamCallable(DLSolver):
        """The reference's DLSolver.__call__ hands feedback_scale to _solve_adam, which does not take it
        (TypeError); this adapter drops that one positional argument and nothing else."""

        def _solve_adam(self, problem_size, batch_size, device, S, pump, dt, iterations, noise_ratio, feedback_scale,
                        *rest):
            return DLSolver._solve_adam(self, problem_size, batch_size, device, S, pump, dt, iterations, noise_ratio,
                                        *rest)

    cls = {"mf": MFSolver, "langevin": LangevinSolver, "pumped_langevin": PumpedLangevinSolver, "dl": DLSolver,
           "mf_adam": MFSolver, "langevin_adam": LangevinSolver, "pumped_langevin_adam": PumpedLangevinSolver,
           "dl_adam": DLSolverAdamCallable}
    path_out = os.path.join(HERE, "equivalence_ref.json")
    res = json.load(open(path_out)) if os.path.exists(path_out) else {}
    res["_meta"] = {"batch": BATCH, "iterations": ITERS, "keys": KEYS, "post_processor": POST,
                    "torch": torch.__version__, "device": "cpu", "adam": ADAM, "adam_per_size": ADAM_PER_SIZE,
                    "what": "unmodified reference Solver.__call__ under torch.manual_seed(seed); "
                            "per instance: [optimal, 1%, 2%, 3%, 4%, 5%, 10%] success fractions + best objective"}
    for name in solvers:
        for seed in seeds:
            tag = f"{name}/seed{seed}"
            for n in SIZES:
                key = f"{tag}/{n}"
                if key in res:
                    continue
                solver = cls[name](device="cpu", batch_size=BATCH)
                solver.parameter_key = {n: dict(KEYS[name])}
                rows = []
                t0 = time.time()
                adam = name.endswith("_adam")
                files = instance_files(n)[:ADAM_PER_SIZE] if adam else instance_files(n)
                for k, path in enumerate(files):
                    inst = ProblemInstance(instance_type="tuning", file_path=path, device="cpu")
                    inst.scale_coefs(solver.get_scaling_factor(inst.q_matrix))
                    torch.manual_seed(seed * 1000 + k)
                    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                        sol = solver(instance=inst, post_processor=POST[name],
                                     algorithm_parameters=AdamParameters(**ADAM) if adam else None)
                    p = sol.solution_performance
                    rows.append([p["optimal"], p["one_percent"], p["two_percent"], p["three_percent"],
                                 p["four_percent"], p["five_percent"], p["ten_percent"],
                                 float(sol.best_objective_value)])
                res[key] = rows
                print(f"{key}: {time.time() - t0:.1f}s  mean p_opt={np.mean([r[0] for r in rows]):.3f}", flush=True)
                json.dump(res, open(path_out, "w"))
    json.dump(res, open(path_out, "w"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["pack", "run"])
    ap.add_argument("--seeds", default="0,1")
    ap.add_argument("--solvers", default="mf,langevin,pumped_langevin,dl")
    ap.add_argument("--threads", type=int, default=6)
    a = ap.parse_args()
    if a.cmd == "pack":
        pack()
    else:
        run([int(s) for s in a.seeds.split(",")], a.solvers.split(","), a.threads)
