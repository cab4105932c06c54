"""Time-to-solution on bundled BoxQP instances: this engine on the GPU next to the UNMODIFIED
reference on the box's host CPU, same instances, same parameter keys, same TTS definition
(TTS = mean per-run solve_time x bootstrapped R99 of the `optimal` success fraction,
boxqp_metadata.py:117-135 / sampleTTSmetric.py:123-214; ccvm_b200/tts.py).

    python tools/tts_compare.py [--per-size 3] [--solvers langevin,mf] [--out profiles/...json]

Instances come from tests/golden/bundled_instances.npz (the first --per-size files of every
Size20..Size70 folder); the reference is imported from baseline/_ref (it travels with the repo to
the GPU box; without it only the engine column is produced).  B = 1000, T = 1500.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ccvm_b200 import tts  # noqa: E402
from tools.equivalence_gpu import load_bundled, SIZES, GOLDEN  # noqa: E402

BATCH, ITERS = 1000, 1500
KEYS = {
    "mf": (dict(pump=0.0, feedback_scale=4000, j=5.0, S=20.0, dt=0.0025, iterations=ITERS), "grad-descent"),
    "langevin": (dict(dt=0.002, S=0.5, sigma=0.5, feedback_scale=1.0, iterations=ITERS), "grad-descent"),
    "pumped_langevin": (dict(pump=2.0, dt=0.002, S=0.5, sigma=0.5, feedback_scale=1.0, iterations=ITERS), "grad-descent"),
    "dl": (dict(pump=8.0, feedback_scale=100, dt=0.001, iterations=ITERS, noise_ratio=10), None),
}
CLS = {"mf": "MFSolver", "langevin": "LangevinSolver", "pumped_langevin": "PumpedLangevinSolver", "dl": "DLSolver"}


def engine_records(name, per_size, batched):
    import ccvm_b200.solvers as S
    key, pp = KEYS[name]
    bundled = load_bundled()
    recs = []
    torch.manual_seed(0)
    for n in SIZES:
        solver = getattr(S, CLS[name])(device="cuda", batch_size=BATCH)
        solver.parameter_key = {n: dict(key)}
        insts = bundled[n][:per_size]
        for inst in insts:
            inst.scale_coefs(solver.get_scaling_factor(inst.q_matrix))
        solver(instance=insts[0], post_processor=pp)  # warm-up (module load, allocator)
        sols = solver.solve_many(insts, post_processor=pp) if batched else \
            [solver(instance=i, post_processor=pp) for i in insts]
        recs += [s.get_metadata_dict() for s in sols]
    return recs


def reference_records(name, per_size, threads):
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "ccvm_simulators")):
        return None
    sys.path.insert(0, ref_dir)
    try:
        import ccvm_simulators.solvers as RS
        from ccvm_simulators.problem_classes.boxqp import ProblemInstance as RefInstance
    except Exception as e:  # noqa: BLE001
        print(f"reference not importable: {e}", file=sys.stderr)
        return None
    finally:
        sys.path.remove(ref_dir)
    torch.set_num_threads(threads)
    z = np.load(os.path.join(GOLDEN, "bundled_instances.npz"))
    key, pp = KEYS[name]
    recs = []
    torch.manual_seed(0)
    for n in SIZES:
        solver = getattr(RS, CLS[name])(device="cpu", batch_size=BATCH)
        solver.parameter_key = {n: dict(key)}
        for k in range(per_size):
            inst = RefInstance(instance_type="tuning", device="cpu", name=str(z[f"name{n}"][k])[:-3])
            inst.problem_size = n
            inst.q_matrix = torch.from_numpy(z[f"q{n}"][k].copy())
            inst.v_vector = torch.from_numpy(z[f"v{n}"][k].copy())
            inst.optimal_sol = inst.best_sol = float(z[f"opt{n}"][k])
            inst.num_frac_values, inst.solution_vector, inst.optimality = 0, [], True
            inst.scale_coefs(solver.get_scaling_factor(inst.q_matrix))
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                sol = solver(instance=inst, post_processor=pp)
            recs.append(sol.get_metadata_dict())
    return recs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-size", type=int, default=3)
    ap.add_argument("--solvers", default="langevin,pumped_langevin,mf,dl")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    report = {"batch": BATCH, "iterations": ITERS, "per_size": args.per_size, "cpu_threads": args.threads,
              "definition": "TTS[s] = mean(solve_time per trajectory) x mean bootstrapped R99(50th percentile) of the "
                            "'optimal' (<= 0.1 % gap) success fraction; inf when no instance of the size was solved"}
    for name in args.solvers.split(","):
        t0 = time.time()
        ref = reference_records(name, args.per_size, args.threads)
        t_ref = time.time() - t0
        row = {}
        for label, recs in (("reference_cpu", ref), ("engine_sequential", engine_records(name, args.per_size, False)),
                            ("engine_batched", engine_records(name, args.per_size, True))):
            if recs is None:
                continue
            table = tts.tts_table(recs, percentiles=(50.0,))
            row[label] = {
                "tts_s": {str(n): table[n][50.0] for n in table},
                "mean_solve_time_s": {str(n): float(np.mean([r["solve_time"] for r in recs if r["problem_size"] == n]))
                                      for n in table},
                "mean_p_optimal": {str(n): float(np.mean([r["solution_performance"]["optimal"] for r in recs
                                                          if r["problem_size"] == n])) for n in table},
            }
        if "reference_cpu" in row:
            row["tts_speedup_sequential"] = {
                n: (row["reference_cpu"]["tts_s"][n] / row["engine_sequential"]["tts_s"][n]
                    if np.isfinite(row["reference_cpu"]["tts_s"][n]) and np.isfinite(row["engine_sequential"]["tts_s"][n])
                    else None) for n in row["reference_cpu"]["tts_s"]}
            row["reference_wall_s"] = t_ref
        report[name] = row
        print(json.dumps({name: row}), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(report, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
