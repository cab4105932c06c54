"""Statistical-equivalence gate of the production (Philox) path on the 300 bundled BoxQP instances
(BASELINE.json configs[1]): every solver, B = 1000, T = 1500, the parameter keys of
the reference's examples, against success fractions recorded from the UNMODIFIED reference on the
CPU (tests/golden/equivalence_ref.json, written by tests/golden/make_equivalence.py).

    python tools/equivalence_gpu.py [--seed 0] [--out profiles/r1_equivalence.json]

Gate (per solver):
  * per (instance, threshold): two-proportion z-test at 95 % between the engine's and the
    reference's success fraction; at most `max_reject` (5 % + the reference's own seed-to-seed
    rejection rate) of the 300 x 7 cells may reject;
  * per size and threshold: the pooled success fraction (50 instances x B; 10 x B for the _solve_adam
    loops) must lie within the two-sample interval of the pooled reference fraction at a family-wise
    95 % over the 42 cells (|z| <= 3.24, the binomial 95 % band of the north star with its Bonferroni
    allowance);
  * where both hit the `optimal` bucket, best objective values agree within 1e-4 relative on all but
    c + 3 instances, c = the reference's own seed-to-seed disagreements under the same
    rule (the bucket is 0.1 % wide, so two runs can legitimately end in different near-optimal
    vertices); the engine is compared with each reference seed and the closer one counts.
With two reference seeds on record the reference side is their union (2 x B trajectories); seed 0
vs seed 1 goes through the same tests: that is the calibration.

How often does a CORRECT implementation fail this gate?  The pooled criterion is a 95 % test per loop:
tools/equivalence_null.py measures the statistic between independent runs of the engine itself (12 seeds,
same design) and finds worst |z| > 3.24 in 2 of 48 cases (4 %) -- and the same frequency for engine vs
reference (profiles/r2_equivalence_null.json).  A single-seed run of all eight loops therefore fails
somewhere about one time in three by chance alone.  `--seeds a,b,c` runs independent engine seeds and
applies every criterion by MAJORITY (a loop passes when it passes under more than half of the seeds):
a chance failure rate of ~0.7 % per loop with three seeds, while a real bias -- which shows under every
seed -- is still caught.  tools/noise_bias.py isolates the generator from the arithmetic.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SIZES = (20, 30, 40, 50, 60, 70)
THRESH = ("optimal", "one_percent", "two_percent", "three_percent", "four_percent", "five_percent", "ten_percent")
Z95 = 1.959964
Z_BONF42 = 3.24   # two-sided 95 % over the 42 (size, threshold) cells of a solver: Phi^-1(1 - 0.025/42)
ALL_LOOPS = ("mf", "langevin", "pumped_langevin", "dl", "mf_adam", "langevin_adam", "pumped_langevin_adam", "dl_adam")


def load_bundled(device="cuda"):
    """The 300 bundled instances as ProblemInstance objects (unscaled), from the packed fixture."""
    from ccvm_b200.problem_classes.boxqp import ProblemInstance
    z = np.load(os.path.join(GOLDEN, "bundled_instances.npz"))
    out = {}
    for n in SIZES:
        insts = []
        for k in range(z[f"q{n}"].shape[0]):
            inst = ProblemInstance(device=device, instance_type="tuning", name=str(z[f"name{n}"][k])[:-3])
            inst.problem_size = n
            inst.q_matrix = torch.from_numpy(z[f"q{n}"][k]).to(device)
            inst.v_vector = torch.from_numpy(z[f"v{n}"][k]).to(device)
            inst.optimal_sol = inst.best_sol = float(z[f"opt{n}"][k])
            inst.num_frac_values, inst.solution_vector, inst.optimality = 0, [], True
            insts.append(inst)
        out[n] = insts
    return out


def run_engine(name, key, post, seed, batch, chunk=50, adam=None, per_size=None, route="many"):
    """rows[n] = per instance [7 success fractions..., best objective] from the CUDA engine.
    `name` ending in "_adam" runs the solver's _solve_adam loop with the AdamParameters `adam` on the
    first `per_size` instances of every size; route "many" goes through solve_many (batched fused
    launches), "single" through one Solver.__call__ (one fused launch) per instance."""
    from ccvm_b200.solvers import DLSolver, MFSolver, LangevinSolver, PumpedLangevinSolver
    from ccvm_b200.solvers.algorithms import AdamParameters
    is_adam = name.endswith("_adam")
    cls = {"mf": MFSolver, "langevin": LangevinSolver, "pumped_langevin": PumpedLangevinSolver,
           "dl": DLSolver}[name[:-5] if is_adam else name]
    params = AdamParameters(**adam) if is_adam else None
    torch.manual_seed(seed)
    bundled = load_bundled()
    rows = {}
    for n in SIZES:
        solver = cls(device="cuda", batch_size=batch)
        solver.parameter_key = {n: dict(key)}
        insts = bundled[n][:per_size] if (is_adam and per_size) else bundled[n]
        for inst in insts:
            inst.scale_coefs(solver.get_scaling_factor(inst.q_matrix))
        sols = []
        if route == "single":
            sols = [solver(instance=inst, post_processor=post, algorithm_parameters=params) for inst in insts]
        else:
            for lo in range(0, len(insts), chunk):
                sols += solver.solve_many(insts[lo:lo + chunk], post_processor=post, algorithm_parameters=params)
        rows[n] = [[s.solution_performance[t] for t in THRESH] + [float(s.best_objective_value)] for s in sols]
    return rows


def compare(a, b, batch_a, batch_b=None):
    """Gate statistics between two result sets {n: rows} drawn with `batch_a` / `batch_b` trajectories
    per instance; returns a dict (see module docstring)."""
    batch_b = batch_b or batch_a
    cells = rejects = 0
    pooled = []
    best_bad = best_cmp = 0
    worst_pool = 0.0
    for n in SIZES:
        ra, rb = np.asarray(a[n], dtype=np.float64), np.asarray(b[n], dtype=np.float64)
        pa, pb = ra[:, :7], rb[:, :7]
        pm = (pa * batch_a + pb * batch_b) / (batch_a + batch_b)
        se = np.sqrt(np.maximum(pm * (1 - pm), 0.0) * (1.0 / batch_a + 1.0 / batch_b))
        # +0.5/batch: the fractions are rounded to 4 dp by the reference (solution.py:118)
        rej = np.abs(pa - pb) > Z95 * se + 0.5 / batch_a
        cells += rej.size
        rejects += int(rej.sum())
        # pooled over the 50 instances of this size (variance = sum of the per-instance binomials)
        var = (pa * (1 - pa) / batch_a + pb * (1 - pb) / batch_b).sum(axis=0) / pa.shape[0] ** 2
        diff = pa.mean(axis=0) - pb.mean(axis=0)
        zs = np.abs(diff) / np.sqrt(np.maximum(var, 1e-12))
        worst_pool = max(worst_pool, float(zs.max()))
        pooled.append({"n": n, "a": pa.mean(axis=0).round(4).tolist(), "b": pb.mean(axis=0).round(4).tolist(),
                       "z": zs.round(2).tolist()})
        both = (pa[:, 0] > 0) & (pb[:, 0] > 0)
        best_cmp += int(both.sum())
        rel = np.abs(ra[both, 7] - rb[both, 7]) / np.abs(rb[both, 7])
        best_bad += int((rel > 1e-4).sum())
    return {"cells": cells, "rejects": rejects, "reject_rate": rejects / cells, "pooled": pooled,
            "pooled_worst_z": worst_pool, "pooled_ok": worst_pool <= Z_BONF42,
            "best_compared": best_cmp, "best_mismatch": best_bad}


def merge_seeds(r0, r1):
    """Two reference runs of the same instances as one run with twice the trajectories: success
    fractions averaged, best objective = the better of the two."""
    out = {}
    for n in SIZES:
        a, b = np.asarray(r0[n], dtype=np.float64), np.asarray(r1[n], dtype=np.float64)
        m = (a + b) / 2
        m[:, 7] = np.maximum(a[:, 7], b[:, 7])
        out[n] = m.tolist()
    return out


def gate(ref, name, engine_rows, batch):
    """Engine (one run of `batch` trajectories per instance) against the reference.  With two
    reference seeds on record the reference is their union (2 x batch trajectories: half the
    sampling noise on that side) and seed 0 vs seed 1 is reported as the calibration: the
    reference's own rejection rate, pooled z and best-objective disagreements under the same tests."""
    ref0 = {n: ref[f"{name}/seed0/{n}"] for n in SIZES}
    out = {}
    if f"{name}/seed1/{SIZES[-1]}" in ref:
        ref1 = {n: ref[f"{name}/seed1/{n}"] for n in SIZES}
        out["engine_vs_ref"] = compare(engine_rows, merge_seeds(ref0, ref1), batch, 2 * batch)
        out["ref_seed0_vs_seed1"] = compare(ref1, ref0, batch)
        out["engine_vs_ref_seed0"] = compare(engine_rows, ref0, batch)
        out["engine_vs_ref_seed1"] = compare(engine_rows, ref1, batch)
    else:
        out["engine_vs_ref"] = compare(engine_rows, ref0, batch)
    cal = out.get("ref_seed0_vs_seed1", {})
    e = out["engine_vs_ref"]
    out["max_reject"] = 0.05 + cal.get("reject_rate", 0.0)
    c = cal.get("best_mismatch", 0)
    out["max_best_mismatch"] = int(c + 3)
    out["best_mismatch"] = min(out[k]["best_mismatch"] for k in ("engine_vs_ref_seed0", "engine_vs_ref_seed1")
                               if k in out) if "engine_vs_ref_seed0" in out else e["best_mismatch"]
    out["pass"] = bool(e["reject_rate"] <= out["max_reject"] and e["pooled_ok"]
                       and out["best_mismatch"] <= out["max_best_mismatch"])
    return out


def majority_gate(runs):
    """A loop passes when EACH criterion holds under more than half of the independent engine seeds."""
    need = len(runs) // 2 + 1
    crit = [sum(r["engine_vs_ref"]["reject_rate"] <= r["max_reject"] for r in runs),
            sum(bool(r["engine_vs_ref"]["pooled_ok"]) for r in runs),
            sum(r["best_mismatch"] <= r["max_best_mismatch"] for r in runs)]
    return all(c >= need for c in crit)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--seeds", default="", help="comma list of engine seeds: majority gate over independent runs")
    ap.add_argument("--solvers", default=",".join(ALL_LOOPS))
    ap.add_argument("--route", default="many", choices=["many", "single"])
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    ref = json.load(open(os.path.join(GOLDEN, "equivalence_ref.json")))
    meta = ref["_meta"]
    report = {"batch": meta["batch"], "iterations": meta["iterations"], "engine_seed": args.seed, "route": args.route,
              "z_pooled_max": Z_BONF42,
              "reference": {"torch": meta["torch"], "device": meta["device"]}}
    seeds = [int(x) for x in args.seeds.split(",")] if args.seeds else [args.seed]
    report["engine_seeds"] = seeds
    ok = True
    for name in args.solvers.split(","):
        if f"{name}/seed0/{SIZES[-1]}" not in ref:
            print(f"{name}: no reference record, skipped")
            continue
        runs = []
        for seed in seeds:
            rows = run_engine(name, meta["keys"][name], meta["post_processor"][name], seed, meta["batch"],
                              adam=meta.get("adam"), per_size=meta.get("adam_per_size"), route=args.route)
            g = gate(ref, name, rows, meta["batch"])
            g["engine_seed"] = seed
            runs.append(g)
            e, c = g["engine_vs_ref"], g.get("ref_seed0_vs_seed1")
            print(f"{name} seed {seed}: pass={g['pass']} cell rejections {e['rejects']}/{e['cells']} ({100 * e['reject_rate']:.2f} %)"
                  f" pooled worst z {e['pooled_worst_z']:.2f}; best mismatches {g['best_mismatch']} (allowed {g['max_best_mismatch']})"
                  + (f" | reference seed0 vs seed1: {100 * c['reject_rate']:.2f} %, worst z {c['pooled_worst_z']:.2f}" if c else ""))
        verdict = majority_gate(runs)
        report[name] = {"pass": verdict, "runs": runs} if len(runs) > 1 else runs[0]
        if len(runs) > 1:
            print(f"{name}: majority over seeds {seeds}: pass={verdict}")
        ok = ok and verdict
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(report, open(args.out, "w"), indent=1)
    print("EQUIVALENCE", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
