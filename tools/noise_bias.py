"""Where does a success-fraction difference between the engine's production mode and the reference come
from?  (VERDICT r1 "what's weak" 2.)  On bundled instances, Langevin / PumpedLangevin keys, the SAME
engine arithmetic is driven by three noise sources with many trajectories per instance:

  A  production mode (in-kernel generator: Philox-seeded xoshiro128+ streams + fast-math Box-Muller)
  B  replay of torch.randn drawn on the GPU (cuRAND Philox + exact Box-Muller)
  D  replay of torch.randn drawn on the CPU the way the reference draws it (mt19937 + torch's CPU normal)

and the 7 success fractions of solution.py:125-136 are compared pairwise with two-proportion z-tests.
A vs B isolates the engine's generator; B vs D the two torch generators; any of them vs the recorded
reference fractions (tests/golden/equivalence_ref.json, B = 1000 x 2 seeds) the arithmetic.

    python tools/noise_bias.py [--per-size 2] [--chunks 5] [--chunk-batch 10000] [--out profiles/...json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ccvm_b200 import engine as E, _native as nat  # noqa: E402
from tools.equivalence_gpu import load_bundled, THRESH  # noqa: E402

KEYS = {
    "langevin": (nat.SOLVER_LANGEVIN, dict(s=0.5, dt=0.002, sigma=0.5, feedback_scale=1.0)),
    "pumped_langevin": (nat.SOLVER_PUMPED_LANGEVIN, dict(s=0.5, pump=2.0, dt=0.002, sigma=0.5, feedback_scale=1.0)),
}
ITERS = 1500


def counts_of(state, inst, want):
    """7 success counters of one batch through the engine's fused epilogue kernels (grad-descent)."""
    pv, en = E.epilogue(state, inst.q_matrix, inst.v_vector, map1=(1.0, 0.5), post_processor="grad-descent",
                        pp_iterations=10, scaled_by=float(inst.scaled_by))
    _, _, counts = E.solution_stats(en, inst.optimal_sol)
    return np.asarray(counts, dtype=np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="20,40,70")
    ap.add_argument("--per-size", type=int, default=2)
    ap.add_argument("--chunks", type=int, default=5)
    ap.add_argument("--chunk-batch", type=int, default=10000)
    ap.add_argument("--cpu-noise-sizes", default="20,70", help="sizes that also get arm D (CPU-drawn noise is slow)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda")
    bundled = load_bundled()
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "equivalence_ref.json")))
    from ccvm_b200.solvers import LangevinSolver
    scaler = LangevinSolver(device="cuda")
    report = {"iterations": ITERS, "trajectories_per_arm": args.chunks * args.chunk_batch, "cells": []}
    cpu_sizes = {int(x) for x in args.cpu_noise_sizes.split(",") if x}
    for n in (int(x) for x in args.sizes.split(",")):
        for k in range(args.per_size):
            inst = bundled[n][k]
            inst.scale_coefs(scaler.get_scaling_factor(inst.q_matrix))
            nb = args.chunk_batch
            tots = {name: {"A": np.zeros(7, np.int64), "B": np.zeros(7, np.int64), "D": np.zeros(7, np.int64)}
                    for name in KEYS}
            for c in range(args.chunks):
                g = torch.Generator(device=dev).manual_seed(5000 + 31 * c + k)
                noise_b = torch.randn((ITERS, 1, n, nb), device=dev, generator=g)
                noise_d = None
                if n in cpu_sizes:
                    gc = torch.Generator().manual_seed(9000 + 31 * c + k)
                    # the reference's draw: one randn(N, B) per iteration from the CPU generator
                    host = torch.empty((ITERS, 1, n, nb), dtype=torch.float32).pin_memory()
                    for t in range(ITERS):
                        torch.randn((n, nb), generator=gc, out=host[t, 0])
                    noise_d = host.to(dev)
                for name, (sid, kw) in KEYS.items():
                    tot = tots[name]
                    outs, _ = E.solve(sid, nat.ALG_ORIGINAL, inst.q_matrix, inst.v_vector, nb, ITERS, seed=1000 + c,
                                      offset=17 * k, **kw)
                    tot["A"] += counts_of(outs[0], inst, 0)
                    outs, _ = E.solve(sid, nat.ALG_ORIGINAL, inst.q_matrix, inst.v_vector, nb, ITERS, noise=noise_b, **kw)
                    tot["B"] += counts_of(outs[0], inst, 0)
                    if noise_d is not None:
                        outs, _ = E.solve(sid, nat.ALG_ORIGINAL, inst.q_matrix, inst.v_vector, nb, ITERS, noise=noise_d,
                                          **kw)
                        tot["D"] += counts_of(outs[0], inst, 0)
                del noise_b, noise_d
            for name in KEYS:
                tot = tots[name]
                m = args.chunks * nb
                frac = {a: (tot[a] / m) for a in tot}
                rows = [ref.get(f"{name}/seed{s}/{n}") for s in (0, 1)]
                ref_frac = np.mean([np.asarray(r[k][:7]) for r in rows if r], axis=0) if rows[0] else None

                def z(pa, pb, ma, mb):
                    pm = (pa * ma + pb * mb) / (ma + mb)
                    se = np.sqrt(np.maximum(pm * (1 - pm), 1e-12) * (1 / ma + 1 / mb))
                    return ((pa - pb) / se).round(2).tolist()

                cell = {"solver": name, "n": n, "instance": k, "thresholds": list(THRESH),
                        "A_production": frac["A"].round(5).tolist(), "B_replay_cuda_randn": frac["B"].round(5).tolist(),
                        "z_A_vs_B": z(frac["A"], frac["B"], m, m)}
                if n in cpu_sizes:
                    cell["D_replay_cpu_randn"] = frac["D"].round(5).tolist()
                    cell["z_A_vs_D"] = z(frac["A"], frac["D"], m, m)
                    cell["z_B_vs_D"] = z(frac["B"], frac["D"], m, m)
                if ref_frac is not None:
                    cell["reference_recorded"] = ref_frac.round(5).tolist()
                    cell["z_A_vs_reference"] = z(frac["A"], ref_frac, m, 2000)
                    cell["z_B_vs_reference"] = z(frac["B"], ref_frac, m, 2000)
                report["cells"].append(cell)
                print(json.dumps(cell), flush=True)
    for key in ("z_A_vs_B", "z_A_vs_D", "z_B_vs_D", "z_A_vs_reference", "z_B_vs_reference"):
        zs = np.asarray([v for c in report["cells"] if key in c for v in c[key]])
        if zs.size:
            report[key + "_summary"] = {"cells": int(zs.size), "mean": float(zs.mean()), "rms": float(np.sqrt((zs ** 2).mean())),
                                        "max_abs": float(np.abs(zs).max())}
            print(key, report[key + "_summary"], flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(report, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
