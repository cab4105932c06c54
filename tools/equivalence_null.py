"""Calibration of the statistical-equivalence gate (tools/equivalence_gpu.py): the distribution of its
pooled worst-|z| statistic (a) between INDEPENDENT RUNS OF THE ENGINE ITSELF (one run vs the union of two
others -- same design as engine vs the two reference seeds, so this is the statistic's null distribution
with every correlation of the design in it: nested thresholds, shared instances) and (b) between engine
runs and the recorded reference.  If (b) looks like (a), the engine is statistically indistinguishable
from the reference at this sample size.

    python tools/equivalence_null.py [--solvers langevin,mf] [--seeds 9] [--out profiles/...json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import equivalence_gpu as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--solvers", default="langevin,mf,dl")
    ap.add_argument("--seeds", type=int, default=9)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    ref = json.load(open(os.path.join(G.GOLDEN, "equivalence_ref.json")))
    meta = ref["_meta"]
    batch = meta["batch"]
    report = {"batch": batch, "seeds": args.seeds, "z_gate": G.Z_BONF42}
    for name in args.solvers.split(","):
        runs = [G.run_engine(name, meta["keys"][name], meta["post_processor"][name], 100 + s, batch,
                             adam=meta.get("adam"), per_size=meta.get("adam_per_size")) for s in range(args.seeds)]
        ref0 = {n: ref[f"{name}/seed0/{n}"] for n in G.SIZES}
        ref1 = {n: ref[f"{name}/seed1/{n}"] for n in G.SIZES}
        ref_union = G.merge_seeds(ref0, ref1)
        null, null_rej = [], []
        for i in range(args.seeds):        # run i against the union of the next two runs (cyclic): the null
            j, k = (i + 1) % args.seeds, (i + 2) % args.seeds
            c = G.compare(runs[i], G.merge_seeds(runs[j], runs[k]), batch, 2 * batch)
            null.append(c["pooled_worst_z"])
            null_rej.append(c["reject_rate"])
        vs_ref, vs_ref_rej = [], []
        for i in range(args.seeds):
            c = G.compare(runs[i], ref_union, batch, 2 * batch)
            vs_ref.append(c["pooled_worst_z"])
            vs_ref_rej.append(c["reject_rate"])
        # the reference against the engine taken as "the reference": union of two engine runs
        ref_vs_engine = [G.compare(r, G.merge_seeds(runs[0], runs[1]), batch, 2 * batch)["pooled_worst_z"] for r in (ref0, ref1)]
        row = {"engine_vs_engine_worst_z": np.round(null, 2).tolist(), "engine_vs_reference_worst_z": np.round(vs_ref, 2).tolist(),
               "engine_vs_engine_reject_rate": np.round(null_rej, 4).tolist(),
               "engine_vs_reference_reject_rate": np.round(vs_ref_rej, 4).tolist(),
               "reference_seed_vs_engine_union_worst_z": np.round(ref_vs_engine, 2).tolist(),
               "reference_seed0_vs_seed1_worst_z": round(G.compare(ref1, ref0, batch)["pooled_worst_z"], 2)}
        report[name] = row
        print(name, json.dumps(row), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(report, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
