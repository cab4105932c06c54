"""Production (Philox) mode against the UNMODIFIED reference on ALL 300 bundled BoxQP instances
(BASELINE.json configs[1], SURVEY.md 8d "statistical-equivalence gate"): B = 1000, T = 1500, the
parameter keys of the reference's examples.  The reference's success fractions were recorded on the
CPU by tests/golden/make_equivalence.py (two seeds: the second calibrates the gate with the
reference's own seed-to-seed spread); the engine runs here through solve_many."""
import json
import os

import pytest

from tools import equivalence_gpu as G

pytestmark = pytest.mark.gpu

REF = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "equivalence_ref.json")))
SOLVERS = [s for s in G.ALL_LOOPS if f"{s}/seed0/70" in REF]


ENGINE_SEEDS = (0, 1, 2)


@pytest.mark.parametrize("route", ["many", "single", "single-mma"])
@pytest.mark.parametrize("name", SOLVERS)
def test_statistically_equivalent_on_bundled_instances(monkeypatch, name, route):
    """All eight loops (the _solve_adam ones on the first 10 instances of every size, with the
    AdamParameters of the reference's examples), through batched launches (solve_many) AND through
    one Solver.__call__ per instance (the single-launch kernel instantiations).

    Every criterion of the gate is a 95 % test, so a correct implementation fails a single-seed run of one
    loop about 4 % of the time (measured: tools/equivalence_null.py, engine vs engine).  The test runs three
    independent engine seeds and applies the criteria by majority (tools/equivalence_gpu.py::majority_gate)."""
    meta = REF["_meta"]
    if route == "single-mma":
        # the small-n tensor-core kernel (csrc/sde_kernel_mma.cuh) forced for every size of the bundled set (it serves
        # n >= 40 at batch >= 2048 by default): the four _solve_adam loops here, all eight in
        # profiles/r2y_equivalence_mma_kernel_single_3seeds.log
        monkeypatch.setenv("CCVM_MMA", "1")
        route = "single"
    if route == "single" and not name.endswith("_adam"):
        pytest.skip("300 single calls per solver and seed: run with tools/equivalence_gpu.py --route single")
    runs = []
    for seed in ENGINE_SEEDS:
        rows = G.run_engine(name, meta["keys"][name], meta["post_processor"][name], seed=seed, batch=meta["batch"],
                            adam=meta.get("adam"), per_size=meta.get("adam_per_size"), route=route)
        runs.append(G.gate(REF, name, rows, meta["batch"]))
    summary = [(round(r["engine_vs_ref"]["reject_rate"], 4), round(r["engine_vs_ref"]["pooled_worst_z"], 2),
                r["best_mismatch"], r["max_best_mismatch"]) for r in runs]
    assert G.majority_gate(runs), (name, route, "per seed (reject rate, pooled worst z, best mismatches, allowed)", summary)
