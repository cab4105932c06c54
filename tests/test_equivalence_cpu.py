"""CPU checks of the statistical-equivalence gate (tools/equivalence_gpu.py) and of the packed
bundled-instance fixture it runs on (tests/golden/bundled_instances.npz, written by
tests/golden/make_equivalence.py from the reference's examples/benchmarking_instances)."""
import json
import os

import numpy as np

from tools import equivalence_gpu as G

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rows(rng, p_inst, batch, shift=0.0):
    """Synthetic per-instance records: nested thresholds drawn from one binomial per instance."""
    out = {}
    for n in G.SIZES:
        rows = []
        for p in p_inst:
            base = np.clip(p + shift, 0, 1)
            ps = np.clip(base + np.linspace(0, 0.3, 7), 0, 1)
            fr = [round(rng.binomial(batch, q) / batch, 4) for q in ps]
            rows.append(fr + [100.0 + n])
        out[n] = rows
    return out


def test_gate_accepts_same_distribution_and_rejects_a_shift():
    rng = np.random.RandomState(0)
    batch = 1000
    p_inst = rng.uniform(0.05, 0.9, 50)
    ref = {"_meta": {}}
    r0, r1 = _rows(rng, p_inst, batch), _rows(rng, p_inst, batch)
    for n in G.SIZES:
        ref[f"x/seed0/{n}"], ref[f"x/seed1/{n}"] = r0[n], r1[n]
    same = G.gate(ref, "x", _rows(rng, p_inst, batch), batch)
    assert same["pass"], same["engine_vs_ref"]
    assert same["engine_vs_ref"]["reject_rate"] < 0.1 and same["engine_vs_ref"]["best_mismatch"] == 0
    assert "ref_seed0_vs_seed1" in same and same["max_best_mismatch"] == 3 and same["best_mismatch"] == 0
    shifted = G.gate(ref, "x", _rows(rng, p_inst, batch, shift=0.05), batch)
    assert not shifted["pass"]
    assert shifted["engine_vs_ref"]["pooled_worst_z"] > G.Z_BONF42


def test_bundled_fixture_shapes_and_reference_conventions():
    z = np.load(os.path.join(GOLDEN, "bundled_instances.npz"))
    for n in G.SIZES:
        q, v, opt = z[f"q{n}"], z[f"v{n}"], z[f"opt{n}"]
        assert q.shape == (50, n, n) and v.shape == (50, n) and opt.shape == (50,)
        assert q.dtype == np.float32 and np.allclose(q, np.swapaxes(q, 1, 2))   # dense symmetric
        assert (opt > 0).all()
        assert str(z[f"name{n}"][0]) == f"tuningH{n:03d}-100-0.in"
    # the fixture holds the reference's in-memory convention (negated on load): the committed
    # copy of Size70/tuningH070-100-1.in in the loop fixtures must be the same matrix up to its scaling
    g = np.load(os.path.join(GOLDEN, "dl_n70_p8.0.npz"))
    if "q" in g.files:
        ratio = g["q"] / z["q70"][1]
        assert np.allclose(ratio, ratio.flat[0], rtol=1e-4)


def test_reference_records_are_complete():
    path = os.path.join(GOLDEN, "equivalence_ref.json")
    ref = json.load(open(path))
    assert ref["_meta"]["batch"] == 1000 and ref["_meta"]["iterations"] == 1500
    solvers = {k.split("/")[0] for k in ref if k != "_meta"}
    assert solvers, "no reference records"
    for name in solvers:
        for n in G.SIZES:
            rows = ref.get(f"{name}/seed0/{n}")
            if rows is None:
                continue
            assert len(rows) == (ref["_meta"]["adam_per_size"] if name.endswith("_adam") else 50)
            assert all(len(r) == 8 for r in rows)
            fr = np.asarray(rows)[:, :7]
            assert (np.diff(fr, axis=1) >= 0).all() and fr.min() >= 0 and fr.max() <= 1   # nested thresholds


def test_majority_gate_needs_more_than_half_of_the_seeds_per_criterion():
    def run(rej, zok, bad):
        return {"engine_vs_ref": {"reject_rate": rej, "pooled_ok": zok}, "max_reject": 0.07, "best_mismatch": bad,
                "max_best_mismatch": 5}
    good, bad_z, bad_best = run(0.03, True, 2), run(0.03, False, 2), run(0.03, True, 9)
    assert G.majority_gate([good, good, bad_z]) and G.majority_gate([good])
    assert not G.majority_gate([good, bad_z, bad_z]) and not G.majority_gate([bad_best])
    assert G.majority_gate([good, bad_z, bad_best])          # each criterion fails under one seed only
    assert not G.majority_gate([bad_z, bad_z, good, good])   # a tie is not a majority
