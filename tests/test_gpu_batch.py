"""GPU tests of the many-instances-per-launch path (``ccvm_solve_batch``, ``ccvm_solution_stats_batch``,
``CCVMSolver.solve_many``): its oracle is the single-instance path over the same generator state --
results must be bit-identical ("same result as ... over the union")."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from ccvm_b200 import engine as E, sweep  # noqa: E402
from ccvm_b200.solvers import DLSolver, MFSolver, LangevinSolver, PumpedLangevinSolver, AdamParameters  # noqa: E402

KEYS = {
    "dl": (DLSolver, dict(pump=8.0, dt=0.001, noise_ratio=10, feedback_scale=100), None),
    "mf": (MFSolver, dict(pump=0.0, feedback_scale=4000, j=5.0, S=20.0, dt=0.0025), "grad-descent"),
    "lv": (LangevinSolver, dict(dt=0.002, S=0.5, sigma=0.5, feedback_scale=1.0), "adam"),
    "plv": (PumpedLangevinSolver, dict(pump=2.0, dt=0.002, S=0.5, sigma=0.5, feedback_scale=1.0), "grad-descent"),
}
# sizes straddle every CTA-size bucket of both batched kernels (TMEM: n <= 128, hybrid: n <= 256) and
# include one instance beyond them (streamed Q, solved by its own launch)
SIZES = (5, 20, 33, 70, 128, 140, 64, 250, 20, 300)


def _solver(name, batch, iters):
    cls, key, pp = KEYS[name]
    solver = cls(device="cuda", batch_size=batch)
    solver.parameter_key = {n: dict(key, iterations=iters) for n in set(SIZES)}
    return solver, pp


def _instances(solver):
    out = []
    for k, n in enumerate(SIZES):
        inst = sweep.synthetic_instance(n, k, solver._scaling_multiplier)
        inst.optimal_sol = 1.0 + k
        out.append(inst)
    return out


@pytest.mark.parametrize("name", ["dl", "mf", "lv", "plv"])
@pytest.mark.parametrize("adam", [False, True])
def test_solve_many_equals_sequential_calls(name, adam):
    batch, iters = 137, 60
    solver, pp = _solver(name, batch, iters)
    insts = _instances(solver)
    alg = AdamParameters(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=True) if adam else None
    torch.manual_seed(7)
    seq = [solver(instance=i, post_processor=pp, algorithm_parameters=alg) for i in insts]
    torch.manual_seed(7)
    many = solver.solve_many(insts, post_processor=pp, algorithm_parameters=alg)
    assert len(many) == len(seq)
    for a, b in zip(seq, many):
        assert a.instance_name == b.instance_name and a.problem_size == b.problem_size
        assert torch.equal(a.objective_values, b.objective_values)
        for key in a.variables:
            assert torch.equal(a.variables[key], b.variables[key]), key
        assert a.best_objective_value == b.best_objective_value
        assert a.best_index == b.best_index
        assert a.solution_performance == b.solution_performance
        assert b.solve_time > 0 and (b.pp_time > 0) == bool(pp)
        assert set(a.get_metadata_dict()) == set(b.get_metadata_dict())


def test_solve_many_rejects_unsupported_modes():
    solver, pp = _solver("lv", 16, 10)
    insts = _instances(solver)[:2]
    with pytest.raises(ValueError):
        solver.solve_many(insts, evolution_step_size=2)
    solver.noise_source = torch.zeros(10, 1, insts[0].problem_size, 16, device="cuda")
    with pytest.raises(ValueError):
        solver.solve_many(insts[:1])
    solver.noise_source = None
    assert solver._deferred is None          # a failed gather must not leave the solver in plan-only mode
    assert solver.solve_many([]) == []


def test_solution_stats_batch_equals_single():
    g = torch.Generator().manual_seed(3)
    sizes = [1, 7, 1000, 1025, 64]
    en = [(-100.0 - 5.0 * torch.rand(b, generator=g)).cuda() for b in sizes]
    en[1][3] = float("nan")
    opts = [104.9, 103.0, 104.99, 102.0, 104.0]
    offs = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    got = E.solution_stats_batch(torch.cat(en), offs, opts)
    for e, o, (best, arg, counts) in zip(en, opts, got):
        b1, a1, c1 = E.solution_stats(e, o)
        assert (best == b1) or (best != best and b1 != b1)
        assert arg == a1 and counts == c1


def test_sweep_chunked_equals_unchunked():
    solver, pp = _solver("plv", 100, 40)
    insts = _instances(solver)
    torch.manual_seed(1)
    a = sweep.solve_sweep(solver, insts, post_processor=pp, rank=0, world_size=1)
    torch.manual_seed(1)
    b = sweep.solve_sweep(solver, insts, post_processor=pp, rank=0, world_size=1, chunk=3)
    strip = lambda r: {k: v for k, v in r.items() if k not in ("solve_time", "pp_time")}  # noqa: E731
    assert [strip(r) for r in a] == [strip(r) for r in b]


def test_sweep_results_do_not_depend_on_rank_count_or_chunking():
    """Every instance's noise stream is keyed by its GLOBAL index in the sweep (sweep.solve_sweep), so the shards of
    a 2-rank or 3-rank sweep (emulated here rank by rank, no process group) reproduce the 1-rank results exactly,
    whatever the chunk size."""
    solver, pp = _solver("lv", 64, 30)
    insts = _instances(solver)
    torch.manual_seed(5)
    one = sweep.solve_sweep(solver, insts, post_processor=pp, rank=0, world_size=1, chunk=4)
    strip = lambda r: {k: v for k, v in r.items() if k not in ("solve_time", "pp_time", "rank")}  # noqa: E731
    for world, chunk in ((2, 1), (3, 8)):
        merged = {}
        for rank in range(world):
            torch.manual_seed(5)          # every rank starts from the same generator state, as under torchrun
            for rec in sweep.solve_sweep(solver, insts, post_processor=pp, rank=rank, world_size=world, chunk=chunk,
                                         gather=False):
                merged[rec["index"]] = rec
        assert sorted(merged) == list(range(len(insts)))
        assert [strip(merged[k]) for k in sorted(merged)] == [strip(r) for r in one]
