"""The N>1 host logic on CPU: world_size-2 gloo processes shard a batch, and the single
all-gather merge must equal the single-process result over the union of the shards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ccvm_b200 import parallel as P
from oracle import ccvm_oracle as O


def test_shard_bounds_cover_everything():
    for total in (1, 7, 1000, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [P.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
            even = [P.shard_bounds(total, world, r, align=2) for r in range(world)]
            assert even[0][0] == 0 and sum(c for _, c in even) == total and all(s % 2 == 0 for s, c in even if c)
            for (s0, c0), (s1, _) in zip(even, even[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in even) - min(c for _, c in even) < 4   # one unit of 2, minus a ragged tail
    assert [P.instance_owner(k, 4) for k in range(6)] == [0, 1, 2, 3, 0, 1]


def test_lpt_placement_balances_heterogeneous_sweeps():
    sizes = [20, 250, 30, 240, 40, 230, 50, 220, 60, 210, 70, 200]
    costs = [n * n for n in sizes]
    owners = P.lpt_owners(costs, 4)
    load = [sum(c for c, o in zip(costs, owners) if o == r) for r in range(4)]
    rr = [sum(c for k, c in enumerate(costs) if k % 4 == r) for r in range(4)]
    assert sorted(set(owners)) == [0, 1, 2, 3]
    assert max(load) / (sum(load) / 4) < 1.20            # 6 large instances on 4 ranks: 1.16 is the optimum
    assert max(rr) / (sum(rr) / 4) > 1.4                 # round-robin is far off on this mix
    assert P.lpt_owners(costs, 4) == owners              # deterministic: every rank derives the same map
    assert P.lpt_owners([5.0, 5.0, 5.0], 2) == [0, 1, 0]  # ties: instance order, lowest rank first


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank can regenerate the whole (deterministic) problem; it only "solves" its shard
        g = torch.Generator().manual_seed(123)
        q, v = O.synthetic_boxqp(n, 3)
        x_all = torch.rand(batch, n, generator=g)
        start, count = P.shard_bounds(batch, world, rank)
        x = x_all[start:start + count]
        e = O.energy(x, q, v, 2.0)
        _, perf = O.solution_stats(e, 40.0)
        counts = [round(p * count) for p in perf.values()]
        rec = P.pack_local_result(e, x, counts, start)
        best, idx, tot, vec = P.merge_results(rec)
        out[rank] = (best.item(), int(idx.item()), tot.tolist(), vec.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_merge_equals_single_process(world):
    batch, n = 101, 12
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), batch, n, out), nprocs=world, join=True)
    g = torch.Generator().manual_seed(123)
    q, v = O.synthetic_boxqp(n, 3)
    x_all = torch.rand(batch, n, generator=g)
    e = O.energy(x_all, q, v, 2.0)
    best_ref, perf = O.solution_stats(e, 40.0)
    idx_ref = int(torch.argmin(e))
    for rank in range(world):
        best, idx, tot, vec = out[rank]
        assert best == best_ref and idx == idx_ref
        assert torch.equal(vec, x_all[idx_ref])
        assert [round(t / batch, 4) for t in tot] == list(perf.values())


def test_merge_without_process_group():
    e = torch.tensor([3.0, -2.0, 5.0])
    x = torch.arange(6.0).reshape(3, 2)
    best, idx, tot, vec = P.merge_results(P.pack_local_result(e, x, [1, 2, 3, 4, 5, 6, 7], 10))
    assert best.item() == 2.0 and idx.item() == 11 and tot.tolist() == [1, 2, 3, 4, 5, 6, 7]
    assert torch.equal(vec, x[1])


class _FakeSolution:
    def __init__(self, name, n):
        self.best_index = n % 7
        self._md = {"instance_name": name, "problem_size": n, "best_objective_value": float(n) * 1.5,
                    "solve_time": 1e-6 * n, "batch_size": 10}

    def get_metadata_dict(self):
        return dict(self._md)


class _FakeInstance:
    def __init__(self, k):
        self.name, self.problem_size = f"inst{k}", 10 + k


class _FakeSolver:
    """Stands in for a CCVMSolver on a machine without a GPU: records which instances it was given."""

    def __init__(self):
        self.seen = []

    def __call__(self, instance, post_processor=None, **kw):
        self.seen.append(instance.name)
        return _FakeSolution(instance.name, instance.problem_size)

    def solve_many(self, instances, post_processor=None, **kw):
        return [self(instance=i, post_processor=post_processor) for i in instances]


def _sweep_worker(rank, world, port, count, chunk, out):
    from ccvm_b200 import sweep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        solver = _FakeSolver()
        built = []

        def get(k):
            built.append(k)
            return _FakeInstance(k)

        md = sweep.solve_sweep(solver, (count, get), post_processor="grad-descent", chunk=chunk)
        out[rank] = (md, list(solver.seen), sorted(set(built)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("chunk", [1, 3])
def test_instance_sweep_shards_round_robin_and_gathers_in_order(chunk):
    world, count = 2, 9
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sweep_worker, args=(world, _free_port(), count, chunk, out), nprocs=world, join=True)
    for rank in range(world):
        md, seen, built = out[rank]
        # every rank ends with ALL records, in instance order, each stamped with its owner
        assert [r["index"] for r in md] == list(range(count))
        assert [r["instance_name"] for r in md] == [f"inst{k}" for k in range(count)]
        assert [r["rank"] for r in md] == [k % world for k in range(count)]
        assert all(r["best_index"] == (10 + r["index"]) % 7 for r in md)
        # ... but only built and solved its own share
        mine = [k for k in range(count) if k % world == rank]
        assert built == mine and seen == [f"inst{k}" for k in mine]
