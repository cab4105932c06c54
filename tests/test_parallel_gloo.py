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
    assert [P.instance_owner(k, 4) for k in range(6)] == [0, 1, 2, 3, 0, 1]


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
