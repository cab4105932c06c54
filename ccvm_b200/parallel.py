"""Multi-GPU layer: one process per GPU, trajectories (or instances) sharded across ranks with NO
data-path collective; a single collective round at the end merges the per-rank results.

The reference has no multi-device code; the contract here is "same result as one
GPU over the union of the shards": noise is keyed by the GLOBAL trajectory index
(``traj_base``), the best objective is the min over ranks, the winner's solution vector comes
from its owner, and success counters add up.

Collective: one ``all_gather`` of a packed [best_energy, winner index (2 words), counts(7), x(N)]
record per rank (N+10 words; latency-bound, NVLink bandwidth irrelevant), after which every rank
selects the winner locally -- no second round, no host synchronisation.
"""
import torch
import torch.distributed as dist

N_COUNTS = 7
HEADER = 3 + N_COUNTS   # CCVM_RECORD_HEADER: energy, index low, index high, 7 counters


def shard_bounds(total, world_size, rank, align=1):
    """Contiguous split of ``total`` items: (start, count) of ``rank``; remainders go to the
    lowest ranks, so shard sizes differ by at most ``align``.  ``align=2`` keeps every shard start
    even -- what trajectory shards need, because the engine's noise streams belong to trajectory
    pairs (``traj_base`` must be even)."""
    total, align = int(total), int(align)
    units = (total + align - 1) // align
    base, rem = divmod(units, int(world_size))
    ucount = base + (1 if rank < rem else 0)
    ustart = rank * base + min(rank, rem)
    start = min(ustart * align, total)
    return start, min((ustart + ucount) * align, total) - start


def instance_owner(index, world_size):
    """Round-robin owner of instance ``index`` in a sweep."""
    return index % world_size


def lpt_owners(costs, world_size):
    """Longest-processing-time placement of a sweep: instances in decreasing order of cost (the drift
    work batch x iterations x N^2 is a good proxy), each to the least loaded rank so far; ties keep
    the instance order and go to the lowest rank, so every rank computes the same map.  Returns the
    owner of every instance.  Round-robin (``instance_owner``) is within a few percent of it for
    thousands of instances; for a few dozen of very different sizes it can be off by 2x."""
    order = sorted(range(len(costs)), key=lambda k: (-float(costs[k]), k))
    load = [0.0] * int(world_size)
    owners = [0] * len(costs)
    for k in order:
        r = min(range(int(world_size)), key=lambda i: (load[i], i))
        owners[k] = r
        load[r] += float(costs[k])
    return owners


def pack_local_result(energy, problem_variables, counts, traj_base):
    """Per-rank record from the local objective values (B_local,), the local solution matrix
    (B_local, N) and the 7 local success counters.  Everything stays on the device.  Layout (32-bit
    words): [min energy (f32), global winner index low / high half (i32 bit patterns), 7 counters
    (i32), winner's vector (f32 x N)] -- integers travel as integers, so nothing is rounded."""
    e_min, idx = torch.min(energy, dim=0)  # torch.min propagates NaN, like the merge below
    rec = torch.empty(HEADER + problem_variables.shape[1], dtype=torch.float32, device=energy.device)
    ints = rec.view(torch.int32)
    rec[0] = e_min
    g = idx.to(torch.int64) + int(traj_base)
    ints[1] = (((g & 0xFFFFFFFF) + (1 << 31)) % (1 << 32) - (1 << 31)).to(torch.int32)   # low half, two's complement
    ints[2] = (g >> 32).to(torch.int32)
    ints[3:HEADER] = torch.as_tensor(counts, dtype=torch.int32, device=energy.device)
    rec[HEADER:] = problem_variables[idx]
    return rec


def pack_from_stats(stats, problem_variables, traj_base):
    """Device-side ``pack_local_result``: ONE kernel builds the record from the 9-word block that
    ``ccvm_solution_stats`` (or a fused solve) wrote (best, argmin, 7 counters) and the local solution matrix."""
    from . import _native as nat
    n = problem_variables.shape[1]
    rec = torch.empty(HEADER + n, dtype=torch.float32, device=problem_variables.device)
    with torch.cuda.device(rec.device):
        nat.check(nat.load().ccvm_pack_record(stats.data_ptr(), problem_variables.data_ptr(), int(n), int(traj_base),
                                              rec.data_ptr(), nat.current_stream_ptr(rec.device)))
    return rec


def unpack_record(rec, slot0_is_objective=False):
    """(best objective value = max(-E) tensor, global winner index (int64 tensor), counts (7,) int32,
    vector) of a record; slot 0 holds the energy (per-rank records) or already the objective (merged)."""
    ints = rec.view(torch.int32)
    idx = (ints[1].to(torch.int64) & 0xFFFFFFFF) | (ints[2].to(torch.int64) << 32)
    return (rec[0] if slot0_is_objective else -rec[0]), idx, ints[3:HEADER], rec[HEADER:]


def merge_results(record, group=None):
    """All-gather the per-rank records and reduce them identically on every rank.

    Returns (best_objective_value tensor = max(-E) over all ranks, global trajectory index of
    the winner (int64), summed counts (7,) int32, winner's solution vector (N,)).  Ties go to the
    lowest rank; a NaN objective wins (torch.max(-E) propagates NaN, solution.py:82)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        gathered = record.unsqueeze(0)
    else:
        world = dist.get_world_size(group)
        flat = torch.empty(world * record.numel(), dtype=record.dtype, device=record.device)
        dist.all_gather_into_tensor(flat, record.contiguous(), group=group)
        gathered = flat.view(world, record.numel())
    if gathered.is_cuda:
        # one kernel instead of half a dozen small torch launches per step
        from . import _native as nat
        n = record.numel() - HEADER
        out = torch.empty_like(record)
        with torch.cuda.device(record.device):
            nat.check(nat.load().ccvm_merge_records(gathered.data_ptr(), int(gathered.shape[0]), int(n), out.data_ptr(),
                                                    nat.current_stream_ptr(record.device)))
        return unpack_record(out, slot0_is_objective=True)   # the kernel stores max(-E) in slot 0
    e = gathered[:, 0]
    nan = torch.isnan(e)
    owner = torch.argmax(nan.to(torch.int8)) if bool(nan.any()) else torch.argmin(e)   # first NaN rank, else min
    merged = gathered[owner].clone()
    merged.view(torch.int32)[3:HEADER] = gathered.view(torch.int32)[:, 3:HEADER].sum(dim=0)
    return unpack_record(merged)
