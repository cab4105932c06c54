"""Instance sweeps: many BoxQP instances through one solver, optionally sharded over the ranks of a
``torch.distributed`` job (one process per GPU).  Instances are independent, so the shards need no
data-path collective; the per-instance metadata records are gathered once at the end.

Reference counterpart: the user loop of ``examples/ccvm_boxqp_*.py`` + ``metadata.py`` (a Python
``for`` over instance files); the multi-GPU part has none."""
import torch
import torch.distributed as dist

from . import parallel


def synthetic_instance(n, seed, scaling_multiplier, device="cuda", name=None, on_device=False):
    """Synthetic dense BoxQP fitted to the bundled instances: symmetric Gaussian Q with
    off-diagonal std 28.5/sqrt(N), V std 20, reference sign convention, scaled like
    ``instance.scale_coefs(solver.get_scaling_factor(Q))``.  No Gurobi optimum exists, so
    ``optimal_sol`` is left at 0 and must be filled from the best value found.

    ``on_device=False`` draws with the host generator SURVEY.md 8d specifies (seed 1000 + k, the
    instances ``bench.py`` and the tests use); ``on_device=True`` draws the same distribution with the
    engine's Philox generator directly in HBM (no host randn, no H2D copy) -- a different stream."""
    from .problem_classes.boxqp import ProblemInstance
    from . import engine
    inst = ProblemInstance(device=device, instance_type="test", name=name or f"synthetic{n:03d}-{seed}")
    inst.problem_size = n
    if on_device:
        inst.q_matrix, inst.v_vector = engine.generate_boxqp(n, 1000 + seed, device=device)
    else:
        g = torch.Generator().manual_seed(1000 + seed)
        a = torch.randn(n, n, generator=g)
        q = -((a + a.T) / 2 ** 0.5 * (28.5 / n ** 0.5)).float()
        v = -(20.0 * torch.randn(n, generator=g)).float()
        inst.q_matrix, inst.v_vector = q.to(device), v.to(device)
    inst.optimal_sol = inst.best_sol = 0.0
    inst.num_frac_values, inst.solution_vector, inst.optimality = 0, [], False
    inst.scale_coefs(engine.scaling_factor(inst.q_matrix, scaling_multiplier))
    return inst


def solve_sweep(solver, instances, post_processor=None, rank=None, world_size=None, gather=True, chunk=1,
                costs=None, **call_kwargs):
    """Solve ``instances`` (a sequence, or a ``(count, index -> instance)`` pair so that ranks only
    build their own) with ``solver``; returns the list of metadata dicts of ALL instances in order
    (on every rank when ``gather``), each extended with ``best_index``, ``rank`` and ``index``.

    ``costs`` (one number per instance, e.g. N^2) switches the placement from round-robin to
    longest-processing-time first (``parallel.lpt_owners``).

    ``chunk`` > 1 solves that many instances of this rank per launch through
    ``CCVMSolver.solve_many`` (one grid over instances x trajectory blocks, one statistics kernel,
    one device->host copy per chunk): a batch-1000 solve occupies a fraction of the GPU's SMs, so
    batching instances is what fills the machine.  Per-instance ``solve_time`` is then the chunk's
    device time apportioned by drift work."""
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    count = instances[0] if isinstance(instances, tuple) else len(instances)
    getter = instances[1] if isinstance(instances, tuple) else instances.__getitem__
    local = {}
    if costs is not None:
        # size-aware placement (cost ~ N^2 per instance): longest-processing-time first
        owners = parallel.lpt_owners(list(costs), world_size)
        mine = [k for k in range(count) if owners[k] == rank]
    else:
        mine = [k for k in range(count) if parallel.instance_owner(k, world_size) == rank]

    # one noise stream per instance, keyed by its GLOBAL index: identical results for any number of ranks
    # and any chunking, and no two ranks ever share a stream (they all start from the same generator state)
    streams = None
    if getattr(solver, "device", "cpu") == "cuda" and hasattr(solver, "noise_streams"):
        gen = torch.cuda.default_generators[torch.cuda.current_device()]
        seed, base = gen.initial_seed() & 0xFFFFFFFFFFFFFFFF, gen.get_offset()
        streams = lambda ks: [(seed, base + 4 * (k + 1)) for k in ks]  # noqa: E731
        gen.set_offset(base + 4 * (count + 1))

    def record(k, sol):
        rec = sol.get_metadata_dict()
        rec["best_index"], rec["rank"], rec["index"] = sol.best_index, rank, k
        local[k] = rec

    try:
        if chunk <= 1:
            for k in mine:
                if streams:
                    solver.noise_streams = streams([k])
                record(k, solver(instance=getter(k), post_processor=post_processor, **call_kwargs))
        elif hasattr(solver, "launch_many"):
            # software pipeline over chunks: chunk c+1 is built, planned and enqueued while chunk c runs;
            # its results are collected (one event wait, no device-wide synchronisation) afterwards
            # the first chunks are small and double up to `chunk`: the GPU starts working after a handful of
            # instances have been planned instead of after a whole chunk (what is not hidden behind kernels
            # is what limits the strong scaling of a sweep over many GPUs)
            in_flight = None
            bounds, lo, size = [], 0, max(1, chunk // 8)
            while lo < len(mine):
                bounds.append((lo, min(lo + size, len(mine))))
                lo += size
                size = min(chunk, size * 2)
            for lo, hi in bounds:
                ks = mine[lo:hi]
                if streams:
                    solver.noise_streams = streams(ks)
                handle = solver.launch_many([getter(k) for k in ks], post_processor=post_processor, **call_kwargs)
                if in_flight is not None:
                    for k, sol in zip(in_flight[0], solver.collect_many(in_flight[1])):
                        record(k, sol)
                in_flight = (ks, handle)
            if in_flight is not None:
                for k, sol in zip(in_flight[0], solver.collect_many(in_flight[1])):
                    record(k, sol)
        else:
            for lo in range(0, len(mine), chunk):
                ks = mine[lo:lo + chunk]
                sols = solver.solve_many([getter(k) for k in ks], post_processor=post_processor, **call_kwargs)
                for k, sol in zip(ks, sols):
                    record(k, sol)
    finally:
        if streams:
            solver.noise_streams = None
    if world_size == 1 or not gather:
        return [local[k] for k in sorted(local)]
    shards = [None] * world_size
    dist.all_gather_object(shards, local)
    merged = {}
    for sh in shards:
        merged.update(sh)
    return [merged[k] for k in sorted(merged)]
