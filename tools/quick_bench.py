"""Quick device-side timing of every solver loop at one shape (CUDA events, Philox mode)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccvm_b200 import engine as E, _native as nat  # noqa: E402


def synth(n, seed, mult, dev):
    g = torch.Generator().manual_seed(1000 + seed)
    a = torch.randn(n, n, generator=g)
    q = -((a + a.T) / 2 ** 0.5 * (28.5 / n ** 0.5))
    v = -(20.0 * torch.randn(n, generator=g))
    f = torch.sqrt(q.abs().sum()) * mult
    return (q / f).to(dev), (v / f).to(dev), float(f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--size", dest="n", type=int, default=70)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=1500)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    # under torchrun every rank solves its own shard of `--batch` trajectories (weak scaling, no
    # data-path collective); times are the max over ranks, rank 0 prints the aggregate
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peak = {m: E.microbench_fp32(m) for m in (0, 1)}
    if rank == 0:
        print(json.dumps({"fp32_peak_tflops_ffma": peak[0], "fp32_peak_tflops_ffma2": peak[1], "n_gpus": world}))
    hp = dict(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=False)
    cases = {
        "dl": (nat.SOLVER_DL, nat.ALG_ORIGINAL, 0.2, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, feedback_scale=100.0, g=0.05)),
        "dl_adam": (nat.SOLVER_DL, nat.ALG_ADAM, 0.2, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, g=0.05, hyperparameters=hp)),
        "mf": (nat.SOLVER_MF, nat.ALG_ORIGINAL, 0.05, dict(s=20.0, pump=0.0, dt=0.0025, j=5.0, feedback_scale=4000.0, g=0.01)),
        "mf_adam": (nat.SOLVER_MF, nat.ALG_ADAM, 0.05, dict(s=20.0, pump=0.0, dt=0.0025, j=5.0, feedback_scale=4000.0, g=0.01, hyperparameters=hp)),
        "langevin": (nat.SOLVER_LANGEVIN, nat.ALG_ORIGINAL, 0.05, dict(s=0.5, dt=0.002, sigma=0.5, feedback_scale=1.0)),
        "langevin_adam": (nat.SOLVER_LANGEVIN, nat.ALG_ADAM, 0.05, dict(s=0.5, dt=0.002, sigma=0.5, feedback_scale=1.0, hyperparameters=hp)),
        "pumped_langevin": (nat.SOLVER_PUMPED_LANGEVIN, nat.ALG_ORIGINAL, 0.05, dict(s=0.5, pump=2.0, dt=0.002, sigma=0.5, feedback_scale=1.0)),
        "pumped_langevin_adam": (nat.SOLVER_PUMPED_LANGEVIN, nat.ALG_ADAM, 0.05, dict(s=0.5, pump=2.0, dt=0.002, sigma=0.5, feedback_scale=1.0, hyperparameters=hp)),
    }
    for name, (sid, alg, mult, kw) in cases.items():
        if args.only and name not in args.only.split(","):
            continue
        q, v, f = synth(args.n, 0, mult, dev)
        # warm-up: at least 3 launches AND 100 ms (the first launches of a process also pay module
        # loading and the clock ramp; three short solves do not cover that)
        import time
        t0, w = time.perf_counter(), 0
        while w < 3 or time.perf_counter() - t0 < 0.1:
            E.solve(sid, alg, q, v, args.batch, args.iters, seed=1, offset=w, **kw)
            torch.cuda.synchronize()
            w += 1
        # per-launch events, median: a sporadic host-side stall (allocator growth, module loading)
        # in one launch must not leak into a kernel number
        per = []
        for r in range(args.reps):
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            outs, _ = E.solve(sid, alg, q, v, args.batch, args.iters, seed=1, offset=10 + r,
                              traj_base=rank * args.batch, **kw)
            e1.record()
            torch.cuda.synchronize()
            per.append(e0.elapsed_time(e1))
        per.sort()
        ms = per[len(per) // 2]
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        if rank != 0:
            continue
        steps = world * args.batch * args.iters / (ms * 1e-3)
        m = 2 if sid == nat.SOLVER_DL else 1
        tflops = steps * 2 * m * args.n ** 2 / 1e12
        finite = bool(torch.isfinite(outs[0]).all())
        print(json.dumps({"solver": name, "n": args.n, "batch": world * args.batch, "iters": args.iters, "ms": round(ms, 4),
                          "traj_steps_per_s": steps, "drift_tflops": round(tflops, 3),
                          "frac_of_ffma2_peak": round(tflops / peak[1] / world, 4), "finite": finite,
                          "n_gpus": world, "batch_per_gpu": args.batch}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
