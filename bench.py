#!/usr/bin/env python
"""Headline benchmark: SDE trajectory-steps/second on the configuration BASELINE.json's metric is
quoted on -- DLSolver `_solve_adam` + ADAM post-processor + BoxQP energy + solution statistics on
a synthetic BoxQP instance, N=70, batch 4096 per GPU, 1500 iterations (BASELINE.json configs[2];
SURVEY.md 8d "Config 3").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation

One "step" = one whole solve of the batch (all 1500 iterations in one persistent kernel) followed
by the fused epilogue and the statistics kernel.  Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N, BATCH, ITERS = 70, 4096, 1500
DL = dict(pump=8.0, dt=0.001, noise_ratio=10.0, feedback_scale=100.0, g=0.05)
HP = dict(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=False)
METRIC = "SDE trajectory-steps/sec (DL-adam + adam post-processor, N=70, B=4096 per GPU)"
UNIT = "trajectory-steps/s"
CPU_SAMPLE_ITERS = 50


def synthetic_instance(n, seed):
    """SURVEY.md 8d generator, reference sign convention (Q, V negated), DL scaling (0.2)."""
    g = torch.Generator().manual_seed(1000 + seed)
    a = torch.randn(n, n, generator=g)
    q = -((a + a.T) / np.sqrt(2.0) * (28.5 / np.sqrt(n))).float()
    v = -(20.0 * torch.randn(n, generator=g)).float()
    f = torch.sqrt(torch.sum(torch.abs(q))) * 0.2
    return q / f, v / f, float(f)


def workload_config(n_gpus):
    """The workload, identical for both arms (arm-specific notes go to the line's `notes`)."""
    return {
        "workload": "configs[2]: DLSolver._solve_adam (pump 8, dt 0.001, noise_ratio 10, g 0.05; Adam alpha 1e-3, "
                    "beta 0.9/0.999, add_assign False) + adam post-processor + compute_energy + solution stats",
        "n": N, "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "iterations": ITERS,
        "instance": "synthetic dense BoxQP, seed 1000 (SURVEY 8d generator)",
    }


GPU_NOTES = {
    "rng": "in-kernel: Philox4x32-10-seeded xoshiro128+ stream per (trajectory pair, variable), Box-Muller",
    "l2": "L2 flushed (256 MiB write) before every timed step, outside the per-step event pair",
    "parallelism": "batch sharded over the GPUs, no data-path collective; one all_gather of N+10 words per step",
    "launches": "one fused kernel per step (schedules + loop + change of variables + post-processor + energy + "
                "statistics); N > 1 adds pack_record + merge_records",
}


def _clean(obj):
    """Strict JSON: non-finite floats (an unsolved size has TTS = inf) become strings."""
    if isinstance(obj, float) and not np.isfinite(obj):
        return str(obj)
    if isinstance(obj, dict):
        return {k: _clean(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_clean(v) for v in obj]
    return obj


# ------------------------------------------------------------------------ CPU reference arm
def _load_reference():
    """The unmodified reference package if a copy travelled with the repo (baseline/_ref), else None."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_dir, "ccvm_simulators")):
        sys.path.insert(0, ref_dir)
        try:
            from ccvm_simulators.solvers import DLSolver  # noqa: F401
            from ccvm_simulators.post_processor.adam import PostProcessorAdam  # noqa: F401
            return True
        except Exception:
            sys.path.remove(ref_dir)
    return False


def cpu_step(kind, q, v, sb, iters):
    """One bounded sample of the workload on the host cores: `iters` iterations of the DL-adam loop
    on the full batch + adam post-processor + energy.  kind == 'reference' runs the reference's own
    functions (function-level, because its DLSolver.__call__ + AdamParameters raises TypeError,
    SURVEY.md 8c(4)); kind == 'port' runs the oracle restatement."""
    s_val = float(np.sqrt(DL["pump"] - 1))
    if kind == "reference":
        import contextlib
        import io
        from ccvm_simulators.solvers import DLSolver
        from ccvm_simulators.post_processor.adam import PostProcessorAdam
        sol = DLSolver(device="cpu", batch_size=BATCH, S=s_val)
        sol.q_matrix, sol.v_vector, sol.solution_bounds = q, v, (0.0, 1.0)
        c, _ = sol._solve_adam(N, BATCH, "cpu", s_val, DL["pump"], DL["dt"], iters, DL["noise_ratio"], True, DL["g"],
                               None, None, dict(HP))
        x = sol.change_variables(c, 0.0, 1.0, s_val)
        with contextlib.redirect_stderr(io.StringIO()):
            pv = PostProcessorAdam().postprocess(x, q, v)
        e = 0.5 * torch.einsum("bi, ij, bj -> b", pv, q, pv) * sb + torch.einsum("bi, i -> b", pv, v) * sb
    else:
        from oracle import ccvm_oracle as O
        c, _ = O.dl_solve_adam(q, v, BATCH, iters, DL["pump"], DL["dt"], DL["noise_ratio"], O.NoiseSource(N, BATCH),
                               dict(HP), True, DL["g"], 1.0)
        pv, e = O.epilogue("mf", c, q, v, sb, s_val, post_processor="adam")
    return float((-e).max())


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    kind = "reference" if _load_reference() else "port"
    q, v, sb = synthetic_instance(N, 0)
    for _ in range(max(args.warmup, 1) if args.warmup else 0):
        cpu_step(kind, q, v, sb, CPU_SAMPLE_ITERS)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(kind, q, v, sb, CPU_SAMPLE_ITERS)
    dt = time.perf_counter() - t0
    value = BATCH * CPU_SAMPLE_ITERS * args.steps / dt
    sample = (f"{CPU_SAMPLE_ITERS} of {ITERS} iterations of the same B={BATCH}, N={N} DL-adam loop + adam "
              f"post-processor + energy per step (per-iteration cost is constant in T)")
    cfg = workload_config(args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "notes": {"parallelism": "single host process, torch intra-op threads", "rng": "torch CPU generator",
                  "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def roofline_record(E, nat, q, v, achieved_tf, fp32_peak_tf, traffic, kernel_ms, loop_ms, tail_ms, flops_per_launch):
    """Roofline of the dominant kernel.  The library serves this shape with the small-n tensor-core kernel
    (csrc/sde_kernel_mma.cuh): the drift flops run on tcgen05 as three FP16 products per FP32-grade product, so the
    roofline they are measured against is the tensor pipe: MEASURED_PEAKS.json's dense bf16 rate / 3.  The kernel is
    NOT bound by it -- the contraction is ~1/3 of an iteration, the rest is the elementwise SDE step (noise, Adam) on
    the SIMT pipes -- so the line also carries the fraction of the FP32 SIMT FMA peak, the roofline of the tiled
    kernel this one replaced (CCVM_MMA=0) and the number round 1 was judged on."""
    plan = E.plan_solve(nat.SOLVER_DL, nat.ALG_ADAM, q, v, BATCH, ITERS, s=1.0, pump=DL["pump"], dt=DL["dt"],
                        noise_ratio=DL["noise_ratio"], g=DL["g"], hyperparameters=HP, seed=1, offset=0)
    info = E.query_launch(plan.desc)
    common = {"unit": "TFLOP/s", "achieved": achieved_tf, "traffic": traffic, "kernel_ms": kernel_ms,
              "loop_ms_in_kernel": loop_ms, "tail_ms_in_kernel": tail_ms, "algorithmic_flops_per_launch": flops_per_launch,
              "launch": info, "fp32_simt_peak": fp32_peak_tf, "frac_of_fp32_simt_peak": achieved_tf / fp32_peak_tf,
              "fp32_simt_peak_source": "measured in-process: register-only FFMA2 probe (ccvm_microbench_fp32)"}
    if info["threads"] == 288:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            peaks = {}
        bf16 = float(peaks.get("bf16_tflops", 0.0)) or 2250.0
        src = ("MEASURED_PEAKS.json bf16_tflops (burst: the kernel is timed alone) / 3" if peaks.get("bf16_tflops")
               else "nominal 2250 TFLOP/s dense 16-bit (B200_PROFILING.md fallback) / 3")
        return {"bound": "tensor", "peak": bf16 / 3.0, "frac": achieved_tf / (bf16 / 3.0), "peak_source": src,
                "kernel": "ccvm::sde_mma_kernel<DL, adam, items per lane 4> (fused: schedules + loop + tail; drift "
                          "contraction on tcgen05 kind::f16, 3 FP16-split products, Qs^T resident in TMEM)",
                "not_the_limit": "elementwise SDE step (noise + Adam, SIMT issue / MUFU) and the per-iteration "
                                 "mbarrier -> MMA -> commit -> tcgen05.ld handshake bound this kernel, see DESIGN.md 4",
                **common}
    return {"bound": "fp32", "peak": fp32_peak_tf, "frac": achieved_tf / fp32_peak_tf,
            "peak_source": common["fp32_simt_peak_source"] + "; MEASURED_PEAKS.json has no FP32 SIMT figure",
            "kernel": "ccvm::sde_tmem_kernel<DL, adam, TMEM, PIPE, 18> (fused: schedules + loop + tail)", **common}


def oracle_check(E, nat, q_host, v_host, q, v, sb, s_val, traj_base):
    """Best-of-batch and per-trajectory objective of the PRODUCTION kernel (same launch geometry as the timed
    steps: full batch) against the CPU oracle on a 64-trajectory slice, the oracle replaying the dumped noise."""
    from oracle import ccvm_oracle as O
    iters, nb = 200, 64
    plan = E.plan_solve(nat.SOLVER_DL, nat.ALG_ADAM, q, v, BATCH, iters, s=1.0, pump=DL["pump"], dt=DL["dt"],
                        noise_ratio=DL["noise_ratio"], g=DL["g"], hyperparameters=HP, seed=77, offset=5,
                        traj_base=traj_base)
    epi = E.plan_epilogue(BATCH, N, q.device, map1=(0.5 / s_val, 0.5), post_processor="adam", scaled_by=sb)
    E.solve_fused(plan, epi, 0.0)
    noise = E.dump_noise(nat.SOLVER_DL, N, nb, iters, 77, 5, traj_base=traj_base, launch_batch=BATCH).cpu()
    c_ref, _ = O.dl_solve_adam(q_host, v_host, nb, iters, DL["pump"], DL["dt"], DL["noise_ratio"],
                               O.NoiseSource(N, nb, replay=noise), dict(HP), True, DL["g"], 1.0)
    _, e_ref = O.epilogue("mf", c_ref, q_host, v_host, sb, s_val, post_processor="adam")
    e_gpu = epi.energy[:nb].cpu()
    rel = ((e_gpu - e_ref).abs() / e_ref.abs().clamp_min(1e-6)).max().item()
    return {"trajectories": nb, "iterations": iters, "max_rel_objective_err": rel, "tolerance": 2e-3,
            "best_gpu": float((-e_gpu).max()), "best_oracle": float((-e_ref).max()), "ok": bool(rel <= 2e-3)}


def strong_record(E, nat, P, rank, world, dev, q, v, sb, s_val):
    """STRONG scaling of the headline workload: ONE batch of 4096 trajectories split over the GPUs (even shard
    starts), same fused step + cross-rank merge, device-timed, max over ranks.  With 512 trajectories per GPU at
    N = 8 an SM owns 3-4 trajectories: the loop is latency-bound there, this record says by how much."""
    import torch.distributed as dist
    start, count = P.shard_bounds(BATCH, world, rank, align=2)
    stream = torch.cuda.current_stream(dev)

    def step(k):
        plan = E.plan_solve(nat.SOLVER_DL, nat.ALG_ADAM, q, v, count, ITERS, s=1.0, pump=DL["pump"], dt=DL["dt"],
                            noise_ratio=DL["noise_ratio"], g=DL["g"], hyperparameters=HP, seed=4321, offset=k,
                            traj_base=start)
        epi = E.plan_epilogue(count, N, dev, map1=(0.5 / s_val, 0.5), post_processor="adam", scaled_by=sb)
        res = E.solve_fused(plan, epi, 0.0)
        if world > 1:
            return P.merge_results(P.pack_from_stats(res, epi.pv, start))
        return res

    for k in range(3):
        step(k)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    steps = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(stream)
    for k in range(steps):
        step(10 + k)
    ev[1].record(stream)
    torch.cuda.synchronize(dev)
    ms = ev[0].elapsed_time(ev[1]) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return {"workload": "the headline step on ONE batch of 4096 trajectories sharded over the GPUs", "scaling": "strong",
            "n_gpus": world, "batch_per_gpu": count, "ms_per_step": ms, "traj_steps_per_s": BATCH * ITERS / (ms * 1e-3)}


SWEEP_SIZES = list(range(20, 251, 10))
SWEEP_COUNT, SWEEP_BATCH, SWEEP_ITERS, SWEEP_CHUNK = 1024, 1000, 1500, 64


def sweep_record(rank, world, dev):
    """BASELINE configs[4] (scaled to 1024 instances so that it fits the bench's time budget): synthetic
    BoxQP instances with N uniform on {20, 30, ..., 250}, B = 1000, T = 1500, Langevin + grad-descent,
    dealt over the ranks (longest-processing-time placement), chunked fused launches (solve_many).
    STRONG scaling: the instance set is fixed, `wall_s` includes building this rank's instances on the
    device, planning, launching and the final gather of the metadata records."""
    import torch.distributed as dist
    from ccvm_b200 import sweep
    from ccvm_b200.solvers import LangevinSolver
    solver = LangevinSolver(device="cuda", batch_size=SWEEP_BATCH)
    solver.parameter_key = {n: dict(dt=0.002, S=0.5, sigma=0.5, feedback_scale=1.0, iterations=SWEEP_ITERS)
                            for n in SWEEP_SIZES}
    draw = np.random.RandomState(0).choice(len(SWEEP_SIZES), SWEEP_COUNT)
    specs = [(SWEEP_SIZES[int(i)], k) for k, i in enumerate(draw)]
    costs = [n * n + 1000 for n, _ in specs]   # measured: solve time ~ a + b n^2 with a / b ~ 1000 (profiles/r2g_quick_bench_sizes_20_250.jsonl)

    def get(i):
        n, k = specs[i]
        return sweep.synthetic_instance(n, k, solver._scaling_multiplier, on_device=True)

    # untimed warm-up: a small sweep of the same shape (every size, the same ramp of chunk sizes up to full
    # chunks, two chunks in flight).  One instance per size is not enough: the FIRST sweep of a process then still
    # grows the library's stream-ordered pool to the working set of two full chunks, which costs 0.5-1 s of
    # blocking driver calls inside the enqueue path (profiles/r2z_sweep_first_vs_repeat.jsonl: wall 2.1-2.5 s for
    # the first sweep, 1.81 s for every later one)
    warm_specs = [(SWEEP_SIZES[k % len(SWEEP_SIZES)], 10_000 + k) for k in range(3 * SWEEP_CHUNK)]
    sweep.solve_sweep(solver, (len(warm_specs), lambda i: sweep.synthetic_instance(
        warm_specs[i][0], warm_specs[i][1], solver._scaling_multiplier, on_device=True)),
        post_processor="grad-descent", chunk=SWEEP_CHUNK, costs=[n * n + 1000 for n, _ in warm_specs],
        rank=0, world_size=1, gather=False)
    torch.cuda.synchronize(dev)
    solver.host_seconds = {"launch": 0.0, "collect": 0.0}
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    md = sweep.solve_sweep(solver, (SWEEP_COUNT, get), post_processor="grad-descent", chunk=SWEEP_CHUNK, costs=costs)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = t.item()
    steps = float(SWEEP_COUNT) * SWEEP_BATCH * SWEEP_ITERS
    return {"workload": "configs[4] mix: 1024 synthetic BoxQP instances, N uniform on {20..250 step 10}, B=1000, T=1500, "
                        "Langevin + grad-descent", "scaling": "strong", "instances": len(md), "n_gpus": world,
            "chunk": SWEEP_CHUNK, "wall_s": wall, "traj_steps_per_s": steps / wall,
            "drift_tflops": sum(2.0 * n * n for n, _ in specs) * SWEEP_BATCH * SWEEP_ITERS / wall / 1e12,
            "ms_per_instance": wall / SWEEP_COUNT * 1e3,
            "host_seconds_rank0": {k: round(v, 4) for k, v in solver.host_seconds.items()},
            "finite": bool(all(np.isfinite(r["best_objective_value"]) for r in md))}


def config4_record(E, nat, rank, world, dev):
    """BASELINE configs[3]: synthetic dense BoxQP N = 1024, batch 8192 per GPU through the tcgen05 3xTF32 path
    (weak scaling; max over ranks), as logical drift TFLOP/s and as a fraction of the measured 3xTF32 ceiling
    (a third of the dense TF32 rate of back-to-back tcgen05.mma, ccvm_microbench_tf32)."""
    import torch.distributed as dist
    n, b, t = 1024, 8192, 100
    q, v, _ = synthetic_instance(n, 1)
    qd, vd = q.to(dev), v.to(dev)
    ceiling = E.microbench_tf32(2) / 3.0
    out = {"n": n, "batch_per_gpu": b, "iterations": t, "n_gpus": world, "ceiling_3xtf32_tflops": ceiling, "loops": {}}
    cases = {"dl": (nat.SOLVER_DL, 2, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, feedback_scale=100.0, g=0.05)),
             "langevin": (nat.SOLVER_LANGEVIN, 1, dict(s=0.5, dt=0.002, sigma=0.5, feedback_scale=1.0))}
    for name, (sid, m, kw) in cases.items():
        for w in range(2):
            E.solve(sid, nat.ALG_ORIGINAL, qd, vd, b, t, seed=3, offset=w, traj_base=rank * b, **kw)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        per = []
        for r in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            outs, _ = E.solve(sid, nat.ALG_ORIGINAL, qd, vd, b, t, seed=3, offset=8 + r, traj_base=rank * b, **kw)
            e1.record()
            torch.cuda.synchronize(dev)
            per.append(e0.elapsed_time(e1))
        ms = sorted(per)[1]
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = tt.item()
        tf = world * 2.0 * m * n * n * b * t / (ms * 1e-3) / 1e12
        out["loops"][name] = {"ms": ms, "logical_tflops": tf, "frac_of_ceiling": tf / ceiling / world,
                              "finite": bool(torch.isfinite(outs[0]).all())}
    return out


def gpu_eager_record(q, v, sb):
    """the UNMODIFIED reference with device='cuda' (eager torch on the same B200), the whole
    workload once (B = 4096, T = 1500), with explicit synchronisation around it (the reference itself stops
    its clock without one, dl_solver.py:851,933)."""
    if not _load_reference():
        return {"unavailable": "baseline/_ref not present"}
    import contextlib
    import io
    from ccvm_simulators.solvers import DLSolver
    from ccvm_simulators.post_processor.adam import PostProcessorAdam
    s_val = float(np.sqrt(DL["pump"] - 1))

    def run(iters):
        sol = DLSolver(device="cuda", batch_size=BATCH, S=s_val)
        sol.q_matrix, sol.v_vector, sol.solution_bounds = q, v, (0.0, 1.0)
        c, _ = sol._solve_adam(N, BATCH, "cuda", s_val, DL["pump"], DL["dt"], iters, DL["noise_ratio"], True, DL["g"],
                               None, None, dict(HP))
        x = sol.change_variables(c, 0.0, 1.0, s_val)
        with contextlib.redirect_stderr(io.StringIO()):
            pv = PostProcessorAdam().postprocess(x, q, v, device="cuda")
        e = 0.5 * torch.einsum("bi, ij, bj -> b", pv, q, pv) * sb + torch.einsum("bi, i -> b", pv, v) * sb
        return float((-e).max())

    run(50)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    best = run(ITERS)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": BATCH * ITERS / dt, "unit": UNIT, "seconds": dt, "kind": "reference, device='cuda' (eager torch)",
            "best_objective": best}


def tts_record():
    """TTS half of the metric on a bundled Size70 subset: DLSolver (B = 1000, T = 1500, the example's parameter
    key), this engine on the GPU vs the unmodified reference on the host CPU, the reference's own definition
    (mean per-run solve time x bootstrapped R99 of the <= 0.1 %-gap success fraction; ccvm_b200/tts.py)."""
    if not _load_reference():
        return {"unavailable": "baseline/_ref not present"}
    import contextlib
    import io
    from ccvm_b200 import tts
    from ccvm_b200.solvers import DLSolver as EngineDL
    from tools.equivalence_gpu import load_bundled
    import ccvm_simulators.solvers as RS
    from ccvm_simulators.problem_classes.boxqp import ProblemInstance as RefInstance
    n, count = 70, 4
    key = dict(pump=8.0, feedback_scale=100, dt=0.001, iterations=1500, noise_ratio=10)
    torch.manual_seed(0)
    eng = EngineDL(device="cuda", batch_size=1000)
    eng.parameter_key = {n: dict(key)}
    insts = load_bundled()[n][:count]
    for inst in insts:
        inst.scale_coefs(eng.get_scaling_factor(inst.q_matrix))
    eng(instance=insts[0])
    recs_e = [eng(instance=i).get_metadata_dict() for i in insts]
    z = np.load(os.path.join(ROOT, "tests", "golden", "bundled_instances.npz"))
    ref = RS.DLSolver(device="cpu", batch_size=1000)
    ref.parameter_key = {n: dict(key)}
    torch.set_num_threads(os.cpu_count())
    recs_r = []
    for k in range(count):
        inst = RefInstance(instance_type="tuning", device="cpu", name=str(z[f"name{n}"][k])[:-3])
        inst.problem_size = n
        inst.q_matrix = torch.from_numpy(z[f"q{n}"][k].copy())
        inst.v_vector = torch.from_numpy(z[f"v{n}"][k].copy())
        inst.optimal_sol = inst.best_sol = float(z[f"opt{n}"][k])
        inst.num_frac_values, inst.solution_vector, inst.optimality = 0, [], True
        inst.scale_coefs(ref.get_scaling_factor(inst.q_matrix))
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            recs_r.append(ref(instance=inst).get_metadata_dict())
    out = {"instances": count, "n": n, "batch": 1000, "iterations": 1500, "solver": "DLSolver", "cores": os.cpu_count()}
    for label, recs in (("engine", recs_e), ("reference_cpu", recs_r)):
        out[label] = {"tts_s": tts.tts_table(recs, percentiles=(50.0,))[n][50.0],
                      "mean_solve_time_s": float(np.mean([r["solve_time"] for r in recs])),
                      "mean_p_optimal": float(np.mean([r["solution_performance"]["optimal"] for r in recs]))}
    out["solve_time_ratio"] = out["reference_cpu"]["mean_solve_time_s"] / out["engine"]["mean_solve_time_s"]
    te, tr = out["engine"]["tts_s"], out["reference_cpu"]["tts_s"]
    out["tts_ratio"] = tr / te if np.isfinite(te) and np.isfinite(tr) and te > 0 else None
    return out


def run_gpu_arm(args):
    import torch.distributed as dist
    from ccvm_b200 import engine as E, _native as nat, parallel as P

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 (NCCL prints its version
    # banner there) are diverted to stderr until the line is printed
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ccvm_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = nat.load()
    stream = torch.cuda.current_stream(dev)

    q_host, v_host, sb = synthetic_instance(N, 0)
    q, v = q_host.to(dev), v_host.to(dev)
    s_val = float(np.sqrt(DL["pump"] - 1))
    traj_base = rank * BATCH
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    launches = {"n": 0}

    def device_step(step_idx, ev_mid=None):
        """ONE kernel: schedules + all iterations + change of variables + adam post-processor + energy +
        solution statistics (ccvm_solve_fused) [-> pack_record -> all_gather -> merge_records]."""
        plan = E.plan_solve(nat.SOLVER_DL, nat.ALG_ADAM, q, v, BATCH, ITERS, s=1.0, pump=DL["pump"], dt=DL["dt"],
                            noise_ratio=DL["noise_ratio"], g=DL["g"], hyperparameters=HP, seed=1234, offset=step_idx,
                            traj_base=traj_base)
        epi = E.plan_epilogue(BATCH, N, dev, map1=(0.5 / s_val, 0.5), post_processor="adam", scaled_by=sb)
        res = E.solve_fused(plan, epi, 0.0)
        if ev_mid is not None:
            ev_mid.record(stream)
        launches["n"] += 1
        if distributed:
            launches["n"] += 2
            return P.merge_results(P.pack_from_stats(res, epi.pv, traj_base)), res, epi
        return res, res, epi

    # context pre-warm (not a step): module loading and the SM clock ramp of a fresh process take
    # tens of milliseconds, more than W short steps cover -- spin the FP32 probe for ~0.2 s first
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 0.2:
        E.microbench_fp32(1)
    for w in range(args.warmup):
        device_step(w)
    torch.cuda.synchronize(dev)

    # ---- timed region: K steps, each bracketed by its own event pair; L2 flushed in between
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    raw_results = []
    if distributed:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches["n"] = 0
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(float(k))
        ev[k][0].record(stream)
        last = device_step(1000 + k, ev[k][1])
        ev[k][2].record(stream)
        raw_results.append(last[1])
    torch.cuda.synchronize(dev)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    step_ms = [ev[k][0].elapsed_time(ev[k][2]) for k in range(args.steps)]
    total_ms = float(sum(step_ms))
    # the dominant kernel (at N = 1 it IS the step: one launch): duration by CUDA events on the launching stream
    solve_avg_ms = float(np.mean([ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)]))
    fused = [E.decode_fused_results(r.cpu())[0] for r in raw_results]
    loop_ms = float(np.mean([f["loop_ns"] for f in fused])) * 1e-6     # in-kernel %globaltimer, max over CTAs
    tail_ms = float(np.mean([f["tail_ns"] for f in fused])) * 1e-6
    if distributed:
        t = torch.tensor([total_ms, solve_avg_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, solve_avg_ms = t[0].item(), t[1].item()
    gpu_launches = launches["n"]

    # ---- end-to-end through the C ABI with HOST buffers (H2D of Q,V and D2H of the results per step)
    qp, vp = q_host.contiguous().pin_memory(), v_host.contiguous().pin_memory()
    h_energy = torch.empty(BATCH, dtype=torch.float32).pin_memory()
    h_stats = torch.empty(9, dtype=torch.int32).pin_memory()
    sd = nat.SolveDesc()
    sd.solver, sd.algorithm, sd.n, sd.batch, sd.iterations, sd.pump_rate_flag = nat.SOLVER_DL, nat.ALG_ADAM, N, BATCH, ITERS, 1
    sd.lower, sd.upper, sd.s = 0.0, 1.0, 1.0
    sd.pump, sd.dt, sd.noise_ratio, sd.g = DL["pump"], DL["dt"], DL["noise_ratio"], DL["g"]
    sd.alpha, sd.beta1, sd.beta2, sd.add_assign = HP["alpha"], HP["beta1"], HP["beta2"], 0
    sd.rng_mode, sd.seed, sd.traj_base = nat.RNG_PHILOX, 1234, traj_base
    ed = nat.EpilogueDesc()
    ed.apply_map1, ed.map1_scale, ed.map1_shift = 1, 0.5 / s_val, 0.5
    ed.post_processor, ed.pp_iterations, ed.pp_step, ed.pp_lower, ed.pp_upper = nat.PP_ADAM, 1, 0.01, 0.0, 1.0
    ed.scaled_by = sb

    h_best = torch.empty((), dtype=torch.float32).pin_memory()
    h_vec = torch.empty(N, dtype=torch.float32).pin_memory()

    def host_step(step_idx):
        """N = 1: the C ABI's host-buffer entry (H2D of Q, V; one fused launch; D2H of energies + statistics).
        N > 1: the same through the package API, plus the cross-rank merge of the step (pack_record,
        all_gather, merge_records) and the D2H of the merged best objective and winner's vector."""
        if not distributed:
            sd.offset = step_idx
            nat.check(lib.ccvm_solve_host(ctypes.byref(sd), ctypes.byref(ed), qp.data_ptr(), vp.data_ptr(), 0.0,
                                          h_energy.data_ptr(), h_stats.data_ptr(), stream.cuda_stream))
            return float(h_stats[:1].view(torch.float32))
        qd, vd = qp.to(dev, non_blocking=True), vp.to(dev, non_blocking=True)
        plan = E.plan_solve(nat.SOLVER_DL, nat.ALG_ADAM, qd, vd, BATCH, ITERS, s=1.0, pump=DL["pump"], dt=DL["dt"],
                            noise_ratio=DL["noise_ratio"], g=DL["g"], hyperparameters=HP, seed=1234, offset=step_idx,
                            traj_base=traj_base)
        epi = E.plan_epilogue(BATCH, N, dev, map1=(0.5 / s_val, 0.5), post_processor="adam", scaled_by=sb)
        res = E.solve_fused(plan, epi, 0.0)
        best, _, _, vec = P.merge_results(P.pack_from_stats(res, epi.pv, traj_base))
        h_energy.copy_(epi.energy, non_blocking=True)
        h_best.copy_(best, non_blocking=True)
        h_vec.copy_(vec, non_blocking=True)
        stream.synchronize()
        return float(h_best)

    for w in range(max(args.warmup, 1)):
        host_step(w)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for k in range(args.steps):
        host_step(2000 + k)   # synchronises the stream itself: results are in host memory on return
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if distributed:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t[0].item()

    # ---- post-run check (outside every timed region): a 64-trajectory slice of the production kernel's last
    # step against the CPU oracle replaying the normals that step drew (ccvm_dump_noise)
    check = None
    if rank == 0:
        try:
            check = oracle_check(E, nat, q_host, v_host, q, v, sb, s_val, traj_base)
        except Exception as e:  # noqa: BLE001
            check = {"error": f"{type(e).__name__}: {e}"}

    extras = {}
    if not args.no_extras:
        try:
            extras["strong"] = strong_record(E, nat, P, rank, world, dev, q, v, sb, s_val)
        except Exception as e:  # noqa: BLE001
            extras["strong"] = {"error": f"{type(e).__name__}: {e}"}
        for name, fn in (("sweep", lambda: sweep_record(rank, world, dev)),
                         ("config4", lambda: config4_record(E, nat, rank, world, dev))):
            try:
                extras[name] = fn()
            except Exception as e:  # noqa: BLE001
                extras[name] = {"error": f"{type(e).__name__}: {e}"}
        if rank == 0 and world == 1:
            for name, fn in (("gpu_eager_baseline", lambda: gpu_eager_record(q, v, sb)),
                             ("tts", lambda: tts_record())):
                try:
                    extras[name] = fn()
                except Exception as e:  # noqa: BLE001
                    extras[name] = {"error": f"{type(e).__name__}: {e}"}

    # ---- roofline of the dominant kernel (the persistent SDE kernel): FP32 SIMT FMA
    traffic = None   # DRAM bytes per launch of that kernel, from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "bench_kernel_traffic.json")) as fh:
            traffic = json.load(fh)["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    peak_tf = None
    cpu_baseline = None
    if rank == 0:
        peak_tf = E.microbench_fp32(1)
        flops_per_launch = 2.0 * 2 * N * N * BATCH * ITERS          # F = 2*M*N^2 per trajectory-step, M = 2 (DL)
        achieved_tf = flops_per_launch / (solve_avg_ms * 1e-3) / 1e12
        if world == 1:
            kind = "reference" if _load_reference() else "port"
            torch.set_num_threads(os.cpu_count())
            cpu_step(kind, q_host, v_host, sb, 10)
            t0 = time.perf_counter()
            reps = 0
            while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 50):
                cpu_step(kind, q_host, v_host, sb, CPU_SAMPLE_ITERS)
                reps += 1
            cdt = time.perf_counter() - t0
            cpu_baseline = {
                "value": BATCH * CPU_SAMPLE_ITERS * reps / cdt, "unit": UNIT, "cores": torch.get_num_threads(),
                "kind": kind,
                "sample": f"{reps} x ({CPU_SAMPLE_ITERS} of {ITERS} iterations, B={BATCH}, N={N}) DL-adam + adam pp + energy",
            }
        value = world * BATCH * ITERS * args.steps / (total_ms * 1e-3)
        roofline = roofline_record(E, nat, q, v, achieved_tf, peak_tf, traffic, solve_avg_ms, loop_ms, tail_ms,
                                   flops_per_launch)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "notes": GPU_NOTES,
            "e2e": {"value": world * BATCH * ITERS * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": (N * N + N) * 4,
                    "d2h_bytes_per_step": BATCH * 4 + (36 if world == 1 else 4 + 4 * N),
                    "api": "ccvm_solve_host (C ABI, pinned host buffers)" if world == 1 else
                           "plan_solve/solve_fused + merge_results (package API, pinned host buffers, cross-rank merge inside)"},
            "gpu_launches": gpu_launches,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "oracle_check": check,
            **extras,
            "wall_s_timed_region": wall,
            "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_max": float(np.max(step_ms)),
        }
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(_clean(line)), flush=True)
        os.dup2(2, 1)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the sweep / config4 / eager-GPU / TTS sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps or 5
        run_reference_arm(args)
    else:
        args.steps = args.steps or 50
        args.warmup = max(args.warmup, 3)
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
