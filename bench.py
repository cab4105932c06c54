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
    return {
        "workload": "configs[2]: DLSolver._solve_adam (pump 8, dt 0.001, noise_ratio 10, g 0.05; Adam alpha 1e-3, "
                    "beta 0.9/0.999, add_assign False) + adam post-processor + compute_energy + solution stats",
        "n": N, "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "iterations": ITERS,
        "instance": "synthetic dense BoxQP, seed 1000 (SURVEY 8d generator)", "rng": "philox4x32-10 in-kernel",
        "l2": "L2 flushed (256 MiB write) before every timed step, outside the per-step event pair",
        "parallelism": f"batch sharded over {n_gpus} GPU(s), no data-path collective; one all_gather of N+9 floats per step",
    }


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
    cfg["parallelism"] = "single host process, torch intra-op threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
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
        """solve (1 schedule + 1 persistent SDE kernel) -> fused epilogue -> stats [-> all_gather]."""
        outs, _ = E.solve(nat.SOLVER_DL, nat.ALG_ADAM, q, v, BATCH, ITERS, s=1.0, pump=DL["pump"], dt=DL["dt"],
                          noise_ratio=DL["noise_ratio"], g=DL["g"], hyperparameters=HP, seed=1234, offset=step_idx,
                          traj_base=traj_base)
        if ev_mid is not None:
            ev_mid.record(stream)
        pv, en = E.epilogue(outs[0], q, v, map1=(0.5 / s_val, 0.5), post_processor="adam", scaled_by=sb)
        res = torch.empty(9, dtype=torch.int32, device=dev)
        nat.check(lib.ccvm_solution_stats(en.data_ptr(), BATCH, 0.0, res.data_ptr(), stream.cuda_stream))
        launches["n"] += 4
        if distributed:
            launches["n"] += 2
            return P.merge_results(P.pack_from_stats(res, pv, traj_base))
        return res

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
    if distributed:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches["n"] = 0
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(float(k))
        ev[k][0].record(stream)
        device_step(1000 + k, ev[k][1])
        ev[k][2].record(stream)
    torch.cuda.synchronize(dev)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    step_ms = [ev[k][0].elapsed_time(ev[k][2]) for k in range(args.steps)]
    solve_ms = [ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)]
    total_ms = float(sum(step_ms))
    solve_avg_ms = float(np.mean(solve_ms))
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

    def host_step(step_idx):
        sd.offset = step_idx
        nat.check(lib.ccvm_solve_host(ctypes.byref(sd), ctypes.byref(ed), qp.data_ptr(), vp.data_ptr(), 0.0,
                                      h_energy.data_ptr(), h_stats.data_ptr(), stream.cuda_stream))
        return float(h_stats[:1].view(torch.float32))

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
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": world * BATCH * ITERS * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": (N * N + N) * 4, "d2h_bytes_per_step": BATCH * 4 + 36,
                    "api": "ccvm_solve_host (C ABI, pinned host buffers)"},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic,
                         "kernel": "ccvm::sde_tmem_kernel<DL, adam>", "kernel_ms": solve_avg_ms,
                         "peak_source": "measured in-process: register-only FFMA2 probe (ccvm_microbench_fp32); "
                                        "MEASURED_PEAKS.json has no FP32 SIMT figure (HBM/bf16 only)",
                         "algorithmic_flops_per_launch": flops_per_launch},
            "cpu_baseline": cpu_baseline,
            "wall_s_timed_region": wall,
            "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_max": float(np.max(step_ms)),
        }
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
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
