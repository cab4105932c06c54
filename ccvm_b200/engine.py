"""Thin tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers and
the current CUDA stream to ``libccvm_b200.so``.  No arithmetic happens here."""
import ctypes as C

import torch

from . import _native as nat

GAP_NAMES = ("optimal", "one_percent", "two_percent", "three_percent", "four_percent",
             "five_percent", "ten_percent")


def _device_of(t):
    if not t.is_cuda:
        raise nat.NativeError("ccvm_b200 engine tensors must live on a CUDA device; there is no CPU path.")
    return t.device


def next_philox_stream(device, increment):
    """(seed, offset) taken from -- and advanced on -- torch's CUDA generator of ``device`` so that
    ``torch.manual_seed`` keeps governing reproducibility (the reference has no seed argument and
    draws from the global generator, dl_solver.py:512-519)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    gen = torch.cuda.default_generators[idx]
    seed = gen.initial_seed()
    offset = gen.get_offset()
    gen.set_offset(offset + 4 * ((int(increment) + 3) // 4))
    return seed & 0xFFFFFFFFFFFFFFFF, offset


class _NvtxRange:
    """NVTX range around one solver call (tracing hooks); a no-op when torch was built
    without NVTX."""

    def __init__(self, name):
        self.name, self.on = name, False

    def __enter__(self):
        try:
            torch.cuda.nvtx.range_push(self.name)
            self.on = True
        except Exception:
            self.on = False
        return self

    def __exit__(self, *exc):
        if self.on:
            torch.cuda.nvtx.range_pop()
        return False


def nvtx_range(name):
    return _NvtxRange(name)


class PlannedSolve:
    """A filled ``ccvm_solve_desc`` together with the tensors its pointers refer to."""

    __slots__ = ("desc", "outputs", "samples", "device", "_keep")

    def __init__(self, desc, outputs, samples, device, keep):
        self.desc, self.outputs, self.samples, self.device, self._keep = desc, outputs, samples, device, keep


def plan_solve(solver, algorithm, q, v, batch, iterations, *, lower=0.0, upper=1.0, s=1.0, s_vec=None,
               pump=0.0, dt=0.0, noise_ratio=1.0, j=1.0, sigma=0.0, feedback_scale=1.0, g=0.0,
               pump_rate_flag=True, hyperparameters=None, noise=None, seed=None, offset=None,
               traj_base=0, evolution_step=None, num_samples=0):
    """Fill the descriptor of one ``_solve`` / ``_solve_adam`` loop and allocate its outputs without
    launching anything (``solve`` launches one, ``solve_batch`` many in a single launch)."""
    nat.require_cuda()
    dev = _device_of(q)
    n = int(q.shape[0])
    qc, vc = nat.as_f32(q, dev), nat.as_f32(v, dev)
    d = nat.SolveDesc()
    d.solver, d.algorithm = solver, algorithm
    d.n, d.batch, d.iterations = n, int(batch), int(iterations)
    d.pump_rate_flag = 1 if pump_rate_flag else 0
    d.q, d.v = nat.ptr(qc), nat.ptr(vc)
    d.lower, d.upper = float(lower), float(upper)
    svc = None
    if s_vec is not None:
        svc = nat.as_f32(s_vec, dev)
        if svc.numel() != n:
            raise ValueError("Tensor S size should be equal to problem size.")
        d.s_vec = nat.ptr(svc)
        d.s = 0.0
    else:
        d.s = float(s)
    d.pump, d.dt, d.noise_ratio, d.j = float(pump), float(dt), float(noise_ratio), float(j)
    d.sigma, d.feedback_scale, d.g = float(sigma), float(feedback_scale), float(g)
    if algorithm == nat.ALG_ADAM:
        d.alpha, d.beta1, d.beta2 = (float(hyperparameters[k]) for k in ("alpha", "beta1", "beta2"))
        d.add_assign = 1 if hyperparameters["add_assign"] else 0
    k_draws = 2 if solver == nat.SOLVER_DL else 1
    nz = None
    if noise is not None:
        nz = nat.as_f32(noise, dev)
        if nz.dim() != 4 or nz.shape[0] < iterations or nz.shape[1] != k_draws or nz.shape[2] != n:
            raise ValueError(f"replay noise must be [iterations][{k_draws}][n][batch], got {tuple(nz.shape)}")
        d.rng_mode, d.noise, d.noise_batch = nat.RNG_REPLAY, nat.ptr(nz), int(nz.shape[3])
    else:
        if seed is None:
            seed, offset = next_philox_stream(dev, 1)
        d.rng_mode, d.seed, d.offset = nat.RNG_PHILOX, int(seed), int(offset or 0)
    d.traj_base = int(traj_base)
    n_out = {nat.SOLVER_DL: 2, nat.SOLVER_MF: 3}.get(solver, 1)
    outs = [torch.empty((batch, n), dtype=torch.float32, device=dev) for _ in range(n_out)]
    d.out0 = nat.ptr(outs[0])
    d.out1 = nat.ptr(outs[1]) if n_out > 1 else None
    d.out2 = nat.ptr(outs[2]) if n_out > 2 else None
    samples = None
    if evolution_step:
        n_state = 1 if solver in (nat.SOLVER_LANGEVIN, nat.SOLVER_PUMPED_LANGEVIN) else 2
        samples = torch.zeros((n_state, num_samples, batch, n), dtype=torch.float32, device=dev)
        d.evolution_step, d.num_samples, d.samples = int(evolution_step), int(num_samples), nat.ptr(samples)
    return PlannedSolve(d, tuple(outs), samples, dev, (qc, vc, svc, nz))


def solve(solver, algorithm, q, v, batch, iterations, **kwargs):
    """Run one ``_solve`` / ``_solve_adam`` loop on the GPU.  Returns (outputs, samples):
    outputs is the tuple of (B, N) state tensors documented in ccvm_b200.h."""
    lib = nat.load()
    plan = plan_solve(solver, algorithm, q, v, batch, iterations, **kwargs)
    with torch.cuda.device(plan.device):
        nat.check(lib.ccvm_solve(C.byref(plan.desc), nat.current_stream_ptr(plan.device)))
    # inputs were kept alive until the launch was enqueued on the current stream
    return plan.outputs, plan.samples


def dump_noise(solver, n, batch, iterations, seed, offset, traj_base=0, device=None, launch_batch=None):
    """The standard normals a Philox-mode solve with these (seed, offset, traj_base) draws, as a replay
    tensor [iterations][K][n][batch] (``ccvm_dump_noise``): validation only.  ``launch_batch``: batch of the launch
    this is a slice of (the kernel family, and with it the generator, is chosen from the launch's batch)."""
    nat.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    d = nat.SolveDesc()
    d.solver, d.n, d.batch, d.iterations = solver, int(n), int(batch), int(iterations)
    d.seed, d.offset, d.traj_base = int(seed), int(offset), int(traj_base)
    if launch_batch is not None:
        d.noise_batch = int(launch_batch)
    k = 2 if solver == nat.SOLVER_DL else 1
    out = torch.empty((iterations, k, n, batch), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_dump_noise(C.byref(d), nat.ptr(out), nat.current_stream_ptr(dev)))
    return out


def solve_batch(plans):
    """Launch many planned solves (same solver and algorithm, Philox noise) as ONE grid over
    instances x trajectory blocks (``ccvm_solve_batch``).  Outputs are those of the plans."""
    if not plans:
        return
    lib = nat.load()
    dev = plans[0].device
    arr = (nat.SolveDesc * len(plans))(*[p.desc for p in plans])
    with torch.cuda.device(dev):
        nat.check(lib.ccvm_solve_batch(arr, len(plans), nat.current_stream_ptr(dev)))


def query_launch(desc):
    info = (C.c_int32 * 5)()
    nat.check(nat.load().ccvm_query_launch(C.byref(desc), info))
    return dict(threads=info[0], ctas=info[1], traj_per_cta=info[2], smem=info[3], regs=info[4])


def _vec_or_scalar(val, n, dev):
    """(scalar, tensor-or-None) for an affine-map coefficient that may be per-variable."""
    if torch.is_tensor(val) and val.numel() > 1:
        if val.dim() == 2:  # (B, N) outer(ones, S): rows are identical by construction
            val = val[0]
        if val.numel() != n:
            raise ValueError("Tensor S size should be equal to problem size.")
        return 0.0, nat.as_f32(val, dev)
    return float(val), None


def _set_scaled_by(desc, scaled_by, dev, keep):
    """``instance.scaled_by`` as a host scalar, or -- when it is a one-element CUDA tensor (the product of
    scaling factors reduced on the device) -- as a device pointer: no ``.item()`` synchronisation."""
    if torch.is_tensor(scaled_by) and scaled_by.is_cuda and scaled_by.numel() == 1:
        t = nat.as_f32(scaled_by, dev).reshape(1)
        keep.append(t)
        desc.scaled_by, desc.scaled_by_dev = 1.0, nat.ptr(t)
    else:
        desc.scaled_by = float(scaled_by.item()) if torch.is_tensor(scaled_by) else float(scaled_by)


class PlannedEpilogue:
    """A filled ``ccvm_epilogue_desc`` with its output tensors (pv, energy)."""

    __slots__ = ("desc", "pv", "energy", "_keep")

    def __init__(self, desc, pv, energy, keep):
        self.desc, self.pv, self.energy, self._keep = desc, pv, energy, keep


def plan_epilogue(b, n, dev, *, map1=None, post_processor=None, pp_iterations=10, pp_step=None, pp_lower=0.0,
                  pp_upper=1.0, map2=None, scaled_by=1.0, want_pv=True, energy_out=None):
    """Descriptor of the tail of ``Solver.__call__`` for a fused launch (``solve_fused`` /
    ``solve_batch_fused``): q, v and the state come from the solve it is attached to."""
    d = nat.EpilogueDesc()
    d.n, d.batch = n, b
    keep = []
    if map1 is not None:
        sc, vec = _vec_or_scalar(map1[0], n, dev)
        keep.append(vec)
        d.apply_map1, d.map1_scale, d.map1_shift, d.map1_scale_vec = 1, sc, float(map1[1]), nat.ptr(vec)
    if map2 is not None:
        sc, vec = _vec_or_scalar(map2[0], n, dev)
        keep.append(vec)
        d.apply_map2, d.map2_scale, d.map2_shift, d.map2_scale_vec = 1, sc, float(map2[1]), nat.ptr(vec)
    if post_processor not in nat.PP_IDS:
        raise AssertionError(f"Method type is not valid. Provided: {post_processor}")
    d.post_processor = nat.PP_IDS[post_processor]
    d.pp_iterations = int(pp_iterations)
    if pp_step is None:
        pp_step = 0.1 if post_processor == "grad-descent" else 0.01
    d.pp_step, d.pp_lower, d.pp_upper = float(pp_step), float(pp_lower), float(pp_upper)
    _set_scaled_by(d, scaled_by, dev, keep)
    pv = torch.empty((b, n), dtype=torch.float32, device=dev) if want_pv else None
    en = energy_out if energy_out is not None else torch.empty((b,), dtype=torch.float32, device=dev)
    if en.numel() != b or en.dtype != torch.float32 or not en.is_contiguous() or en.device != dev:
        raise ValueError("energy_out must be a contiguous fp32 tensor of batch elements on the state's device")
    d.problem_variables, d.energy = nat.ptr(pv), nat.ptr(en)
    return PlannedEpilogue(d, pv, en, keep)


def decode_fused_results(raw):
    """Host view of ``count`` 56-byte result blocks (a CPU uint8 tensor): list of dicts with best,
    arg_best, counts, ctas, loop_ns, tail_ns."""
    raw = raw.contiguous().view(-1, nat.FUSED_RESULT_BYTES)
    best = raw[:, 0:4].contiguous().view(torch.float32).reshape(-1).tolist()
    ints = raw[:, 4:40].contiguous().view(torch.int32).reshape(-1, 9).tolist()
    ns = raw[:, 40:56].contiguous().view(torch.int64).reshape(-1, 2).tolist()
    return [dict(best=best[i], arg_best=ints[i][0], counts=ints[i][1:8], ctas=ints[i][8], loop_ns=ns[i][0],
                 tail_ns=ns[i][1]) for i in range(raw.shape[0])]


def solve_fused(plan, epi, optimal_value=0.0, want_stats=True):
    """ONE launch for a whole ``Solver.__call__``: schedules, loop, change of variables, post-processor,
    energy, statistics (``ccvm_solve_fused``).  Returns the device result block (uint8[56]) or None."""
    lib = nat.load()
    dev = plan.device
    res = torch.empty(nat.FUSED_RESULT_BYTES, dtype=torch.uint8, device=dev) if want_stats else None
    with torch.cuda.device(dev):
        nat.check(lib.ccvm_solve_fused(C.byref(plan.desc), C.byref(epi.desc), float(optimal_value), nat.ptr(res),
                                       nat.current_stream_ptr(dev)))
    return res


def solve_batch_fused(plans, epis, optimal_values):
    """Many planned solves AND their tails in one launch per bucket (``ccvm_solve_batch_fused``).
    Returns the device result blocks, uint8[count][56]."""
    lib = nat.load()
    dev = plans[0].device
    count = len(plans)
    arr = (nat.SolveDesc * count)(*[p.desc for p in plans])
    earr = (nat.EpilogueDesc * count)(*[e.desc for e in epis])
    opt = (C.c_double * count)(*[float(o) for o in optimal_values])
    res = torch.empty((count, nat.FUSED_RESULT_BYTES), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        nat.check(lib.ccvm_solve_batch_fused(arr, earr, opt, count, nat.ptr(res), nat.current_stream_ptr(dev)))
    return res


def epilogue(state, q, v, *, map1=None, post_processor=None, pp_iterations=10, pp_step=None,
             pp_lower=0.0, pp_upper=1.0, map2=None, scaled_by=1.0, want_energy=True, want_pv=True,
             energy_out=None):
    """Fused tail of ``Solver.__call__``: x = state*m1s + m1o -> post-processor -> pv ;
    energy((pv*m2s + m2o)).  ``map1`` / ``map2`` are (scale, shift) or None."""
    nat.require_cuda()
    lib = nat.load()
    dev = _device_of(state)
    st = nat.as_f32(state, dev)
    b, n = st.shape
    qc, vc = nat.as_f32(q, dev), nat.as_f32(v, dev)
    d = nat.EpilogueDesc()
    d.n, d.batch = n, b
    d.q, d.v, d.state = nat.ptr(qc), nat.ptr(vc), nat.ptr(st)
    keep = []
    if map1 is not None:
        sc, vec = _vec_or_scalar(map1[0], n, dev)
        keep.append(vec)
        d.apply_map1, d.map1_scale, d.map1_shift, d.map1_scale_vec = 1, sc, float(map1[1]), nat.ptr(vec)
    if map2 is not None:
        sc, vec = _vec_or_scalar(map2[0], n, dev)
        keep.append(vec)
        d.apply_map2, d.map2_scale, d.map2_shift, d.map2_scale_vec = 1, sc, float(map2[1]), nat.ptr(vec)
    if post_processor not in nat.PP_IDS:
        raise AssertionError(f"Method type is not valid. Provided: {post_processor}")
    d.post_processor = nat.PP_IDS[post_processor]
    d.pp_iterations = int(pp_iterations)
    if pp_step is None:
        pp_step = 0.1 if post_processor == "grad-descent" else 0.01
    d.pp_step, d.pp_lower, d.pp_upper = float(pp_step), float(pp_lower), float(pp_upper)
    _set_scaled_by(d, scaled_by, dev, keep)
    pv = torch.empty((b, n), dtype=torch.float32, device=dev) if want_pv else None
    en = None
    if want_energy:
        en = energy_out if energy_out is not None else torch.empty((b,), dtype=torch.float32, device=dev)
        if en.numel() != b or en.dtype != torch.float32 or not en.is_contiguous() or en.device != dev:
            raise ValueError("energy_out must be a contiguous fp32 tensor of batch elements on the state's device")
    d.problem_variables, d.energy = nat.ptr(pv), nat.ptr(en)
    with torch.cuda.device(dev):
        nat.check(lib.ccvm_epilogue(C.byref(d), nat.current_stream_ptr(dev)))
    del keep, qc, vc, st
    return pv, en


def compute_energy(x, q, v, scaled_by=1.0):
    """E_b = (1/2 x_b Q x_b + V x_b) * scaled_by on the device; result returned where x lives."""
    home = x.device
    xg = to_engine_device(x)
    dev = xg.device
    xc, qc, vc = nat.as_f32(xg, dev), nat.as_f32(q, dev), nat.as_f32(v, dev)
    b, n = xc.shape
    en = torch.empty((b,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_compute_energy(nat.ptr(xc), nat.ptr(qc), nat.ptr(vc), float(scaled_by), b, n,
                                                 nat.ptr(en), nat.current_stream_ptr(dev)))
    return en.to(home)


def postprocess_grad_descent(x, q, v, iterations, step_size, lower, upper):
    """In place on ``x`` (must be a contiguous fp32 CUDA tensor)."""
    nat.require_cuda()
    dev = _device_of(x)
    qc, vc = nat.as_f32(q, dev), nat.as_f32(v, dev)
    b, n = x.shape
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_postprocess_grad_descent(nat.ptr(x), nat.ptr(qc), nat.ptr(vc), b, n,
                                                           int(iterations), float(step_size), float(lower),
                                                           float(upper), nat.current_stream_ptr(dev)))
    return x


def postprocess_adam(x, q, v, lr, lower, upper):
    nat.require_cuda()
    dev = _device_of(x)
    qc, vc = nat.as_f32(q, dev), nat.as_f32(v, dev)
    b, n = x.shape
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_postprocess_adam(nat.ptr(x), nat.ptr(qc), nat.ptr(vc), b, n, float(lr),
                                                   float(lower), float(upper), nat.current_stream_ptr(dev)))
    return x


def solution_stats(energy, optimal_value):
    """(best_objective_value, arg_best, counts[7]) with ONE device->host read of 36 bytes
    (the reference does eight .item() syncs, solution.py:82,113-136)."""
    nat.require_cuda()
    dev = _device_of(energy)
    en = nat.as_f32(energy, dev)
    res = torch.empty(9, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_solution_stats(nat.ptr(en), en.numel(), float(optimal_value), nat.ptr(res),
                                                 nat.current_stream_ptr(dev)))
    host = res.cpu()
    best = host[:1].view(torch.float32).item()
    return best, int(host[1]), [int(c) for c in host[2:9]]


def solution_stats_batch(energy, offsets, optimal_values):
    """Statistics of many instances with one kernel and ONE device->host copy.  ``energy`` is the
    concatenation of the instances' energy vectors (device), ``offsets`` the count+1 boundaries,
    ``optimal_values`` the per-instance optima.  Returns a list of (best, arg_best, counts[7])."""
    nat.require_cuda()
    dev = _device_of(energy)
    count = len(optimal_values)
    en = nat.as_f32(energy, dev)
    off = torch.tensor(list(offsets), dtype=torch.int64).to(dev, non_blocking=True)
    opt = torch.tensor(list(optimal_values), dtype=torch.float32).to(dev, non_blocking=True)
    res = torch.empty((count, 9), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_solution_stats_batch(nat.ptr(en), nat.ptr(off), nat.ptr(opt), count,
                                                       nat.ptr(res), nat.current_stream_ptr(dev)))
    host = res.cpu()
    best = host[:, 0].contiguous().view(torch.float32).tolist()
    rest = host[:, 1:].tolist()
    return [(best[i], rest[i][0], rest[i][1:]) for i in range(count)]


def scaling_factor(q, multiplier):
    """0-d fp32 tensor sqrt(sum|Q|) * multiplier (ccvm_solver.py:134-150), reduced on the device and
    returned where q lives."""
    home = q.device
    qg = to_engine_device(q)
    dev = qg.device
    qc = nat.as_f32(qg, dev)
    out = torch.empty((), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_scaling_factor(nat.ptr(qc), int(qc.shape[0]), float(multiplier), nat.ptr(out),
                                                 nat.current_stream_ptr(dev)))
    return out.to(home)


def generate_boxqp(n, seed, device=None, q_offdiag_std=None, v_std=20.0):
    """(Q, V) of a synthetic dense symmetric BoxQP instance drawn on the device (``ccvm_generate_boxqp``):
    coefficient statistics of the bundled benchmarking instances (off-diagonal std
    28.5/sqrt(N), V std 20), reference sign convention, deterministic in (n, seed)."""
    nat.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    q = torch.empty((n, n), dtype=torch.float32, device=dev)
    v = torch.empty((n,), dtype=torch.float32, device=dev)
    std = 28.5 / n ** 0.5 if q_offdiag_std is None else float(q_offdiag_std)
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_generate_boxqp(nat.ptr(q), nat.ptr(v), int(n), int(seed) & 0xFFFFFFFFFFFFFFFF, std,
                                                 float(v_std), nat.current_stream_ptr(dev)))
    return q, v


def microbench_fp32(mode=1):
    """Measured register-only FP32 FMA throughput in TFLOP/s (0 = FFMA, 1 = packed FFMA2)."""
    nat.require_cuda()
    val = C.c_double(0.0)
    nat.check(nat.load().ccvm_microbench_fp32(int(mode), C.byref(val), nat.current_stream_ptr()))
    return val.value


def microbench_tf32(mode=2):
    """Measured dense TF32 tensor-pipe throughput in TFLOP/s of the tcgen05.mma shape the SDE kernels
    use (1 = cta_group::1, 2 = cta_group::2): the roofline denominator of the tensor-core path."""
    nat.require_cuda()
    val = C.c_double(0.0)
    nat.check(nat.load().ccvm_microbench_tf32(int(mode), C.byref(val), nat.current_stream_ptr()))
    return val.value


def to_engine_device(t):
    """Tensors handed to a kernel must be on a CUDA device; CPU tensors are copied to the current
    one (plumbing).  Without a GPU this raises -- there is no CPU implementation to fall back to."""
    nat.require_cuda()
    return t if t.is_cuda else t.to("cuda")


def _s_args(S, n, dev):
    """(scalar, device-vector-or-None) for a saturation value that may be a 1-D / (B, N) tensor."""
    if torch.is_tensor(S) and S.numel() > 1:
        vec = S if S.dim() == 1 else S[0]
        if vec.numel() != n:
            raise ValueError("Tensor S size should be equal to problem size.")
        return 0.0, nat.as_f32(vec, dev)
    return (float(S.item()) if torch.is_tensor(S) else float(S)), None


def eval_hook(solver, kind, q, v, inputs, lower, upper, S, pump=0.0, rate=0.0, feedback_scale=0.0, j=0.0,
              g=0.0):
    """calculate_drift / calculate_grads of ``solver`` on explicit inputs; returns a tuple of
    tensors on the inputs' device."""
    src = inputs[0]
    home = src.device
    ins = [nat.as_f32(to_engine_device(t), to_engine_device(t).device) for t in inputs]
    dev = ins[0].device
    b, n = ins[0].shape
    qc, vc = nat.as_f32(q, dev), nat.as_f32(v, dev)
    d = nat.HookDesc()
    d.solver, d.kind, d.n, d.batch = solver, 1 if kind == "drift" else 0, n, b
    d.q, d.v = nat.ptr(qc), nat.ptr(vc)
    d.in0 = nat.ptr(ins[0])
    d.in1 = nat.ptr(ins[1]) if len(ins) > 1 else None
    d.in2 = nat.ptr(ins[2]) if len(ins) > 2 else None
    d.lower, d.upper = float(lower), float(upper)
    s_val, s_vec = _s_args(S, n, dev)
    d.s, d.s_vec = s_val, nat.ptr(s_vec)
    d.pump, d.rate, d.feedback_scale, d.j, d.g = float(pump), float(rate), float(feedback_scale), float(j), float(g)
    two = solver == nat.SOLVER_DL or (solver == nat.SOLVER_MF and kind == "drift")
    o0 = torch.empty((b, n), dtype=torch.float32, device=dev)
    o1 = torch.empty((b, n), dtype=torch.float32, device=dev) if two else None
    d.out0, d.out1 = nat.ptr(o0), nat.ptr(o1)
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_eval_hook(C.byref(d), nat.current_stream_ptr(dev)))
    outs = (o0, o1) if two else (o0,)
    return tuple(o.to(home) for o in outs)


def change_variables(x, lower, upper, S):
    home = x.device
    xc = nat.as_f32(to_engine_device(x), to_engine_device(x).device)
    dev = xc.device
    b, n = xc.shape
    s_val, s_vec = _s_args(S, n, dev)
    out = torch.empty_like(xc)
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_change_variables(nat.ptr(xc), nat.ptr(out), b, n, float(lower), float(upper),
                                                   s_val, nat.ptr(s_vec), nat.current_stream_ptr(dev)))
    return out.to(home)


def clamp(x, lo, hi):
    """torch.clamp(x, lo, hi) with scalar or tensor bounds ((N,) or x-shaped), on the device."""
    home = x.device
    xc = nat.as_f32(to_engine_device(x), to_engine_device(x).device)
    dev = xc.device
    b, n = xc.shape
    lo_t = nat.as_f32(lo, dev) if torch.is_tensor(lo) and lo.numel() > 1 else None
    hi_t = nat.as_f32(hi, dev) if torch.is_tensor(hi) and hi.numel() > 1 else None
    blen = 0
    for t in (lo_t, hi_t):
        if t is not None:
            if blen and t.numel() != blen:
                raise ValueError("clamp bounds must have matching shapes")
            blen = t.numel()
    lo_s = 0.0 if lo_t is not None else float(lo)
    hi_s = 0.0 if hi_t is not None else float(hi)
    if (lo_t is None) != (hi_t is None):  # mixed scalar/tensor bounds: materialise the scalar one
        fill = torch.full((blen,), lo_s if lo_t is None else hi_s, dtype=torch.float32, device=dev)
        lo_t, hi_t = (fill, hi_t) if lo_t is None else (lo_t, fill)
    out = torch.empty_like(xc)
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_fit_to_constraints(nat.ptr(xc), nat.ptr(out), b, n, lo_s, hi_s, nat.ptr(lo_t),
                                                     nat.ptr(hi_t), blen, nat.current_stream_ptr(dev)))
    return out.to(home)


def scale_coefs(q, v, factor):
    """(q / factor, v / factor) computed on the device; results live where q lived."""
    home = q.device
    qc = nat.as_f32(to_engine_device(q), to_engine_device(q).device)
    dev = qc.device
    vc = nat.as_f32(v, dev)
    n = qc.shape[0]
    f = factor if torch.is_tensor(factor) else torch.tensor(float(factor))
    f = nat.as_f32(f, dev).reshape(-1)
    if f.numel() not in (1, n * n):
        raise ValueError("scaling_factor must be a scalar or an (n, n) tensor")
    q_out = torch.empty_like(qc)
    v_out = torch.empty((n,) if f.numel() == 1 else (n, n), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nat.check(nat.load().ccvm_scale_coefs(nat.ptr(qc), nat.ptr(vc), n, nat.ptr(f), f.numel(), nat.ptr(q_out),
                                              nat.ptr(v_out), nat.current_stream_ptr(dev)))
    return q_out.to(home), v_out.to(home)
