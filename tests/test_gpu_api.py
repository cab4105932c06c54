"""GPU tests of the reference-facing API (solver classes, post-processors, instance, Solution)
against golden vectors from the unmodified reference and against the CPU oracle."""
import io
import os

import numpy as np
import pytest
import torch

from oracle import ccvm_oracle as O
from tests import _cases as C

pytestmark = pytest.mark.gpu

from ccvm_b200.solvers import DLSolver, MFSolver, LangevinSolver, PumpedLangevinSolver, AdamParameters  # noqa: E402
from ccvm_b200.problem_classes.boxqp import ProblemInstance  # noqa: E402
from ccvm_b200.post_processor import PostProcessorFactory  # noqa: E402
from ccvm_b200.solution import Solution  # noqa: E402
from ccvm_b200 import engine as E  # noqa: E402

PERF = ("optimal", "one_percent", "two_percent", "three_percent", "four_percent", "five_percent", "ten_percent")


def make_instance(z, device="cuda", name="golden"):
    inst = ProblemInstance(device=device, instance_type="test", name=name)
    q = torch.from_numpy(z["q"]).to(device)
    inst.problem_size = q.shape[0]
    inst.q_matrix, inst.v_vector = q, torch.from_numpy(z["v"]).to(device)
    inst.scaled_by = torch.tensor(float(z["scaled_by"]), device=device) if "scaled_by" in z else 1
    inst.optimal_sol = float(z["optimal"]) if "optimal" in z else 1.0
    inst.best_sol, inst.num_frac_values, inst.solution_vector = inst.optimal_sol, 0, []
    return inst


def close(got, exp, tol, what=""):
    got = np.asarray(got.detach().cpu() if torch.is_tensor(got) else got, dtype=np.float64)
    exp = np.asarray(exp, dtype=np.float64)
    err = np.abs(got - exp) / np.maximum(1.0, np.abs(exp))
    assert err.max() <= tol, f"{what}: {err.max():.3e} > {tol}"


@pytest.mark.parametrize("solver,pp", [(s, p) for s in ("dl", "mfadam", "lv", "plv") for p in ("none", "gd", "adam")])
def test_full_call_matches_reference(solver, pp):
    """Solver.__call__ end to end in noise-replay mode vs the reference's own __call__."""
    z = C.load(f"call_{solver}_{pp}")
    b, t = int(z["batch"]), int(z["iterations"])
    ppn = {"none": None, "gd": "grad-descent", "adam": "adam"}[pp]
    inst = make_instance(z)
    kw = {}
    if solver == "dl":
        s = DLSolver(device="cuda", batch_size=b)
        s.parameter_key = {20: dict(pump=8.0, dt=0.001, iterations=t, noise_ratio=10, feedback_scale=100)}
    elif solver == "mfadam":
        s = MFSolver(device="cuda", batch_size=b)
        s.parameter_key = {20: dict(pump=0.0, feedback_scale=4000, j=5.0, S=20.0, dt=0.0025, iterations=t)}
        kw["algorithm_parameters"] = AdamParameters(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=True)
    elif solver == "lv":
        s = LangevinSolver(device="cuda", batch_size=b)
        s.parameter_key = {20: dict(dt=0.002, S=0.5, iterations=t, sigma=0.5, feedback_scale=1.0)}
    else:
        s = PumpedLangevinSolver(device="cuda", batch_size=b)
        s.parameter_key = {20: dict(pump=2.0, dt=0.002, S=0.5, iterations=t, sigma=0.5, feedback_scale=1.0)}
    s.noise_source = torch.from_numpy(z["noise"]).cuda()
    sol = s(instance=inst, post_processor=ppn, **kw)
    assert sol.variables["problem_variables"].is_cuda and sol.objective_values.is_cuda
    close(sol.variables["problem_variables"], z["pv"], 2e-4, "problem_variables")
    close(sol.objective_values, z["obj"], 2e-4, "objective")
    assert abs(sol.best_objective_value - float(z["best"])) <= 2e-4 * abs(float(z["best"]))
    for k, name in enumerate(PERF):
        assert abs(sol.solution_performance[name] - z["perf"][k]) <= 1.0 / b + 1e-9
    assert sol.solve_time > 0 and (sol.pp_time > 0) == (ppn is not None)
    if solver == "dl":
        close(sol.variables["s"], z["s"], 2e-4, "s")
    if solver == "mfadam":
        close(sol.variables["mu"], z["mu"], 2e-4, "mu")
        close(sol.variables["sigma"], z["sigma"], 2e-4, "sigma")
    meta = sol.get_metadata_dict()
    assert set(meta) == {"problem_size", "batch_size", "instance_name", "iterations", "solve_time", "pp_time",
                         "optimal_value", "best_value", "num_frac_values", "solution_vector", "evolution_file",
                         "solution_performance", "best_objective_value"}


def test_dl_adam_through_call_works():
    """The reference's DLSolver.__call__ + AdamParameters raises TypeError; here it
    runs and equals _solve_adam called directly."""
    z = C.load("call_dl_none")
    b, t = int(z["batch"]), int(z["iterations"])
    inst = make_instance(z)
    s = DLSolver(device="cuda", batch_size=b)
    s.parameter_key = {20: dict(pump=8.0, dt=0.001, iterations=t, noise_ratio=10, feedback_scale=100)}
    s.noise_source = torch.from_numpy(z["noise"]).cuda()
    hp = AdamParameters(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=False)
    sol = s(instance=inst, algorithm_parameters=hp)
    q, v = torch.from_numpy(z["q"]), torch.from_numpy(z["v"])
    c_ref, s_ref = O.dl_solve_adam(q, v, b, t, 8.0, 0.001, 10, O.NoiseSource(20, b, replay=torch.from_numpy(z["noise"])),
                                   hp.to_dict())
    close(sol.variables["problem_variables"], c_ref.numpy(), 2e-4, "c")
    close(sol.variables["s"], s_ref.numpy(), 2e-4, "s")


def test_reference_hook_known_answers():
    """tests/unit/solvers/test_mf_solver.py:63-204 of the reference, on the device."""
    s = MFSolver(device="cuda", batch_size=3)
    s.q_matrix, s.v_vector = torch.ones(2, 2, device="cuda"), torch.ones(2, device="cuda")
    z = torch.zeros(3, 2, device="cuda")
    grads = s._calculate_grads_boxqp(mu_tilde=z, S=20.0, fs=400)
    assert torch.equal(grads.cpu(), torch.full((3, 2), -20.0))
    dmu, dsg = s._calculate_drift_boxqp(mu=z, mu_tilde=z, sigma=z, pump=2.5, j=399, g=0.1, S=20.0, fs=400)
    assert torch.equal(dmu.cpu(), torch.full((3, 2), -20.0))
    assert torch.equal(dsg.cpu(), torch.full((3, 2), 200.5))
    x4 = torch.full((2, 2), 4.0, device="cuda")
    assert torch.equal(s._change_variables_boxqp(problem_variables=x4, S=2).cpu(), torch.full((2, 2), 1.5))
    assert torch.equal(s._change_variables_boxqp(problem_variables=x4, lower_limit=0.2, upper_limit=0.8, S=2).cpu(),
                       torch.full((2, 2), 1.1))
    mt = torch.tensor([[0.5, -0.5], [3.0, -3.0]], device="cuda")
    assert torch.equal(s._fit_to_constraints_boxqp(mt, -1.0, 1.0).cpu(), torch.tensor([[0.5, -0.5], [1.0, -1.0]]))
    lo = torch.tensor([[-1.0, -0.25], [-2.0, -1.0]], device="cuda")
    hi = -lo
    assert torch.equal(s._fit_to_constraints_boxqp(mt, lo, hi).cpu(), torch.tensor([[0.5, -0.25], [2.0, -1.0]]))


@pytest.mark.parametrize("cls", ["dl", "lv", "plv"])
def test_hooks_match_oracle(cls):
    g = torch.Generator().manual_seed(4)
    n, b = 13, 5
    q, v = torch.randn(n, n, generator=g), torch.randn(n, generator=g)  # NON-symmetric Q: yQ, not Qy
    c, s_ = torch.randn(b, n, generator=g), torch.randn(b, n, generator=g)
    lo, up, S = -0.5, 1.5, 0.7
    if cls == "dl":
        sol = DLSolver("cuda")
        sol.q_matrix, sol.v_vector = q.cuda(), v.cuda()
        gc, gs = sol._calculate_grads_boxqp(c.cuda(), s_.cuda(), lo, up, S)
        g1, g3 = O._dl_feedback(c, q, v, lo, up, S)
        close(gc, (-g1 - g3).numpy(), 1e-5)
        dc, ds = sol._calculate_drift_boxqp(c.cuda(), s_.cuda(), 3.0, 0.4, 50.0, lo, up)
        sd = np.sqrt(2.0)
        g1, g3 = O._dl_feedback(c, q, v, lo, up, sd)
        ref = -(50.0 * 0.9) * (g1 + g3) + (-1 + 1.2 - c**2 - s_**2) * c
        close(dc, ref.numpy(), 1e-5)
    elif cls == "lv":
        sol = LangevinSolver("cuda")
        sol.q_matrix, sol.v_vector = q.cuda(), v.cuda()
        close(sol._calculate_drift_boxqp(c.cuda(), lo, up, S), O._langevin_grad(c, q, v, lo, up, S).numpy(), 1e-5)
        close(sol._calculate_grads_boxqp(c.cuda(), lo, up, S), O._langevin_grad(c, q, v, lo, up, S).numpy(), 1e-5)
    else:
        sol = PumpedLangevinSolver("cuda")
        sol.q_matrix, sol.v_vector, sol.solution_bounds = q.cuda(), v.cuda(), (lo, up)
        close(sol._calculate_grads_boxqp(c.cuda(), lo, up, S), O._pl_grad(c, q, v, lo, up, S).numpy(), 1e-5)
        ref = (-1 + 1.3 - c**2) * c + 2.0 * O._pl_grad(c, q, v, lo, up, S)
        close(sol._calculate_drift_boxqp(c.cuda(), 1.3, S, 2.0), ref.numpy(), 1e-5)


def test_postprocessors_energy_stats_vs_reference_golden():
    z = C.load("postproc_n20")
    q, v = torch.from_numpy(z["q"]).cuda(), torch.from_numpy(z["v"]).cuda()
    x0 = torch.from_numpy(z["x0"]).cuda()
    keep = x0.clone()
    gd = PostProcessorFactory.create_postprocessor("grad-descent")
    close(gd.postprocess(x0, q, v), z["x_gd"], 1e-5, "gd")
    assert gd.pp_time > 0 and torch.equal(x0, keep)
    close(gd.postprocess(x0, q, v, lower_clamp=0.1, upper_clamp=0.9, num_iter_pp=5, step_size=0.05), z["x_gd5"], 1e-5)
    ad = PostProcessorFactory.create_postprocessor("adam")
    close(ad.postprocess(x0, q, v), z["x_adam"], 1e-6, "adam")
    close(ad.postprocess(x0, q, v, num_iter=5), z["x_adam"], 1e-6, "adam x5 == x1")
    # CPU tensors in, CPU tensors out (the kernel still runs on the GPU)
    out = gd.postprocess(x0.cpu(), q.cpu(), v.cpu())
    assert not out.is_cuda
    close(out, z["x_gd"], 1e-5)
    inst = make_instance(z)
    close(inst.compute_energy(x0), z["e0"], 1e-5, "energy")
    close(inst.compute_energy(torch.from_numpy(z["x_gd"]).cuda()), z["e_gd"], 1e-5)
    obj = torch.from_numpy(z["fake_obj"]).cuda()
    sol = Solution(problem_size=20, batch_size=10, instance_name="x", iterations=1, objective_values=obj,
                   solve_time=0.0, pp_time=0.0, optimal_value=float(z["optimal"]), best_value=0.0, num_frac_values=0,
                   solution_vector=[], variables={"problem_variables": x0}, device="cuda")
    assert sol.best_objective_value == float(z["best"])
    assert [sol.solution_performance[k] for k in PERF] == list(z["perf"])
    assert sol.best_index == 0


def test_solution_stats_reference_known_answers_and_nan():
    """test_solution.py:140-173 style known answers; NaN objectives -> best nan, fractions 0."""
    obj = -torch.tensor([100.0, 99.95, 90.0])
    best, arg, counts = E.solution_stats(obj.cuda(), 100.0)
    assert best == 100.0 and arg == 0 and counts == [2, 2, 2, 2, 2, 2, 2]
    best, arg, counts = E.solution_stats(torch.tensor([float("nan"), -1.0]).cuda(), 100.0)
    assert np.isnan(best) and counts == [0] * 7
    big = -torch.linspace(50, 100, 100000)
    best, arg, counts = E.solution_stats(big.cuda(), 100.0)
    ref_best, ref_perf = O.solution_stats(big, 100.0)
    assert best == ref_best and arg == 99999
    assert [round(c / 100000, 4) for c in counts] == list(ref_perf.values())


def test_instance_scaling_on_device(golden_dir):
    z = np.load(os.path.join(golden_dir, "instance007.npz"))
    inst = ProblemInstance(instance_type="test", file_path=os.path.join(golden_dir, "synthetic007.in"), device="cuda")
    assert inst.q_matrix.is_cuda and np.array_equal(inst.q_matrix.cpu().numpy(), z["q"])
    f = DLSolver("cuda").get_scaling_factor(inst.q_matrix)
    assert f.is_cuda and f.dim() == 0 and abs(f.item() - float(z["factor"])) <= 1e-6 * float(z["factor"])
    inst.scale_coefs(f)
    close(inst.q_matrix, z["q_scaled"], 1e-6)
    close(inst.v_vector, z["v_scaled"], 1e-6)
    assert abs(float(inst.scaled_by) - float(z["scaled_by"])) <= 1e-6 * float(z["scaled_by"])
    # reference test_problem_instance.py:137-158: an (n, n) tensor factor broadcasts against v
    inst2 = ProblemInstance(device="cuda", instance_type="test")
    inst2.q_matrix = torch.tensor([[41.0, 35.0], [20.0, 30.0]], device="cuda")
    inst2.v_vector = torch.tensor([-31.0, -37.0], device="cuda")
    fac = torch.tensor([[1.0, 2.0], [10.0, 10.0]], device="cuda")
    inst2.scale_coefs(fac)
    assert torch.equal(inst2.q_matrix.cpu(), torch.tensor([[41.0, 17.5], [2.0, 3.0]]))
    assert torch.equal(inst2.v_vector.cpu(), torch.tensor([[-31.0, -18.5], [-3.1, -3.7]]))
    assert torch.equal(inst2.scaled_by.cpu(), fac.cpu())


@pytest.mark.parametrize("solver", ["dl", "mf", "lv", "plv"])
def test_evolution_sampling(solver, tmp_path):
    """Snapshots at i % step == 0 and at the last iteration (dl_solver.py:557-564), file format."""
    z = C.load("call_dl_none" if solver == "dl" else f"call_{'mfadam' if solver == 'mf' else solver}_none")
    b, t, step = int(z["batch"]), 23, 5
    inst = make_instance(z, name="evo")
    noise = torch.from_numpy(z["noise"])[:t]
    q, v = torch.from_numpy(z["q"]), torch.from_numpy(z["v"])
    snaps = []
    rec = lambda i, *arrs: snaps.append([a.clone() for a in arrs]) if (i % step == 0 or i + 1 >= t) else None  # noqa: E731
    src = O.NoiseSource(20, b, replay=noise)
    if solver == "dl":
        s = DLSolver(device="cuda", batch_size=b)
        s.parameter_key = {20: dict(pump=8.0, dt=0.001, iterations=t, noise_ratio=10, feedback_scale=100)}
        O.dl_solve(q, v, b, t, 8.0, 0.001, 10, 100, src, snapshots=rec)
        names = ("c_sample", "s_sample")
    elif solver == "mf":
        s = MFSolver(device="cuda", batch_size=b)
        s.parameter_key = {20: dict(pump=0.0, feedback_scale=4000, j=5.0, S=20.0, dt=0.0025, iterations=t)}
        O.mf_solve(q, v, b, t, 20.0, 0.0, 0.0025, 5.0, 4000, src, snapshots=rec)
        names = ("mu_sample", "sigma_sample")
    elif solver == "lv":
        s = LangevinSolver(device="cuda", batch_size=b)
        s.parameter_key = {20: dict(dt=0.002, S=0.5, iterations=t, sigma=0.5, feedback_scale=1.0)}
        O.langevin_solve(q, v, b, t, 0.5, 0.002, 0.5, 1.0, src, snapshots=rec)
        names = ("c_sample",)
    else:
        s = PumpedLangevinSolver(device="cuda", batch_size=b)
        s.parameter_key = {20: dict(pump=2.0, dt=0.002, S=0.5, iterations=t, sigma=0.5, feedback_scale=1.0)}
        O.pumped_langevin_solve(q, v, b, t, 0.5, 2.0, 0.002, 0.5, 1.0, src, snapshots=rec)
        names = ("c_sample",)
    s.noise_source = noise.cuda()
    path = str(tmp_path / "evo.txt")
    sol = s(instance=inst, evolution_step_size=step, evolution_file=path)
    assert sol.evolution_file == path
    num_samples = int(t / step) + 1 + (1 if t % step else 0)
    assert len(snaps) == num_samples - 0 or len(snaps) == num_samples  # 0,5,10,15,20,22 -> 6
    for k, name in enumerate(names):
        got = getattr(s, name)
        assert got.shape == (b, 20, num_samples) and not got.is_cuda
        for si, arrs in enumerate(snaps):
            close(got[:, :, si], arrs[k].numpy(), 2e-4, f"{name}[{si}]")
    lines = open(path).read().split("\n")
    assert len(lines) == 20 * len(names) + 1
    best = sol.best_index
    row0 = [str(round(getattr(s, names[0])[best, 0, i].item(), 4)) for i in range(num_samples)]
    expect = "\t".join(row0) if solver == "mf" else "".join(c + "\t" for c in row0)
    assert lines[0] == expect
    with pytest.raises(ValueError, match="evolution step size must be greater than or equal to 1"):
        s(instance=inst, evolution_step_size=-1)


def test_errors_on_gpu_path():
    z = C.load("call_lv_none")
    inst = make_instance(z)
    s = LangevinSolver(device="cuda", batch_size=4)
    s.parameter_key = {20: dict(dt=0.002, S=torch.ones(7), iterations=5, sigma=0.5, feedback_scale=1.0)}
    with pytest.raises(ValueError, match="Tensor S size should be equal to problem size."):
        s(instance=inst)
    s.parameter_key = {20: dict(dt=0.002, S=0.5, iterations=5, sigma=0.5, feedback_scale=1.0)}
    with pytest.raises(ValueError, match="is not supported"):
        s(instance=inst, algorithm_parameters={"alpha": 1})
    with pytest.raises(AssertionError, match="Method type is not valid"):
        s(instance=inst, post_processor="nope")
    s.fit_to_constraints = lambda *a: None
    with pytest.raises(RuntimeError, match="was replaced"):
        s(instance=inst)


def test_device_instance_generator():
    """ccvm_generate_boxqp: dense symmetric, the coefficient statistics of the bundled instances
    (SURVEY.md 8d), deterministic in (n, seed), usable as a ProblemInstance source."""
    from ccvm_b200 import sweep
    n = 250
    q, v = E.generate_boxqp(n, 17)
    q2, v2 = E.generate_boxqp(n, 17)
    q3, _ = E.generate_boxqp(n, 18)
    assert torch.equal(q, q2) and torch.equal(v, v2) and not torch.equal(q, q3)
    assert torch.equal(q, q.T)
    off = q[~torch.eye(n, dtype=torch.bool, device=q.device)]
    sigma = 28.5 / n ** 0.5
    m = off.numel() / 2          # independent entries
    assert abs(off.mean().item()) < 5 * sigma / m ** 0.5
    assert abs(off.std().item() / sigma - 1) < 5 / (2 * m) ** 0.5
    assert abs(q.diagonal().std().item() / (2 ** 0.5 * sigma) - 1) < 5 / (2 * n) ** 0.5
    big_v = torch.cat([E.generate_boxqp(200, s)[1] for s in range(50)])
    assert abs(big_v.std().item() / 20.0 - 1) < 0.05 and abs(big_v.mean().item()) < 1.0
    # statistics of a bundled instance of the same size, for reference
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "bundled_instances.npz"))
    q70 = torch.from_numpy(z["q70"][0])
    g70, _ = E.generate_boxqp(70, 3)
    mask = ~torch.eye(70, dtype=torch.bool)
    assert abs(g70.cpu()[mask].std().item() / q70[mask].std().item() - 1) < 0.1
    inst = sweep.synthetic_instance(40, 5, 0.05, on_device=True)
    solver = LangevinSolver(device="cuda", batch_size=64)
    solver.parameter_key = {40: dict(dt=0.002, S=0.5, sigma=0.5, feedback_scale=1.0, iterations=50)}
    sol = solver(instance=inst, post_processor="grad-descent")
    assert torch.isfinite(sol.objective_values).all()


def test_roofline_probes():
    """The two roofline denominators (bench.py, tools/tensor_peak.py) land where a B200 can be:
    FP32 FFMA2 near 148 x 128 x 2 x f, dense TF32 tcgen05 between half and all of the nominal 1.19 PFLOP/s."""
    assert 55.0 < E.microbench_fp32(1) < 80.0
    assert 55.0 < E.microbench_fp32(0) < 80.0
    for mode in (1, 2):
        tf = E.microbench_tf32(mode)
        assert 600.0 < tf < 1300.0, (mode, tf)


def test_device_record_pack_and_merge_match_host_logic():
    """ccvm_pack_record / ccvm_merge_records (one kernel each) against the torch restatement in
    ccvm_b200/parallel.py that the gloo tests pin on CPU."""
    from ccvm_b200 import parallel as P, _native as nat
    torch.manual_seed(3)
    n, b = 37, 501
    recs_dev, recs_ref = [], []
    for r in range(4):
        en = torch.randn(b, device="cuda") * 5 - 100
        if r == 2:
            en[17] = en.min() - 1.0          # rank 2 holds the global winner ...
        if r == 3:
            en[5] = en[5]                    # (no tie here; ties are checked below)
        pv = torch.rand(b, n, device="cuda")
        res = torch.empty(9, dtype=torch.int32, device="cuda")
        nat.check(nat.load().ccvm_solution_stats(en.data_ptr(), b, 95.0, res.data_ptr(), nat.current_stream_ptr()))
        counts = res[2:9]
        ref = P.pack_local_result(en, pv, counts, 1000 * r + (1 << 33))   # index beyond 2^32: both halves in use
        dev = P.pack_from_stats(res, pv, 1000 * r + (1 << 33))
        assert torch.equal(ref.view(torch.int32), dev.view(torch.int32))
        recs_dev.append(dev)
        recs_ref.append(ref.cpu())
    H = P.HEADER
    gathered = torch.stack(recs_dev).contiguous()
    out = torch.empty_like(recs_dev[0])
    nat.check(nat.load().ccvm_merge_records(gathered.data_ptr(), 4, n, out.data_ptr(), nat.current_stream_ptr()))
    g = torch.stack(recs_ref)
    owner = int(torch.argmin(g[:, 0]))
    assert owner == 2
    best, idx, counts, vec = P.unpack_record(out, slot0_is_objective=True)
    assert best.item() == -g[owner, 0].item() and idx.item() == 2017 + (1 << 33)
    assert torch.equal(counts.cpu(), g.view(torch.int32)[:, 3:H].sum(dim=0, dtype=torch.int32))
    assert torch.equal(vec.cpu(), g[owner, H:])
    # ties go to the lowest rank
    gathered[3, 0] = gathered[2, 0]
    gathered[1, 0] = gathered[2, 0]
    nat.check(nat.load().ccvm_merge_records(gathered.data_ptr(), 4, n, out.data_ptr(), nat.current_stream_ptr()))
    assert P.unpack_record(out, True)[1].item() == P.unpack_record(recs_dev[1])[1].item()
    # a NaN objective wins on the device and in the host branch alike
    gathered[3, 0] = float("nan")
    nat.check(nat.load().ccvm_merge_records(gathered.data_ptr(), 4, n, out.data_ptr(), nat.current_stream_ptr()))
    assert torch.isnan(out[0])
    # merge_results on a CUDA record without a process group = identity reduction of one record
    best, idx, counts, vec = P.merge_results(recs_dev[0])
    assert best.item() == -recs_dev[0][0].item() and torch.equal(vec, recs_dev[0][H:])


@pytest.mark.parametrize("n,b", [(1, 3), (5, 17), (33, 129), (70, 1000), (129, 40), (250, 333), (300, 50), (600, 9)])
@pytest.mark.parametrize("pp", [None, "grad-descent"])
def test_tiled_epilogue_matches_per_trajectory_kernel(monkeypatch, n, b, pp):
    """The tiled epilogue (8 trajectories share every Q element) against the per-trajectory kernel it
    replaced (CCVM_EPILOGUE_LEGACY): identical dot-product order, so the post-processed variables
    are bit-identical; the energy differs only by the association of its final sum."""
    torch.manual_seed(n + b)
    q0, v0 = O.synthetic_boxqp(n, 2)
    f = O.scaling_factor(q0, 0.05)
    q, v = (q0 / f).cuda(), (v0 / f).cuda()
    state = (torch.rand(b, n, device="cuda") - 0.5)
    kw = dict(map1=(1.0, 0.5), post_processor=pp, pp_iterations=10, map2=(0.5, 0.25), scaled_by=float(f))
    monkeypatch.delenv("CCVM_EPILOGUE_LEGACY", raising=False)
    pv_t, e_t = E.epilogue(state, q, v, **kw)
    monkeypatch.setenv("CCVM_EPILOGUE_LEGACY", "1")
    pv_l, e_l = E.epilogue(state, q, v, **kw)
    assert torch.equal(pv_t, pv_l)
    assert torch.allclose(e_t, e_l, rtol=2e-6, atol=1e-5 * float(e_l.abs().max()))


def test_schedule_table_from_its_own_kernel_matches_the_inline_one(monkeypatch):
    """Very large batch x iteration counts fall back to ONE schedule table built by a kernel of its own (the
    per-CTA copies of the in-kernel prologue would add up): same table, bit-identical results."""
    q0, v0 = O.synthetic_boxqp(33, 5)
    f = O.scaling_factor(q0, 0.2)
    q, v = (q0 / f).cuda(), (v0 / f).cuda()
    kw = dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, g=0.05, seed=3, offset=8,
              hyperparameters=dict(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=True))
    monkeypatch.delenv("CCVM_NO_SCHED_INLINE", raising=False)
    from ccvm_b200 import _native as nat
    a, _ = E.solve(nat.SOLVER_DL, nat.ALG_ADAM, q, v, 130, 77, **kw)
    a = [t.clone() for t in a]
    monkeypatch.setenv("CCVM_NO_SCHED_INLINE", "1")
    b, _ = E.solve(nat.SOLVER_DL, nat.ALG_ADAM, q, v, 130, 77, **kw)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
