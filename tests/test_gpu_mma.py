"""The small-n tensor-core kernel (csrc/sde_kernel_mma.cuh: drift contraction on tcgen05 with FP16-split operands,
Qs^T resident in tensor memory) against the oracle and against the tiled kernels.

It serves single-instance Philox-mode launches with 40 <= n <= 128 and batch >= 2048 ... 3584 (by n and solver), and with
128 < n <= 192 (two M tiles) from batch 256 (DL) / 640 (CCVM_MMA=1 forces it for any
shape it can run, CCVM_MMA=0 disables it).  Same bar as every production kernel: the oracle replays the normals
``ccvm_dump_noise`` writes for the launch, per-trajectory objective within 1e-3 relative (2e-3 DL-adam; reference loops
dl_solver.py:468-769, mf_solver.py:493-764, langevin_solver.py:368-561, pumped_langevin_solver.py:232-449)."""
import struct

import numpy as np
import pytest
import torch

from oracle import ccvm_oracle as O
from ccvm_b200 import engine as E, _native as nat
from tests.test_gpu_parity import HP, instance, parity_case
from tests.test_gpu_production_parity import TILES, launch_info, tol_of

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_mma(monkeypatch):
    monkeypatch.setenv("CCVM_MMA", "1")


def test_size_rule(monkeypatch):
    """Which launches take the tensor-core kernel (288 threads: two update warpgroups + the MMA issuer warp)."""
    monkeypatch.delenv("CCVM_MMA", raising=False)
    assert launch_info("dl", True, 70, 4096, 10)["threads"] == 288
    info = launch_info("dl", True, 70, 4096, 10)
    assert info["ctas"] == 147 and info["traj_per_cta"] == 28       # 148 SMs: 7 pairs per warpgroup
    assert launch_info("lv", False, 128, 2048, 10)["threads"] == 288
    assert launch_info("lv", False, 40, 4096, 10)["threads"] == 288
    assert launch_info("lv", False, 40, 3000, 10)["threads"] != 288   # measured crossover of the K = 1 loops at n = 40: ~3600
    assert launch_info("dl", False, 40, 2048, 10)["threads"] == 288
    assert launch_info("mf", True, 70, 2560, 10)["threads"] == 288
    assert launch_info("mf", True, 70, 2048, 10)["threads"] != 288
    assert launch_info("lv", False, 36, 4096, 10)["threads"] != 288   # too few variables: tiled kernel
    assert launch_info("lv", False, 70, 1000, 10)["threads"] != 288   # too few trajectories per SM
    # two M tiles: 128 < n <= 192, DL from batch 256, the K = 1 loops from 640; above 192 the hybrid kernel
    assert launch_info("lv", False, 129, 4096, 10)["threads"] == 288
    assert launch_info("lv", False, 160, 4096, 10)["ctas"] == 147          # 9 items per lane: one wave of CTAs
    assert launch_info("dl", True, 160, 4096, 10)["ctas"] == 256           # 8 items per lane at most: two waves, 4 pairs per warpgroup
    assert launch_info("mf", False, 192, 640, 10)["threads"] == 288
    assert launch_info("mf", False, 192, 600, 10)["threads"] != 288
    assert launch_info("dl", False, 150, 256, 10)["threads"] == 288
    assert launch_info("dl", False, 150, 200, 10)["threads"] != 288
    assert launch_info("lv", False, 193, 4096, 10)["threads"] != 288  # hybrid kernel
    monkeypatch.setenv("CCVM_MMA", "0")
    assert launch_info("dl", True, 70, 4096, 10)["threads"] == 256


# the shapes the size rule selects: every tile, K extents 48 ... 192, one and two M tiles, 4 ... 8 pairs per warpgroup,
# 3 ... 11 items per lane
@pytest.mark.parametrize("solver,adam", TILES)
@pytest.mark.parametrize("n,b,t", [(40, 3600, 60), (64, 2600, 60), (70, 4096, 60), (100, 2048, 40), (128, 2240, 40),
                                   (129, 1000, 30), (150, 4096, 30), (177, 700, 40), (192, 4096, 20)])
def test_production_parity_selected_shapes(solver, adam, n, b, t):
    assert launch_info(solver, adam, n, b, t)["threads"] == 288
    parity_case(solver, adam, n, b, t, tol_of(solver, adam), philox=(31, 7 * n + b))


# forced: ragged sizes (n % 4 != 0, n % 16 != 0), odd batches (a half-filled pair, a partly filled CTA), one pair per
# warpgroup, tiny n
@pytest.mark.parametrize("solver,adam", TILES)
@pytest.mark.parametrize("n,b,t", [(33, 129, 100), (47, 301, 100), (70, 1001, 80), (113, 75, 60), (20, 64, 100), (5, 37, 60),
                                   (131, 75, 50), (161, 301, 40), (190, 37, 40)])
def test_production_parity_forced_shapes(force_mma, solver, adam, n, b, t):
    assert launch_info(solver, adam, n, b, t)["threads"] == 288
    parity_case(solver, adam, n, b, t, tol_of(solver, adam), philox=(77, 5 * n + b))


def test_per_variable_saturation_and_bounds(force_mma):
    """Tensor S (per-variable clamp and drift scaling) and non-default solution bounds on the tensor-core path: the
    oracle on the dumped noise."""
    n, b, t = 50, 300, 120
    q, v, sb = instance(n, 9, 0.05)
    s_vec = torch.linspace(0.3, 0.9, n)
    noise = E.dump_noise(nat.SOLVER_LANGEVIN, n, b, t, 5, 9).cpu()
    assert launch_info("lv", False, n, b, t)["threads"] == 288
    c_ref = O.langevin_solve(q, v, b, t, s_vec, 0.002, 0.5, 1.0, O.NoiseSource(n, b, replay=noise), bounds=(-1.0, 2.0))
    outs, _ = E.solve(nat.SOLVER_LANGEVIN, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), b, t, s_vec=s_vec.cuda(), dt=0.002,
                      sigma=0.5, feedback_scale=1.0, lower=-1.0, upper=2.0, seed=5, offset=9)
    assert (outs[0].cpu() - c_ref).abs().max().item() <= 2e-4


def test_reproducible_and_shard_invariant():
    """Same bits for the same (seed, offset); a batch split at an even index gives the same trajectories although the
    shards run other launch geometries (noise is keyed by global trajectory pair and variable)."""
    n, t = 70, 80
    q, v, _ = instance(n, 2, 0.05)
    qg, vg = q.cuda(), v.cuda()
    kw = dict(s=0.5, pump=2.0, dt=0.002, sigma=0.5, feedback_scale=1.0, seed=11, offset=4)
    full, _ = E.solve(nat.SOLVER_PUMPED_LANGEVIN, nat.ALG_ORIGINAL, qg, vg, 5400, t, **kw)   # two waves of CTAs
    full = full[0].clone()
    again, _ = E.solve(nat.SOLVER_PUMPED_LANGEVIN, nat.ALG_ORIGINAL, qg, vg, 5400, t, **kw)
    assert torch.equal(full, again[0])
    assert launch_info("plv", False, n, 2600, t)["threads"] == 288
    a, _ = E.solve(nat.SOLVER_PUMPED_LANGEVIN, nat.ALG_ORIGINAL, qg, vg, 2600, t, traj_base=0, **kw)
    bb, _ = E.solve(nat.SOLVER_PUMPED_LANGEVIN, nat.ALG_ORIGINAL, qg, vg, 2800, t, traj_base=2600, **kw)
    assert torch.equal(torch.cat([a[0], bb[0]]), full)
    other, _ = E.solve(nat.SOLVER_PUMPED_LANGEVIN, nat.ALG_ORIGINAL, qg, vg, 5400, t, **dict(kw, seed=12))
    assert not torch.equal(full, other[0])


def test_noise_is_standard_normal():
    """Q = V = 0 and a huge S: Langevin's c_T is sigma sqrt(dt) times a sum of T normals (the per-(pair, variable)
    streams of the tensor-core kernel)."""
    n, b, t = 64, 2560, 64
    q, v = torch.zeros(n, n), torch.zeros(n)
    assert launch_info("lv", False, n, b, t)["threads"] == 288
    outs, _ = E.solve(nat.SOLVER_LANGEVIN, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), b, t, s=1e6, dt=1.0 / t, sigma=1.0,
                      feedback_scale=1.0, seed=5, offset=0)
    x = outs[0].double().cpu().flatten()
    m = x.numel()
    assert abs(x.mean().item()) < 5 / np.sqrt(m)
    assert abs(x.var().item() - 1.0) < 5 * np.sqrt(2.0 / m)
    assert abs(((x - x.mean()) ** 4).mean().item() / x.var().item() ** 2 - 3.0) < 0.1
    c = np.corrcoef(outs[0].cpu().numpy()[:, :8].T)          # across variables
    assert np.abs(c - np.eye(8)).max() < 5 / np.sqrt(b)
    c = np.corrcoef(outs[0].cpu().numpy()[:8, :])            # across trajectories (incl. the two of a pair)
    assert np.abs(c - np.eye(8)).max() < 5 / np.sqrt(n)
    noise = E.dump_noise(nat.SOLVER_LANGEVIN, n, b, t, 5, 0)
    assert torch.allclose(noise.sum(0)[0].T * (1.0 / t) ** 0.5, outs[0], atol=2e-5)


def test_nan_for_nan(force_mma):
    """The README quick-start with feedback_scale = 100 diverges to NaN in the reference: so must the
    tensor-core path (its FP16 operands saturate instead of overflowing; the FP32 state still blows up)."""
    q, v, _ = instance(20, 1, 0.2)
    b, t = 16, 600
    noise = E.dump_noise(nat.SOLVER_DL, 20, b, t, 3, 1).cpu()
    c_ref, _ = O.dl_solve(q, v, b, t, 2.0, 0.005, 10.0, 100.0, O.NoiseSource(20, b, replay=noise))
    outs, _ = E.solve(nat.SOLVER_DL, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), b, t, s=1.0, pump=2.0, dt=0.005,
                      noise_ratio=10.0, feedback_scale=100.0, g=0.05, seed=3, offset=1)
    assert torch.isnan(c_ref).all() and torch.isnan(outs[0]).all()


@pytest.mark.parametrize("pp", [None, "grad-descent", "adam"])
def test_fused_tail_equals_stand_alone_epilogue(monkeypatch, pp):
    """One launch per Solver.__call__ on the tensor-core path too: the tail of the kernel (change of variables ->
    post-processor -> energy -> statistics) gives what the stand-alone epilogue and statistics kernels give on the
    same final state."""
    monkeypatch.delenv("CCVM_MMA", raising=False)
    n, b, t = 70, 2100, 50
    q0, v0 = O.synthetic_boxqp(n, 4)
    f = float(O.scaling_factor(q0, 0.2))
    q, v = (q0 / f).cuda(), (v0 / f).cuda()
    kw = dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, g=0.05, hyperparameters=HP, seed=9, offset=2)
    s_map = float(np.sqrt(7.0))
    ekw = dict(map1=(0.5 / s_map, 0.5), post_processor=pp, pp_iterations=3, scaled_by=f)
    plan = E.plan_solve(nat.SOLVER_DL, nat.ALG_ADAM, q, v, b, t, **kw)
    epi = E.plan_epilogue(b, n, torch.device("cuda", torch.cuda.current_device()), **ekw)
    res = E.solve_fused(plan, epi, optimal_value=100.0)
    torch.cuda.synchronize()
    pv_f, en_f, c_f = epi.pv.clone(), epi.energy.clone(), plan.outputs[0].clone()
    outs, _ = E.solve(nat.SOLVER_DL, nat.ALG_ADAM, q, v, b, t, **kw)
    assert torch.equal(outs[0], c_f)
    pv_s, en_s = E.epilogue(outs[0], q, v, **ekw)
    assert torch.equal(pv_s, pv_f)
    assert torch.allclose(en_s, en_f, rtol=2e-6, atol=1e-5 * float(en_s.abs().max()))
    best, arg = struct.unpack("fi", bytes(res.cpu().numpy()[:8]))
    assert arg == int(torch.argmax(-en_f).item()) and abs(best - float((-en_f).max())) <= 1e-6 * abs(best)


def test_solver_call_through_tensor_core_path(monkeypatch):
    """DLSolver.__call__ (adam algorithm + adam post-processor, the benchmark configuration) on a bundled N = 70
    instance at a batch that selects the tensor-core kernel: same solution quality as the tiled kernel (different
    noise streams, so statistics, not bits)."""
    from ccvm_b200.solvers import DLSolver
    from ccvm_b200.solvers.algorithms import AdamParameters
    from tools.equivalence_gpu import load_bundled
    inst = load_bundled()[70][3]
    hp = AdamParameters(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=False)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("CCVM_MMA", mode)
        torch.manual_seed(5)
        solver = DLSolver(device="cuda", batch_size=4096)
        solver.parameter_key = {70: dict(pump=8.0, dt=0.001, iterations=1500, noise_ratio=10, feedback_scale=100)}
        if mode == "1":
            inst.scale_coefs(solver.get_scaling_factor(inst.q_matrix))
        res[mode] = solver(instance=inst, post_processor="adam", algorithm_parameters=hp)
    a, b = res["1"], res["0"]
    # two independent samples of 4096 trajectories (calibration: tools/mma_call_stats.py, four seeds per kernel --
    # objective mean 72, standard deviation 38, best_objective_value = max(-E) between 40 and 58 for EITHER kernel):
    # means within 5 standard errors; the 1 % / 99 % quantiles within 5 standard errors of a sample quantile
    # (sqrt(p (1 - p) / n) / density, Gaussian density at 2.33 sigma: 0.058 sigma each, 0.082 sigma for the difference);
    # the best value within 4 Gumbel scales sigma / sqrt(2 ln n) of a sample maximum
    oa, ob = a.objective_values.double().cpu(), b.objective_values.double().cpu()
    se = float(np.sqrt(oa.var().item() / oa.numel() + ob.var().item() / ob.numel()))
    assert abs(oa.mean().item() - ob.mean().item()) <= 5 * se, (oa.mean().item(), ob.mean().item(), se)
    sd = float(np.sqrt(0.5 * (oa.var().item() + ob.var().item())))
    assert abs(oa.std().item() - ob.std().item()) <= 5 * sd / np.sqrt(oa.numel()), (oa.std().item(), ob.std().item())
    for p in (0.01, 0.99):
        qa, qb = torch.quantile(oa, p).item(), torch.quantile(ob, p).item()
        assert abs(qa - qb) <= 5 * 0.082 * sd, (p, qa, qb, sd)
    assert abs(a.best_objective_value - b.best_objective_value) <= 4 * sd / np.sqrt(2 * np.log(oa.numel())), \
        (a.best_objective_value, b.best_objective_value, sd)
    for key in ("one_percent", "five_percent", "ten_percent"):
        pa, pb = a.solution_performance[key], b.solution_performance[key]
        assert abs(pa - pb) <= 4 * np.sqrt(max(pb * (1 - pb), 1e-3) * 2 / 4096) + 1e-3, (key, pa, pb)
