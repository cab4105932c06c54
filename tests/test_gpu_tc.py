"""GPU parity of the tensor-core drift path (sde_kernel_tc.cuh: tcgen05 kind::tf32, 3xTF32 split, FP32
accumulation in TMEM) used for n >= 256 with >= 1024 contraction rows (BASELINE config 4).

Checked against (a) the CPU oracle under noise replay -- the same 1e-3 relative bar on the
per-trajectory objective as the SIMT kernels (2e-3 DL-adam), and (b) the SIMT kernel under the same
Philox stream (both paths draw identical noise), elementwise."""
import ctypes as C

import pytest
import torch

from oracle import ccvm_oracle as O
from ccvm_b200 import engine as E, _native as nat
from tests.test_gpu_parity import test_replay_parity_vs_oracle as replay_case, instance, HP

pytestmark = pytest.mark.gpu

TC_THREADS = 320


def _launch_info(solver, n, batch, alg=nat.ALG_ORIGINAL):
    q = torch.zeros(n, n, device="cuda")
    plan = E.plan_solve(solver, alg, q, torch.zeros(n, device="cuda"), batch, 10, s=1.0, pump=2.0, dt=0.001,
                        sigma=0.5, seed=1, offset=0, hyperparameters=HP)
    return E.query_launch(plan.desc)


def test_path_selection(monkeypatch):
    monkeypatch.delenv("CCVM_TC", raising=False)
    assert _launch_info(nat.SOLVER_DL, 1024, 8192)["threads"] == TC_THREADS      # config 4
    assert _launch_info(nat.SOLVER_DL, 1024, 8192)["ctas"] == 128
    # Langevin-type loops have one row per trajectory: 64 row blocks at B = 8192 -> two CTA pairs per block of
    # 256 rows split its four output chunks (128 CTAs); small batches split four ways
    assert _launch_info(nat.SOLVER_LANGEVIN, 1024, 8192)["ctas"] == 128
    assert _launch_info(nat.SOLVER_LANGEVIN, 1024, 2048)["ctas"] == 64
    monkeypatch.setenv("CCVM_TC_NO_SPLIT", "1")
    assert _launch_info(nat.SOLVER_LANGEVIN, 1024, 8192)["ctas"] == 64
    monkeypatch.delenv("CCVM_TC_NO_SPLIT")
    assert _launch_info(nat.SOLVER_LANGEVIN, 320, 1024)["threads"] == TC_THREADS
    assert _launch_info(nat.SOLVER_DL, 320, 512)["threads"] == TC_THREADS        # 1024 rows
    assert _launch_info(nat.SOLVER_LANGEVIN, 320, 512)["threads"] != TC_THREADS  # too few rows: SIMT
    # n = 256 is also within reach of the hybrid SIMT kernel: tensor cores only from ~10k rows
    assert _launch_info(nat.SOLVER_DL, 256, 8192)["threads"] == TC_THREADS
    assert _launch_info(nat.SOLVER_LANGEVIN, 256, 8192)["threads"] != TC_THREADS
    assert _launch_info(nat.SOLVER_DL, 256, 1024)["threads"] != TC_THREADS
    assert _launch_info(nat.SOLVER_DL, 250, 8192)["threads"] != TC_THREADS       # n < 256: SIMT
    assert _launch_info(nat.SOLVER_DL, 70, 4096)["threads"] != TC_THREADS
    monkeypatch.setenv("CCVM_TC", "0")
    assert _launch_info(nat.SOLVER_DL, 1024, 8192)["threads"] != TC_THREADS


TC_CASES = [
    # (solver, adam, n, batch, iterations, tol)   -- sizes the oracle finishes in seconds
    ("dl", False, 256, 512, 60, 1e-3), ("dl", True, 256, 512, 60, 2e-3),
    ("mf", False, 256, 1024, 60, 1e-3), ("mf", True, 256, 1024, 40, 1e-3),
    ("lv", False, 256, 1024, 60, 1e-3), ("lv", True, 256, 1024, 40, 1e-3),
    ("plv", False, 256, 1024, 60, 1e-3), ("plv", True, 256, 1024, 40, 1e-3),
    # ragged: n not a multiple of the 256-column chunk / 16-wide k-block, rows not a multiple of 128
    ("dl", False, 300, 70, 50, 1e-3), ("plv", False, 513, 131, 30, 1e-3), ("mf", False, 270, 200, 40, 1e-3),
    # several output chunks and k-ring wrap-arounds
    ("dl", False, 1024, 64, 12, 1e-3), ("lv", False, 1000, 128, 12, 1e-3),
]


# CCVM_TC selects the kernel: "2" = CTA-pair kernel (cta_group::2, the default for eligible sizes),
# "1" = single-CTA kernel (kept as the reference implementation of the same decomposition)
@pytest.mark.parametrize("ver", ["2", "1"])
@pytest.mark.parametrize("solver,adam,n,b,t,tol", TC_CASES)
def test_tc_replay_parity_vs_oracle(monkeypatch, ver, solver, adam, n, b, t, tol):
    monkeypatch.setenv("CCVM_TC", ver)
    replay_case(solver, adam, n, b, t, tol)


@pytest.mark.parametrize("ver", ["2", "1"])
@pytest.mark.parametrize("solver", ["dl", "lv", "mf"])
def test_tc_matches_simt_under_same_philox_stream(monkeypatch, solver, ver):
    n, b, t = 512, 1024, 40
    q, v, _ = instance(n, 5, 0.2 if solver == "dl" else 0.05)
    if solver == "dl":
        sid, kw = nat.SOLVER_DL, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, feedback_scale=100.0, g=0.05)
    elif solver == "lv":
        sid, kw = nat.SOLVER_LANGEVIN, dict(s=0.5, dt=0.002, sigma=0.5, feedback_scale=1.0)
    else:
        sid, kw = nat.SOLVER_MF, dict(s=20.0, pump=0.0, dt=0.0025, j=5.0, feedback_scale=4000.0, g=0.01)
    # the tensor-core path draws Philox in counter mode, the SIMT kernels draw per-thread streams: the
    # SIMT side replays the normals the tensor-core run drew (ccvm_dump_noise under the same switch)
    monkeypatch.setenv("CCVM_TC", ver)
    outs, _ = E.solve(sid, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), b, t, seed=9, offset=8, **kw)
    res = {ver: [o.clone() for o in outs]}
    noise = E.dump_noise(sid, n, b, t, 9, 8)
    monkeypatch.setenv("CCVM_TC", "0")
    outs, _ = E.solve(sid, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), b, t, noise=noise, **kw)
    res["0"] = [o.clone() for o in outs]
    for a, c in zip(res["0"], res[ver]):
        assert torch.isfinite(c).all()
        scale = a.abs().max().item()
        assert (a - c).abs().max().item() <= 2e-4 * max(scale, 1.0)
    # and the tensor-core path is reproducible bit for bit
    monkeypatch.setenv("CCVM_TC", ver)
    outs, _ = E.solve(sid, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), b, t, seed=9, offset=8, **kw)
    assert all(torch.equal(o, r) for o, r in zip(outs, res[ver]))


@pytest.mark.parametrize("solver,n,b", [("lv", 1024, 2048), ("mf", 512, 1024), ("plv", 1024, 4096), ("dl", 512, 512)])
def test_tc_column_split_matches_unsplit(monkeypatch, solver, n, b):
    """Two / four CTA pairs per row block exchanging the state through L2 (chunk flags) against one pair per
    block: same noise, same arithmetic up to the order of the k-blocks inside the tensor-core accumulation."""
    t = 25
    q, v, _ = instance(n, 9, 0.2 if solver == "dl" else 0.05)
    sid, kw = {"dl": (nat.SOLVER_DL, dict(s=1.0, pump=8.0, dt=0.001, noise_ratio=10.0, feedback_scale=100.0, g=0.05)),
               "lv": (nat.SOLVER_LANGEVIN, dict(s=0.5, dt=0.002, sigma=0.5, feedback_scale=1.0)),
               "plv": (nat.SOLVER_PUMPED_LANGEVIN, dict(s=0.5, pump=2.0, dt=0.002, sigma=0.5, feedback_scale=1.0)),
               "mf": (nat.SOLVER_MF, dict(s=20.0, pump=0.0, dt=0.0025, j=5.0, feedback_scale=4000.0, g=0.01))}[solver]
    monkeypatch.delenv("CCVM_TC_NO_SPLIT", raising=False)
    split, _ = E.solve(sid, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), b, t, seed=2, offset=6, **kw)
    split = [o.clone() for o in split]
    again, _ = E.solve(sid, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), b, t, seed=2, offset=6, **kw)
    assert all(torch.equal(a, c) for a, c in zip(split, again))       # the exchange is race-free: bit-reproducible
    monkeypatch.setenv("CCVM_TC_NO_SPLIT", "1")
    one, _ = E.solve(sid, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), b, t, seed=2, offset=6, **kw)
    for a, c in zip(one, split):
        assert torch.isfinite(c).all()
        assert (a - c).abs().max().item() <= 2e-4 * max(a.abs().max().item(), 1.0)


@pytest.mark.parametrize("ver", ["2", "1"])
def test_tc_shard_invariance(monkeypatch, ver):
    monkeypatch.setenv("CCVM_TC", ver)
    n, t = 256, 30
    q, v, _ = instance(n, 3, 0.05)
    kw = dict(s=0.5, pump=2.0, dt=0.002, sigma=0.5, feedback_scale=1.0, seed=4, offset=0)
    full, _ = E.solve(nat.SOLVER_PUMPED_LANGEVIN, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), 300, t, **kw)
    a, _ = E.solve(nat.SOLVER_PUMPED_LANGEVIN, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), 130, t, traj_base=0, **kw)
    c, _ = E.solve(nat.SOLVER_PUMPED_LANGEVIN, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), 170, t, traj_base=130, **kw)
    assert torch.equal(torch.cat([a[0], c[0]]), full[0])


@pytest.mark.parametrize("ver", ["2", "1"])
def test_tc_nan_for_nan(monkeypatch, ver):
    """A diverging trajectory must end as NaN and must not contaminate its neighbours'
    rows of the GEMM."""
    monkeypatch.setenv("CCVM_TC", ver)
    n, b, t = 256, 64, 400
    q, v, _ = instance(n, 1, 0.2)
    outs, _ = E.solve(nat.SOLVER_DL, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), b, t, s=1.0, pump=2.0, dt=0.005,
                      noise_ratio=10.0, feedback_scale=100.0, g=0.05, seed=1, offset=0)
    noise = E.dump_noise(nat.SOLVER_DL, n, b, t, 1, 0)    # the normals the tensor-core run drew
    monkeypatch.setenv("CCVM_TC", "0")
    ref, _ = E.solve(nat.SOLVER_DL, nat.ALG_ORIGINAL, q.cuda(), v.cuda(), b, t, s=1.0, pump=2.0, dt=0.005,
                     noise_ratio=10.0, feedback_scale=100.0, g=0.05, noise=noise)
    assert torch.equal(torch.isnan(outs[0]).any(dim=1), torch.isnan(ref[0]).any(dim=1))
