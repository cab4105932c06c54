"""Pin the oracle: it must reproduce every golden vector recorded from the unmodified
reference (tests/golden/make_golden.py) exactly, and the reference's own known answers."""
import numpy as np
import pytest
import torch

from oracle import ccvm_oracle as O
from tests import _cases as C


@pytest.mark.parametrize("name", C.loop_fixtures())
def test_loops_bit_exact(name):
    z = C.load(name)
    got = C.run_oracle(name, z)
    for key, exp in C.expected_outputs(name, z).items():
        assert np.array_equal(got[key].numpy(), exp), f"{name}:{key}"


def test_mf_tensor_s_bit_exact():
    z = C.load("mf_tensorS")
    z["S"] = np.float64(0)  # unused, s_vec wins
    got = C.run_oracle("mf_tensorS", z)
    for key, exp in C.expected_outputs("mf_tensorS", z).items():
        assert np.array_equal(got[key].numpy(), exp), key


def test_postprocessors_energy_stats():
    z = C.load("postproc_n20")
    q, v = torch.from_numpy(z["q"]), torch.from_numpy(z["v"])
    x0 = torch.from_numpy(z["x0"])
    assert np.array_equal(O.pp_grad_descent(x0, q, v).numpy(), z["x_gd"])
    assert np.array_equal(
        O.pp_grad_descent(x0, q, v, lower=0.1, upper=0.9, num_iter_pp=5, step_size=0.05).numpy(), z["x_gd5"])
    # torch.optim.Adam's fused arithmetic differs in rounding order; 1 step agrees to ~1e-8
    np.testing.assert_allclose(O.pp_adam(x0, q, v).numpy(), z["x_adam"], rtol=0, atol=5e-8)
    sb = torch.tensor(z["scaled_by"])
    assert np.array_equal(O.energy(x0, q, v, sb).numpy(), z["e0"])
    assert np.array_equal(O.energy(torch.from_numpy(z["x_gd"]), q, v, sb).numpy(), z["e_gd"])
    best, perf = O.solution_stats(torch.from_numpy(z["fake_obj"]), float(z["optimal"]))
    assert best == float(z["best"])
    assert list(perf.values()) == list(z["perf"])


@pytest.mark.parametrize("solver,pp", [(s, p) for s in ("dl", "mfadam", "lv", "plv")
                                       for p in ("none", "gd", "adam")])
def test_full_call_epilogues(solver, pp):
    z = C.load(f"call_{solver}_{pp}")
    q, v = torch.from_numpy(z["q"]), torch.from_numpy(z["v"])
    b, t = int(z["batch"]), int(z["iterations"])
    noise = O.NoiseSource(20, b, replay=torch.from_numpy(z["noise"]))
    ppn = {"none": None, "gd": "grad-descent", "adam": "adam"}[pp]
    sb = torch.tensor(z["scaled_by"])
    if solver == "dl":
        c, s = O.dl_solve(q, v, b, t, 8.0, 0.001, 10, 100, noise)
        pv, obj = O.epilogue("dl", c, q, v, sb, 1, post_processor=ppn)
        assert np.array_equal(s.numpy(), z["s"])
    elif solver == "mfadam":
        hp = dict(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=True)
        mu, mt, sg = O.mf_solve_adam(q, v, b, t, 20.0, 0.0, 0.0025, 5.0, 4000, noise, hp)
        pv, obj = O.epilogue("mf", mt, q, v, sb, 20.0, post_processor=ppn)
        assert np.array_equal(mu.numpy(), z["mu"]) and np.array_equal(sg.numpy(), z["sigma"])
    elif solver == "lv":
        c = O.langevin_solve(q, v, b, t, 0.5, 0.002, 0.5, 1.0, noise)
        pv, obj = O.epilogue("langevin", c, q, v, sb, 0.5, post_processor=ppn)
    else:
        c = O.pumped_langevin_solve(q, v, b, t, 0.5, 2.0, 0.002, 0.5, 1.0, noise)
        pv, obj = O.epilogue("pumped_langevin", c, q, v, sb, 0.5, post_processor=ppn)
    tol = dict(rtol=0, atol=5e-8) if pp == "adam" else dict(rtol=0, atol=0)
    np.testing.assert_allclose(pv.numpy(), z["pv"], **tol)
    np.testing.assert_allclose(obj.numpy(), z["obj"], rtol=1e-6 if pp == "adam" else 0, atol=0)
    best, perf = O.solution_stats(obj, float(z["optimal"]))
    if pp != "adam":
        assert best == float(z["best"])
        assert list(perf.values()) == list(z["perf"])


def test_reference_known_answers():
    """test_mf_solver.py:63-154 closed-form hook values; test_solution.py:140-173."""
    q, v = torch.ones(2, 2), torch.ones(2)
    mu_tilde = torch.zeros(3, 2)  # S=20, fs=400 (test_mf_solver.py:63-88)
    t1, t2 = O._mf_feedback(mu_tilde, q, v, 0, 1, 20.0)
    assert torch.equal(400 * (t1 + t2), torch.full((3, 2), -20.0))
    x4 = torch.full((2, 2), 4.0)  # test_mf_solver.py:132-154
    assert torch.equal(O.change_variables(x4, 0, 1, 2), torch.full((2, 2), 1.5))
    assert torch.equal(O.change_variables(x4, 0.2, 0.8, 2), torch.full((2, 2), 1.1))
    obj = -torch.tensor([100.0, 99.95, 90.0])
    best, perf = O.solution_stats(obj, 100.0)
    assert best == 100.0 and perf["optimal"] == 0.6667 and perf["ten_percent"] == 0.6667


def test_instance_loader():
    z = C.load("instance007")
    q, v, info = O.load_instance(C.GOLDEN + "/synthetic007.in")
    assert np.array_equal(q.numpy(), z["q"]) and np.array_equal(v.numpy(), z["v"])
    f = O.scaling_factor(q, 0.2)
    qs, vs, sb = O.scale_coefs(q, v, 1, f)
    assert np.array_equal(qs.numpy(), z["q_scaled"]) and np.array_equal(vs.numpy(), z["v_scaled"])
    assert info["optimal_sol"] == float(z["optimal"]) and info["num_frac_values"] == int(z["num_frac"])
    assert info["solution_vector"] == list(z["solution_vector"])
