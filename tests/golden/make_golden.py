"""Generate golden fixtures by running the UNMODIFIED reference (needs /root/reference).

    python tests/golden/make_golden.py

Writes small ``.npz`` files next to this script.  Every solver loop of the reference is run
under ``torch.manual_seed(seed)``; the noise it consumed is regenerated with the recipe of
SURVEY.md 8c (``randn(N, B)`` per draw from the same seed) and stored as ``noise``
``[T][K][N][B]`` so that the oracle and the CUDA engine can replay it.  The committed
fixtures were produced in the build container (torch 2.11.0+cu128, CPU).
"""

import io
import os
import sys
import contextlib

import numpy as np
import torch

REF = os.environ.get("CCVM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from ccvm_simulators.solvers import (  # noqa: E402
    DLSolver, MFSolver, LangevinSolver, PumpedLangevinSolver)
from ccvm_simulators.solvers.algorithms import AdamParameters  # noqa: E402
from ccvm_simulators.problem_classes.boxqp import ProblemInstance  # noqa: E402
from ccvm_simulators.post_processor.factory import PostProcessorFactory  # noqa: E402
from ccvm_simulators.solution import Solution  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
INST20 = f"{REF}/ccvm_simulators/tests/data/test_instances/test020-100-10.in"
INST70 = f"{REF}/examples/benchmarking_instances/Size70/tuningH070-100-1.in"


def replay_noise(seed, t, k, n, b):
    # one randn(N, B) call per reference draw (a single big randn is NOT the same stream
    # on CPU: the vectorised normal fill treats the tail of each call specially)
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randn(n, b, generator=g) for _ in range(t * k)]).reshape(t, k, n, b)


def load(path, solver, bounds=(0.0, 1.0)):
    inst = ProblemInstance(instance_type="test", file_path=path, device="cpu",
                           solution_bounds=bounds)
    inst.scale_coefs(solver.get_scaling_factor(inst.q_matrix))
    return inst


def bind(solver, inst):
    solver.q_matrix, solver.v_vector = inst.q_matrix, inst.v_vector
    solver.solution_bounds = inst.solution_bounds


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: tuple(np.shape(v)) for k, v in out.items()})


HYPERS = {
    "a": dict(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=False),
    "b": dict(alpha=0.05, beta1=0.8, beta2=1.0, add_assign=True),
    "c": dict(alpha=0.01, beta1=0.9, beta2=0.99, add_assign=True),
    "d": dict(alpha=0.02, beta1=0.7, beta2=1.0, add_assign=False),
}


def loops():
    seed = 7
    cases = []
    # (tag, instance path, B, T, bounds, flag)
    shapes = [("n20", INST20, 6, 40, (0.0, 1.0), True),
              ("n20b", INST20, 5, 25, (-0.5, 1.5), False),
              ("n70", INST70, 4, 30, (0.0, 1.0), True)]
    for tag, path, b, t, bounds, flag in shapes:
        # ---- DL
        for pump, s_ctor in ((8.0, 1.0), (0.9, 1.5)):
            sol = DLSolver(device="cpu", batch_size=b, S=s_ctor)
            inst = load(path, sol, bounds)
            bind(sol, inst)
            n = inst.problem_size
            par = dict(pump=pump, dt=0.001, noise_ratio=10.0, feedback_scale=100.0 if pump > 1 else 20.0,
                       g=0.05)
            torch.manual_seed(seed)
            c, s = sol._solve(n, b, "cpu", s_ctor, par["pump"], par["dt"], t, par["noise_ratio"],
                              par["feedback_scale"], flag, par["g"], None, None)
            save(f"dl_{tag}_p{pump}", q=inst.q_matrix, v=inst.v_vector, noise=replay_noise(seed, t, 2, n, b),
                 out_c=c, out_s=s, batch=b, iterations=t, bounds=bounds, flag=flag, s_ctor=s_ctor, **par)
            for hk, hp in HYPERS.items():
                torch.manual_seed(seed)
                c, s = sol._solve_adam(n, b, "cpu", s_ctor, par["pump"], par["dt"], t, par["noise_ratio"],
                                       flag, par["g"], None, None, AdamParameters(**hp).to_dict())
                save(f"dladam_{tag}_p{pump}_{hk}", q=inst.q_matrix, v=inst.v_vector,
                     noise=replay_noise(seed, t, 2, n, b), out_c=c, out_s=s, batch=b, iterations=t,
                     bounds=bounds, flag=flag, s_ctor=s_ctor, **par, **hp)
        # ---- MF
        sol = MFSolver(device="cpu", batch_size=b)
        inst = load(path, sol, bounds)
        bind(sol, inst)
        n = inst.problem_size
        par = dict(pump=0.5, dt=0.0025, j=5.0, feedback_scale=4000.0, S=20.0, g=0.01)
        torch.manual_seed(seed)
        mu, mt, sg = sol._solve(n, b, "cpu", par["S"], par["pump"], par["dt"], t, par["j"],
                                par["feedback_scale"], flag, par["g"], None, None)
        save(f"mf_{tag}", q=inst.q_matrix, v=inst.v_vector, noise=replay_noise(seed, t, 1, n, b),
             out_mu=mu, out_mu_tilde=mt, out_sigma=sg, batch=b, iterations=t, bounds=bounds, flag=flag, **par)
        for hk, hp in HYPERS.items():
            torch.manual_seed(seed)
            mu, mt, sg = sol._solve_adam(n, b, "cpu", par["S"], par["pump"], par["dt"], t, par["j"],
                                         par["feedback_scale"], flag, par["g"], None, None,
                                         AdamParameters(**hp).to_dict())
            save(f"mfadam_{tag}_{hk}", q=inst.q_matrix, v=inst.v_vector, noise=replay_noise(seed, t, 1, n, b),
                 out_mu=mu, out_mu_tilde=mt, out_sigma=sg, batch=b, iterations=t, bounds=bounds, flag=flag,
                 **par, **hp)
        # ---- Langevin
        sol = LangevinSolver(device="cpu", batch_size=b)
        inst = load(path, sol, bounds)
        bind(sol, inst)
        par = dict(dt=0.002, S=0.5, sigma=0.5, feedback_scale=1.0)
        torch.manual_seed(seed)
        c = sol._solve(n, b, "cpu", par["S"], par["dt"], t, par["sigma"], par["feedback_scale"], None, None)
        save(f"lv_{tag}", q=inst.q_matrix, v=inst.v_vector, noise=replay_noise(seed, t, 1, n, b),
             out_c=c, batch=b, iterations=t, bounds=bounds, **par)
        for hk, hp in HYPERS.items():
            torch.manual_seed(seed)
            c = sol._solve_adam(n, b, "cpu", par["S"], par["dt"], t, par["sigma"], par["feedback_scale"],
                                None, None, AdamParameters(**hp).to_dict())
            save(f"lvadam_{tag}_{hk}", q=inst.q_matrix, v=inst.v_vector, noise=replay_noise(seed, t, 1, n, b),
                 out_c=c, batch=b, iterations=t, bounds=bounds, **par, **hp)
        # ---- pumped Langevin
        sol = PumpedLangevinSolver(device="cpu", batch_size=b)
        inst = load(path, sol, bounds)
        bind(sol, inst)
        par = dict(pump=2.0, dt=0.002, S=0.5, sigma=0.5, feedback_scale=1.0)
        torch.manual_seed(seed)
        c = sol._solve(n, b, "cpu", par["S"], par["pump"], par["dt"], t, par["sigma"], flag,
                       par["feedback_scale"], None, None)
        save(f"plv_{tag}", q=inst.q_matrix, v=inst.v_vector, noise=replay_noise(seed, t, 1, n, b),
             out_c=c, batch=b, iterations=t, bounds=bounds, flag=flag, **par)
        for hk, hp in HYPERS.items():
            torch.manual_seed(seed)
            c = sol._solve_adam(n, b, "cpu", par["S"], par["pump"], par["dt"], t, par["sigma"], flag,
                                par["feedback_scale"], None, None, AdamParameters(**hp).to_dict())
            save(f"plvadam_{tag}_{hk}", q=inst.q_matrix, v=inst.v_vector, noise=replay_noise(seed, t, 1, n, b),
                 out_c=c, batch=b, iterations=t, bounds=bounds, flag=flag, **par, **hp)
    return cases


def tensor_s_case():
    """MF with a per-variable saturation vector S (mf_solver.py:833-839)."""
    seed, b, t = 11, 5, 30
    sol = MFSolver(device="cpu", batch_size=b)
    inst = load(INST20, sol)
    bind(sol, inst)
    n = inst.problem_size
    s_vec = torch.linspace(10.0, 30.0, n)
    s2d = torch.outer(torch.ones(b), s_vec)
    torch.manual_seed(seed)
    mu, mt, sg = sol._solve(n, b, "cpu", s2d, 0.0, 0.0025, t, 5.0, 4000.0, True, 0.01, None, None)
    save("mf_tensorS", q=inst.q_matrix, v=inst.v_vector, noise=replay_noise(seed, t, 1, n, b), s_vec=s_vec,
         out_mu=mu, out_mu_tilde=mt, out_sigma=sg, batch=b, iterations=t, pump=0.0, dt=0.0025, j=5.0,
         feedback_scale=4000.0, g=0.01)


def postproc_energy_stats():
    sol = MFSolver(device="cpu", batch_size=16)
    inst = load(INST20, sol)
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(16, inst.problem_size, generator=g)
    with contextlib.redirect_stderr(io.StringIO()):
        x_gd = PostProcessorFactory.create_postprocessor("grad-descent").postprocess(
            x0.clone(), inst.q_matrix, inst.v_vector)
        x_gd5 = PostProcessorFactory.create_postprocessor("grad-descent").postprocess(
            x0.clone(), inst.q_matrix, inst.v_vector, num_iter_pp=5, step_size=0.05,
            lower_clamp=0.1, upper_clamp=0.9)
        x_adam = PostProcessorFactory.create_postprocessor("adam").postprocess(
            x0.clone(), inst.q_matrix, inst.v_vector)
    e0 = inst.compute_energy(x0)
    e_gd = inst.compute_energy(x_gd)
    fake_obj = -torch.tensor([inst.optimal_sol * f for f in
                              (1.0, 0.9995, 0.995, 0.985, 0.975, 0.965, 0.955, 0.92, 0.85, 0.5)])
    solu = Solution(problem_size=20, batch_size=10, instance_name="x", iterations=1, objective_values=fake_obj,
                    solve_time=0.0, pp_time=0.0, optimal_value=inst.optimal_sol, best_value=inst.best_sol,
                    num_frac_values=0, solution_vector=[], variables={"problem_variables": x0})
    perf = solu.solution_performance
    save("postproc_n20", q=inst.q_matrix, v=inst.v_vector, scaled_by=inst.scaled_by, x0=x0, x_gd=x_gd,
         x_gd5=x_gd5, x_adam=x_adam, e0=e0, e_gd=e_gd, optimal=inst.optimal_sol, fake_obj=fake_obj,
         best=solu.best_objective_value, perf=np.array([perf[k] for k in
             ("optimal", "one_percent", "two_percent", "three_percent", "four_percent", "five_percent",
              "ten_percent")]))


def full_calls():
    """Solver.__call__ end to end (epilogue conventions)."""
    seed, b, t = 5, 8, 60
    with contextlib.redirect_stderr(io.StringIO()):
        for pp in (None, "grad-descent", "adam"):
            tag = {None: "none", "grad-descent": "gd", "adam": "adam"}[pp]
            sol = DLSolver(device="cpu", batch_size=b)
            sol.parameter_key = {20: dict(pump=8.0, dt=0.001, iterations=t, noise_ratio=10, feedback_scale=100)}
            inst = load(INST20, sol)
            torch.manual_seed(seed)
            r = sol(instance=inst, post_processor=pp)
            save(f"call_dl_{tag}", q=inst.q_matrix, v=inst.v_vector, scaled_by=inst.scaled_by,
                 noise=replay_noise(seed, t, 2, 20, b), pv=r.variables["problem_variables"], s=r.variables["s"],
                 obj=r.objective_values, best=r.best_objective_value, optimal=inst.optimal_sol,
                 perf=np.array(list(r.solution_performance.values())), batch=b, iterations=t)

            sol = MFSolver(device="cpu", batch_size=b)
            sol.parameter_key = {20: dict(pump=0.0, feedback_scale=4000, j=5.0, S=20.0, dt=0.0025, iterations=t)}
            inst = load(INST20, sol)
            torch.manual_seed(seed)
            r = sol(instance=inst, post_processor=pp,
                    algorithm_parameters=AdamParameters(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=True))
            save(f"call_mfadam_{tag}", q=inst.q_matrix, v=inst.v_vector, scaled_by=inst.scaled_by,
                 noise=replay_noise(seed, t, 1, 20, b), pv=r.variables["problem_variables"],
                 mu=r.variables["mu"], sigma=r.variables["sigma"], obj=r.objective_values,
                 best=r.best_objective_value, optimal=inst.optimal_sol,
                 perf=np.array(list(r.solution_performance.values())), batch=b, iterations=t)

            sol = LangevinSolver(device="cpu", batch_size=b)
            sol.parameter_key = {20: dict(dt=0.002, S=0.5, iterations=t, sigma=0.5, feedback_scale=1.0)}
            inst = load(INST20, sol)
            torch.manual_seed(seed)
            r = sol(instance=inst, post_processor=pp)
            save(f"call_lv_{tag}", q=inst.q_matrix, v=inst.v_vector, scaled_by=inst.scaled_by,
                 noise=replay_noise(seed, t, 1, 20, b), pv=r.variables["problem_variables"],
                 obj=r.objective_values, best=r.best_objective_value, optimal=inst.optimal_sol,
                 perf=np.array(list(r.solution_performance.values())), batch=b, iterations=t)

            sol = PumpedLangevinSolver(device="cpu", batch_size=b)
            sol.parameter_key = {20: dict(pump=2.0, dt=0.002, S=0.5, iterations=t, sigma=0.5, feedback_scale=1.0)}
            inst = load(INST20, sol)
            torch.manual_seed(seed)
            r = sol(instance=inst, post_processor=pp)
            save(f"call_plv_{tag}", q=inst.q_matrix, v=inst.v_vector, scaled_by=inst.scaled_by,
                 noise=replay_noise(seed, t, 1, 20, b), pv=r.variables["problem_variables"],
                 obj=r.objective_values, best=r.best_objective_value, optimal=inst.optimal_sol,
                 perf=np.array(list(r.solution_performance.values())), batch=b, iterations=t)


def instance_file():
    """A synthetic instance file in the reference's on-disk format plus what the reference's
    own loader makes of it (problem_instance.py:116-224)."""
    n = 7
    g = torch.Generator().manual_seed(21)
    a = torch.randn(n, n, generator=g)
    qf = ((a + a.T) * 5).round()
    vf = (torch.randn(n, generator=g) * 20).round()
    sol_vec = torch.rand(n, generator=g).round()
    path = os.path.join(HERE, "synthetic007.in")
    with open(path, "w") as fh:
        fh.write("\t".join(["7", "123.456789", "120.5", "True", "0.0123", "0.0045", "42", "2"]) + "\n")
        fh.write("\t".join(str(float(x)) for x in vf) + "\n")
        for r in range(n):
            fh.write("\t".join(str(float(x)) for x in qf[r]) + "\n")
        fh.write("\t".join(str(float(x)) for x in sol_vec) + "\t\n")
    inst = ProblemInstance(instance_type="test", file_path=path, device="cpu")
    sol = DLSolver(device="cpu")
    f = sol.get_scaling_factor(inst.q_matrix)
    q0, v0 = inst.q_matrix.clone(), inst.v_vector.clone()
    inst.scale_coefs(f)
    save("instance007", q=q0, v=v0, q_scaled=inst.q_matrix, v_scaled=inst.v_vector, factor=f,
         scaled_by=inst.scaled_by, optimal=inst.optimal_sol, best=inst.best_sol,
         num_frac=inst.num_frac_values, sol_time_gb=inst.sol_time_gb, sol_time_bfgs=inst.sol_time_bfgs,
         solution_vector=np.array(inst.solution_vector))


if __name__ == "__main__":
    torch.set_num_threads(1)
    loops()
    tensor_s_case()
    postproc_energy_stats()
    full_calls()
    instance_file()
