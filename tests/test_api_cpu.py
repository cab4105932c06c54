"""Host-side logic that needs no GPU: argument validation with the reference's messages, the
C-ABI library loads and exports every symbol the header declares, and compute calls fail loudly
(no CPU fallback) when CUDA is absent."""
import os
import re

import pytest
import torch

import ccvm_b200
from ccvm_b200 import _native as nat
from ccvm_b200.solvers import (DLSolver, MFSolver, LangevinSolver, PumpedLangevinSolver, AdamParameters,
                               CCVMSolver, MachineType)
from ccvm_b200.problem_classes.boxqp import ProblemInstance
from ccvm_b200.post_processor import PostProcessorFactory, PostProcessorAdam, PostProcessorGradDescent

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOLVERS = (DLSolver, MFSolver, LangevinSolver, PumpedLangevinSolver)
KEYS = {
    DLSolver: dict(pump=2.0, dt=0.005, iterations=100, noise_ratio=10, feedback_scale=100),
    MFSolver: dict(pump=0.0, feedback_scale=4000, j=5.0, S=20.0, dt=0.0025, iterations=100),
    LangevinSolver: dict(dt=0.002, S=0.5, iterations=100, sigma=0.5, feedback_scale=1.0),
    PumpedLangevinSolver: dict(pump=2.0, dt=0.002, S=0.5, iterations=100, sigma=0.5, feedback_scale=1.0),
}


def test_library_exports_every_declared_symbol():
    lib = nat.load()
    header = open(os.path.join(ROOT, "include", "ccvm_b200.h")).read()
    declared = set(re.findall(r"\b(ccvm_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ccvm_b200.h but not exported"
    assert set(nat.EXPORTS) == declared
    assert lib.ccvm_abi_version() == nat.ABI_VERSION == int(re.search(r"#define CCVM_ABI_VERSION (\d+)", header).group(1))


@pytest.mark.parametrize("cls", SOLVERS)
def test_constructor_device_validation(cls):
    with pytest.raises(ValueError, match="Given device is not available"):
        cls(device="tpu")
    s = cls(device="cpu")
    assert s.device == "cpu" and s.batch_size == 1000 and not s.is_tuned
    with pytest.raises(ValueError, match="not a valid problem category"):
        cls(device="cpu", problem_category="maxcut")


@pytest.mark.parametrize("cls", SOLVERS)
def test_parameter_key_exact_key_set(cls):
    s = cls(device="cpu")
    good = {20: dict(KEYS[cls])}
    s.parameter_key = good
    assert s.parameter_key == good
    bad = {20: {k: v for k, v in list(KEYS[cls].items())[:-1]}}
    with pytest.raises(ValueError, match="The parameter key is not valid for this solver"):
        s.parameter_key = bad
    extra = {20: dict(KEYS[cls], bogus=1)}
    with pytest.raises(ValueError):
        s.parameter_key = extra


def test_scaling_multipliers_and_hook_binding():
    assert DLSolver("cpu")._scaling_multiplier == 0.2
    for cls in (MFSolver, LangevinSolver, PumpedLangevinSolver):
        assert cls("cpu")._scaling_multiplier == 0.05
    s = MFSolver("cpu")
    assert s.calculate_grads == s._calculate_grads_boxqp and s.change_variables == s._change_variables_boxqp
    assert s.calculate_drift == s._calculate_drift_boxqp and s.fit_to_constraints == s._fit_to_constraints_boxqp


def test_adam_parameters_validation():
    p = AdamParameters(alpha=0.001, beta1=0.9, beta2=0.999, add_assign=False)
    assert p.to_dict() == {"alpha": 0.001, "beta1": 0.9, "beta2": 0.999, "add_assign": False}
    assert AdamParameters(beta2=1.0).beta2 == 1.0
    for kw in (dict(alpha=-1), dict(beta1=0), dict(beta1=1), dict(beta2=0), dict(beta2=1.5)):
        with pytest.raises(ValueError, match="AdamAlgorithm: Invalid"):
            AdamParameters(**kw)


def test_device_mismatch_and_missing_size_messages(golden_dir):
    inst = ProblemInstance(instance_type="test", file_path=os.path.join(golden_dir, "synthetic007.in"), device="cpu")
    s = MFSolver(device="cuda", batch_size=4)
    s.parameter_key = {7: dict(KEYS[MFSolver])}
    with pytest.raises(ValueError, match=r"The device type of the instance \(cpu\) and the solver \(cuda\) must match."):
        s(instance=inst)
    s2 = MFSolver(device="cpu", batch_size=4)
    s2.parameter_key = {20: dict(KEYS[MFSolver])}
    with pytest.raises(KeyError, match="for the given instance size is not defined"):
        s2(instance=inst)


def test_no_cpu_fallback(golden_dir):
    """A cpu solver validates but refuses to solve: the package has no CPU implementation."""
    inst = ProblemInstance(instance_type="test", file_path=os.path.join(golden_dir, "synthetic007.in"), device="cpu")
    s = LangevinSolver(device="cpu", batch_size=4)
    s.parameter_key = {7: dict(KEYS[LangevinSolver])}
    with pytest.raises(nat.NativeError):
        s(instance=inst)
    if not torch.cuda.is_available():
        with pytest.raises(nat.NativeError):
            inst.compute_energy(torch.zeros(2, 7))


def test_instance_loader_host_side(golden_dir):
    import numpy as np
    z = np.load(os.path.join(golden_dir, "instance007.npz"))
    inst = ProblemInstance(instance_type="test", file_path=os.path.join(golden_dir, "synthetic007.in"), device="cpu")
    assert inst.problem_size == 7 and inst.name == "synthetic007" and inst.optimality is True
    assert inst.optimal_sol == float(z["optimal"]) and inst.best_sol == float(z["best"])
    assert inst.num_frac_values == int(z["num_frac"]) and inst.scaled_by == 1
    assert inst.sol_time_gb == float(z["sol_time_gb"]) and inst.sol_time_bfgs == float(z["sol_time_bfgs"])
    assert np.array_equal(inst.q_matrix.numpy(), z["q"]) and np.array_equal(inst.v_vector.numpy(), z["v"])
    assert inst.solution_vector == list(z["solution_vector"])
    with pytest.raises(ValueError, match="instance_type must be tuning or test"):
        ProblemInstance(instance_type="nope")
    with pytest.raises(ValueError, match="Minimum solution bound must be less than maximum"):
        ProblemInstance(solution_bounds=(1.0, 0.0))
    with pytest.raises(ValueError, match="tuple of size 2"):
        ProblemInstance(solution_bounds=(1.0,))
    with pytest.raises(Exception, match="Error reading instance file"):
        ProblemInstance(file_path=os.path.join(golden_dir, "make_golden.py"))


def test_post_processor_factory_and_type_errors():
    assert isinstance(PostProcessorFactory.create_postprocessor("adam"), PostProcessorAdam)
    assert isinstance(PostProcessorFactory.create_postprocessor("Grad-Descent"), PostProcessorGradDescent)
    with pytest.raises(AssertionError, match="Method type is not valid"):
        PostProcessorFactory.create_postprocessor("nope")
    with pytest.raises(NotImplementedError):
        PostProcessorFactory.create_postprocessor("lbfgs")
    q, v = torch.zeros(3, 3), torch.zeros(3)
    for pp in (PostProcessorAdam(), PostProcessorGradDescent()):
        with pytest.raises(TypeError, match="parameter c must be a tensor"):
            pp.postprocess([[0.0] * 3], q, v)
        with pytest.raises(TypeError, match="parameter q_matrix must be a tensor"):
            pp.postprocess(torch.zeros(2, 3), "q", v)
        with pytest.raises(TypeError, match="parameter v_vector must be a tensor"):
            pp.postprocess(torch.zeros(2, 3), q, None)
        with pytest.raises(Exception):
            pp.postprocess(torch.zeros(2, 4), q, v)


def test_machine_time_energy_models():
    """Closed-form bookkeeping (reference test_ccvm_solver.py:372-494 style)."""
    import pandas as pd
    df = pd.DataFrame({"solve_time": [100.0, 120.0], "pp_time": [1.0, 3.0], "iterations": [1000, 1000]})
    s = DLSolver("cpu")
    s.parameter_key = {20: dict(KEYS[DLSolver])}
    assert s.machine_time("cpu")(dataframe=df, problem_size=20) == 110.0
    assert s.machine_time("gpu")(dataframe=df, problem_size=20) == 110.0
    assert s.machine_energy("cpu")(df, 20) == pytest.approx(4.93 * 110.0)
    assert s.machine_energy("gpu")(df, 20) == pytest.approx(28.93 * 110.0)
    assert s.machine_time("dl-ccvm")(df, 20) == pytest.approx(20 * 10e-12 * 1000 + 2.0)
    assert s.machine_energy("dl-ccvm")(df, 20) > 4.96 * 2.0
    with pytest.raises(ValueError, match="Mismatch between the solver and the machine type"):
        s.machine_time("fpga")
    with pytest.raises(ValueError, match="The given machine type is not valid"):
        s.machine_energy("abacus")
    lv = LangevinSolver("cpu")
    assert lv.machine_time("fpga")(df, 20) == pytest.approx(133e-6 + 2.0)
    assert lv.machine_energy("fpga")(df, 20) == pytest.approx(17.18 * 133e-6)
    mf = MFSolver("cpu")
    mf.parameter_key = {20: dict(KEYS[MFSolver])}
    rt = (34 + 0.1 * 20) * 3.33e-9 + 20 * 100e-12 + 3.33e-9
    assert mf.machine_time("mf-ccvm")(df, 20) == pytest.approx(rt * 1000 + 2.0)
    assert mf.machine_energy("mf-ccvm")(df, 20) == pytest.approx(
        (rt * (15.74 + 1000e-6 * (0.0 + 1 + 5.0)) - 15.74 * 3.33e-9) * 1000 + 4.87 * 2.0)
    assert {m.value for m in MachineType} == {"cpu", "gpu", "fpga", "dl-ccvm", "mf-ccvm"}


def test_overridden_hooks_are_rejected():
    s = MFSolver("cuda", batch_size=2)
    s.calculate_grads = lambda *a, **k: None
    with pytest.raises(RuntimeError, match="was replaced"):
        s._require_stock_hooks()
