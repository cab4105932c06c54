"""LangevinSolver -- plain Langevin dynamics.  API of the reference's
``solvers/langevin_solver.py``; the loops (_solve 368-435, _solve_adam 437-561) run as one
persistent sm_100a kernel."""
import numpy as np
import torch

from .. import engine
from .._native import SOLVER_LANGEVIN, ALG_ORIGINAL, ALG_ADAM
from .ccvm_solver import CCVMSolver

LANGEVIN_SCALING_MULTIPLIER = 0.05
"""Multiplier used by LangevinSolver in get_scaling_factor()."""


def unit_box_map(S):
    """(scale, shift) of x = (c + S) / (2 S) -- Langevin and PumpedLangevin ignore the instance's
    solution bounds here (reference langevin_solver.py:717-722)."""
    if torch.is_tensor(S) and S.numel() > 1:
        vec = (S if S.ndim == 1 else S[0]).detach().double().cpu().numpy()
        return torch.from_numpy((0.5 / vec).astype(np.float32)), 0.5
    return 0.5 / float(S), 0.5


class LangevinSolver(CCVMSolver):
    """Langevin dynamics as a system of SDEs.

    Args:
        device (str): "cuda" to solve ("cpu" is accepted for construction/validation only).
        problem_category (str): "boxqp".
        batch_size (int): trajectories per solve.  Default 1000.
    """

    _PARAMETER_KEYS = frozenset(["dt", "S", "iterations", "sigma", "feedback_scale"])

    def __init__(self, device, problem_category="boxqp", batch_size=1000):
        super().__init__(device)
        self.batch_size = batch_size
        self._scaling_multiplier = LANGEVIN_SCALING_MULTIPLIER
        self._method_selector(problem_category)
        self._default_fpga_machine_parameters = {
            "fpga_power": {20: 17.18, 30: 18.13, 40: 18.45, 50: 19.03, 60: 19.22, 70: 19.32},
            "fpga_runtimes": {20: 133e-6, 30: 265e-6, 40: 327e-6, 50: 437e-6, 60: 511e-6, 70: 662e-6},
        }

    # --------------------------------------------------------------------- hooks
    def _calculate_drift_boxqp(self, c, lower_limit=0, upper_limit=1, S=1):
        """-((c a/(2S) + b/2) Q + V) a/(2S) (reference langevin_solver.py:117-139)."""
        return engine.eval_hook(SOLVER_LANGEVIN, "drift", self.q_matrix, self.v_vector, (c,), lower_limit,
                                upper_limit, S)[0]

    def _calculate_grads_boxqp(self, c, lower_limit=0, upper_limit=1, S=1):
        """Same expression as the drift (reference 141-166)."""
        return engine.eval_hook(SOLVER_LANGEVIN, "grads", self.q_matrix, self.v_vector, (c,), lower_limit,
                                upper_limit, S)[0]

    def _append_samples_to_file(self, c_sample, evolution_file_object, s_sample=None):
        """c rows, tab after every value.  (The reference's signature also demands an s_sample
        that its own caller never passes -> TypeError; it is optional here.)"""
        self._append_rows(c_sample, evolution_file_object)
        if s_sample is not None:
            self._append_rows(s_sample, evolution_file_object)

    # ------------------------------------------------------------ machine models
    def _validate_fpga_machine_parameters(self, machine_parameters):
        missing_keys = [key for key in ("fpga_power", "fpga_runtimes") if key not in machine_parameters]
        if missing_keys:
            raise ValueError(f"Invalid fpga_machine_parameters: Missing required keys - {missing_keys}")

    def tune(self, instances, post_processor=None, pump_rate_flag=True, g=0.05):
        """Placeholder, as in the reference."""
        self._is_tuned = True

    def _fpga_machine_energy(self, machine_parameters=None):
        """fpga_power[N] * fpga_runtimes[N] (reference 250-293)."""
        if machine_parameters is None:
            machine_parameters = self._default_fpga_machine_parameters
        else:
            self._validate_fpga_machine_parameters(machine_parameters)

        def _fpga_machine_energy_callable(matching_df, problem_size):
            return machine_parameters["fpga_power"][problem_size] * machine_parameters["fpga_runtimes"][problem_size]

        return _fpga_machine_energy_callable

    def _fpga_machine_time(self, machine_parameters=None):
        """fpga_runtimes[N] + mean pp_time (reference 295-366)."""
        if machine_parameters is None:
            machine_parameters = self._default_fpga_machine_parameters
        else:
            self._validate_fpga_machine_parameters(machine_parameters)

        def _fpga_machine_time_callable(dataframe, problem_size):
            try:
                postprocessing_time = np.mean(dataframe["pp_time"].values)
            except KeyError as e:
                raise ValueError(f"The given dataframe is missing required column: {e.args[0]}")
            try:
                return machine_parameters["fpga_runtimes"][problem_size] + postprocessing_time
            except KeyError:
                raise ValueError(
                    f"The fpga_runtimes dict in given machine_parameters does not have an entry for problem size {problem_size}."
                )

        return _fpga_machine_time_callable

    # --------------------------------------------------------------------- loops
    def _solve(self, problem_size, batch_size, device, S, dt, iterations, sigma, feedback_scale,
               evolution_step_size, samples_taken):
        """Original Langevin loop -> c (clamped to [-S, S] every iteration)."""
        (c,) = self._engine_solve(SOLVER_LANGEVIN, ALG_ORIGINAL, batch_size, iterations, S, evolution_step_size,
                                  dt=dt, sigma=sigma, feedback_scale=feedback_scale)
        self._publish_samples(("c_sample",))
        return c

    def _solve_adam(self, problem_size, batch_size, device, S, dt, iterations, sigma, feedback_scale,
                    evolution_step_size, samples_taken, hyperparameters):
        """Langevin loop with Adam on the gradient -> c."""
        (c,) = self._engine_solve(SOLVER_LANGEVIN, ALG_ADAM, batch_size, iterations, S, evolution_step_size,
                                  hyperparameters, dt=dt, sigma=sigma, feedback_scale=feedback_scale)
        self._publish_samples(("c_sample",))
        return c

    def __call__(self, instance, post_processor=None, evolution_step_size=None, evolution_file=None,
                 algorithm_parameters=None):
        """Solve ``instance``; returns a Solution whose problem_variables are (c + S) / (2S),
        post-processed if asked (reference langevin_solver.py:563-762)."""
        self._check_device(instance)
        problem_size = instance.problem_size
        self._bind_instance(instance)
        dt, S, iterations, sigma, feedback_scale = self._read_parameters(
            problem_size, ("dt", "S", "iterations", "sigma", "feedback_scale"))
        S = self._normalise_s(S, problem_size)

        def solve_args(adam):
            return (problem_size, self.batch_size, self.device, S, dt, iterations, sigma, feedback_scale,
                    evolution_step_size, 0 if evolution_step_size else None)

        def finish(c):
            return c, unit_box_map(S), None, lambda pv: {"problem_variables": pv}

        return self._run(SOLVER_LANGEVIN, instance, post_processor, evolution_step_size, evolution_file,
                         algorithm_parameters, iterations, S, solve_args, finish)
