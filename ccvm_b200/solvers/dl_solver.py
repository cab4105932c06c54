"""DLSolver -- delay-line CCVM (two quadratures c, s).  API of the reference's
``solvers/dl_solver.py``; the loops (_solve 468-569, _solve_adam 571-769) run as one
persistent sm_100a kernel."""
import numpy as np
import torch

from .. import engine
from .._native import SOLVER_DL, ALG_ORIGINAL, ALG_ADAM
from .ccvm_solver import CCVMSolver

DL_SCALING_MULTIPLIER = 0.2
"""Multiplier used by DLSolver in get_scaling_factor()."""

_OPTICS_KEYS = ("laser_power", "modulators_power", "squeezing_power", "electronics_power",
                "amplifiers_power", "electronics_latency", "laser_clock", "postprocessing_power")


class DLSolver(CCVMSolver):
    """Delay-line coherent continuous-variable machine.

    Args:
        device (str): "cuda" to solve ("cpu" is accepted for construction/validation only).
        problem_category (str): "boxqp".
        batch_size (int): trajectories per solve.  Default 1000.
        S (float or torch.Tensor): enforced saturation value.  Default 1.
    """

    _PARAMETER_KEYS = frozenset(["pump", "dt", "iterations", "noise_ratio", "feedback_scale"])

    def __init__(self, device, problem_category="boxqp", batch_size=1000, S=1):
        super().__init__(device)
        self.batch_size = batch_size
        self.S = S
        self._default_optics_machine_parameters = {
            "laser_power": 1200e-6,
            "modulators_power": 10e-3,
            "squeezing_power": 180e-3,
            "electronics_power": 0.0,
            "amplifiers_power": 222.2e-3,
            "electronics_latency": 1e-9,
            "laser_clock": 10e-12,
            "postprocessing_power": {20: 4.96, 30: 5.1, 40: 4.95, 50: 5.26, 60: 5.11, 70: 5.09},
        }
        self._scaling_multiplier = DL_SCALING_MULTIPLIER
        self._method_selector(problem_category)

    # --------------------------------------------------------------------- hooks
    def _calculate_drift_boxqp(self, c, s, pump, rate, feedback_scale=100, lower_limit=0, upper_limit=1, S=1):
        """(c_drift, s_drift) of the DL SDE (reference dl_solver.py:117-172); inside the drift S
        becomes sqrt(pump-1) when pump > 1."""
        return engine.eval_hook(SOLVER_DL, "drift", self.q_matrix, self.v_vector, (c, s), lower_limit,
                                upper_limit, S, pump=pump, rate=rate, feedback_scale=feedback_scale)

    def _calculate_grads_boxqp(self, c, s, lower_limit=0, upper_limit=1, S=1):
        """(c_grads, s_grads) = -(1/4 ((y a/S + b)Q) a/S + V a/(2S)) (reference 174-217)."""
        return engine.eval_hook(SOLVER_DL, "grads", self.q_matrix, self.v_vector, (c, s), lower_limit,
                                upper_limit, S)

    def _append_samples_to_file(self, c_sample, s_sample, evolution_file_object):
        """c rows then s rows, tab after every value (reference 252-281)."""
        self._append_rows(c_sample, evolution_file_object)
        self._append_rows(s_sample, evolution_file_object)

    # ------------------------------------------------------------ machine models
    def _is_valid_optics_machine_parameters(self, machine_parameters):
        missing_keys = [key for key in _OPTICS_KEYS if key not in machine_parameters]
        if missing_keys:
            raise ValueError(f"Invalid optics_machine_parameters: Missing required keys - {missing_keys}")

    def tune(self, instances, post_processor=None, pump_rate_flag=True, g=0.05):
        """Placeholder, as in the reference (which raises AttributeError here);
        this one just records the flag."""
        self._is_tuned = True

    def _optics_machine_energy(self, machine_parameters=None):
        """Energy model of the optical DL-CCVM (reference 331-406)."""
        if machine_parameters is None:
            machine_parameters = self._default_optics_machine_parameters
        else:
            self._is_valid_optics_machine_parameters(machine_parameters)

        def _optics_machine_energy_callable(dataframe, problem_size):
            self._validate_machine_energy_dataframe_columns(dataframe)
            try:
                pump = self.parameter_key[problem_size]["pump"]
            except KeyError:
                raise KeyError(f"Pump for the given instance size: {problem_size} is not defined.")
            mp, n = machine_parameters, float(problem_size)
            t_clock, t_elec = mp["laser_clock"], mp["electronics_latency"]
            per_iteration = (
                pump * mp["laser_power"] * (t_elec + t_clock * n)
                + 2 * mp["modulators_power"] * t_clock * n * (n - 1)
                + mp["squeezing_power"] * (t_elec + t_clock * n)
                + mp["electronics_power"] * (t_elec + t_clock * n)
                + mp["amplifiers_power"] * (t_elec * (n - 1) + t_clock * n * (n - 1))
            )
            optics_energy = per_iteration * np.mean(dataframe["iterations"].values)
            pp_energy = mp["postprocessing_power"][problem_size] * np.mean(dataframe["pp_time"].values)
            return optics_energy + pp_energy

        return _optics_machine_energy_callable

    def _optics_machine_time(self, machine_parameters=None):
        """N * laser_clock * iterations + pp_time (reference 408-466)."""
        if machine_parameters is None:
            machine_parameters = self._default_optics_machine_parameters
        else:
            self._is_valid_optics_machine_parameters(machine_parameters)

        def _optics_machine_time_callable(dataframe, problem_size):
            try:
                iterations = np.mean(dataframe["iterations"].values)
                postprocessing_time = np.mean(dataframe["pp_time"].values)
            except KeyError as e:
                raise KeyError(
                    f"The given dataframe is missing the {e.args[0]} "
                    f"column. Required columns are: ['iterations', 'pp_time']."
                )
            return float(problem_size) * machine_parameters["laser_clock"] * iterations + postprocessing_time

        return _optics_machine_time_callable

    # --------------------------------------------------------------------- loops
    def _solve(self, problem_size, batch_size, device, S, pump, dt, iterations, noise_ratio,
               feedback_scale, pump_rate_flag, g, evolution_step_size, samples_taken):
        """Original DL-CCVM loop -> (c, s).  Same signature as the reference (468-483)."""
        c, s = self._engine_solve(SOLVER_DL, ALG_ORIGINAL, batch_size, iterations, S, evolution_step_size, pump=pump, dt=dt,
                                  noise_ratio=noise_ratio, feedback_scale=feedback_scale,
                                  pump_rate_flag=pump_rate_flag, g=g)
        self._publish_samples(("c_sample", "s_sample"))
        return c, s

    def _solve_adam(self, problem_size, batch_size, device, S, pump, dt, iterations, noise_ratio,
                    pump_rate_flag, g, evolution_step_size, samples_taken, hyperparameters):
        """DL-CCVM loop with Adam -> (c, s).  Like the reference (571-586) it takes NO
        feedback_scale."""
        c, s = self._engine_solve(SOLVER_DL, ALG_ADAM, batch_size, iterations, S, evolution_step_size, hyperparameters, pump=pump,
                                  dt=dt, noise_ratio=noise_ratio, pump_rate_flag=pump_rate_flag, g=g)
        self._publish_samples(("c_sample", "s_sample"))
        return c, s

    def __call__(self, instance, post_processor=None, pump_rate_flag=True, g=0.05, evolution_step_size=None,
                 evolution_file=None, algorithm_parameters=None):
        """Solve ``instance``; returns a Solution (reference 771-999).

        Output conventions kept from the reference: without a post-processor
        ``variables["problem_variables"]`` holds the raw clamped amplitudes c and the objective is
        evaluated on change_variables(c); with one, the post-processor runs on
        change_variables(c) and the objective is evaluated on change_variables(pp output) (the
        reference applies the map twice, 941-958).

        One defect of the reference is NOT reproduced: its ``__call__`` passes feedback_scale to
        ``_solve_adam``, which does not accept it (TypeError); here the Adam path
        simply works.
        """
        self._check_device(instance)
        problem_size = instance.problem_size
        self._bind_instance(instance)
        pump, dt, iterations, noise_ratio, feedback_scale = self._read_parameters(
            problem_size, ("pump", "dt", "iterations", "noise_ratio", "feedback_scale"))
        S = self._normalise_s(self.S, problem_size)
        lower, upper = self.solution_bounds

        def solve_args(adam):
            head = (problem_size, self.batch_size, self.device, S, pump, dt, iterations, noise_ratio)
            tail = (pump_rate_flag, g, evolution_step_size, 0 if evolution_step_size else None)
            return head + tail if adam else head + (feedback_scale,) + tail

        def finish(outs):
            c, s = outs
            cov = (_cov_scale(S, lower, upper), 0.5 * (upper + lower))
            map1 = cov if post_processor else None

            def variables(pv):
                return {"problem_variables": pv, "s": s}

            return c, map1, cov, variables

        return self._run(SOLVER_DL, instance, post_processor, evolution_step_size, evolution_file,
                         algorithm_parameters, iterations, S, solve_args, finish)


def _cov_scale(S, lower, upper):
    """0.5 * (u - l) / S: scalar, or per-variable tensor when S is one (host-side, N values)."""
    half = 0.5 * (upper - lower)
    if torch.is_tensor(S) and S.numel() > 1:
        vec = (S if S.ndim == 1 else S[0]).detach().double().cpu().numpy()
        return torch.from_numpy((half / vec).astype(np.float32))
    return half / float(S)
