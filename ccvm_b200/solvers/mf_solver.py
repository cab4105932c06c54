"""MFSolver -- measurement-feedback CCVM (state mu, sigma).  API of the reference's
``solvers/mf_solver.py``; the loops (_solve 493-593, _solve_adam 595-764) run as one persistent
sm_100a kernel."""
import numpy as np

from .. import engine
from .._native import SOLVER_MF, ALG_ORIGINAL, ALG_ADAM
from .ccvm_solver import CCVMSolver
from .dl_solver import _cov_scale

MF_SCALING_MULTIPLIER = 0.05
"""Multiplier used by MFSolver in get_scaling_factor()."""

_OPTICS_KEYS = ("laser_clock", "FPGA_clock", "FPGA_fixed", "FPGA_var_fac", "FPGA_power", "buffer_time",
                "laser_power", "postprocessing_power")


class MFSolver(CCVMSolver):
    """Measurement-feedback coherent continuous-variable machine.

    Args:
        device (str): "cuda" to solve ("cpu" is accepted for construction/validation only).
        problem_category (str): "boxqp".
        batch_size (int): trajectories per solve.  Default 1000.
    """

    _PARAMETER_KEYS = frozenset(["pump", "feedback_scale", "j", "S", "dt", "iterations"])
    _EVOLUTION_TRAILING_TAB = False

    def __init__(self, device, problem_category="boxqp", batch_size=1000):
        super().__init__(device)
        self.batch_size = batch_size
        self._scaling_multiplier = MF_SCALING_MULTIPLIER
        self._default_optics_machine_parameters = {
            "laser_clock": 100e-12,
            "FPGA_clock": 3.33e-9,
            "FPGA_fixed": 34,
            "FPGA_var_fac": 0.1,
            "FPGA_power": {20: 15.74, 30: 16.97, 40: 18.54, 50: 20.25, 60: 22.08, 70: 24.01},
            "buffer_time": 3.33e-9,
            "laser_power": 1000e-6,
            "postprocessing_power": {20: 4.87, 30: 5.14, 40: 5.11, 50: 5.08, 60: 5.09, 70: 5.3},
        }
        self._method_selector(problem_category)

    # --------------------------------------------------------------------- hooks
    def _calculate_drift_boxqp(self, mu, mu_tilde, sigma, pump, j, g, S, fs, lower_limit=0, upper_limit=1):
        """(drift_mu, drift_sigma) (reference mf_solver.py:141-198)."""
        return engine.eval_hook(SOLVER_MF, "drift", self.q_matrix, self.v_vector, (mu, mu_tilde, sigma),
                                lower_limit, upper_limit, S, pump=pump, j=j, g=g, feedback_scale=fs)

    def _calculate_grads_boxqp(self, mu_tilde, S, fs, lower_limit=0, upper_limit=1):
        """fs * (-1/4 ((mu_tilde a/S + b)Q) a/S - V a/(2S)) (reference mf_solver.py:200-233)."""
        return engine.eval_hook(SOLVER_MF, "grads", self.q_matrix, self.v_vector, (mu_tilde,), lower_limit,
                                upper_limit, S, feedback_scale=fs)[0]

    def _append_samples_to_file(self, mu_sample, sigma_sample, evolution_file_object):
        """mu rows then sigma rows, tab-separated without trailing tab (reference 267-300)."""
        self._append_rows(mu_sample, evolution_file_object)
        self._append_rows(sigma_sample, evolution_file_object)

    # ------------------------------------------------------------ machine models
    def _is_valid_optics_machine_parameters(self, machine_parameters):
        missing_keys = [key for key in _OPTICS_KEYS if key not in machine_parameters]
        if missing_keys:
            raise ValueError(f"Invalid optics_machine_parameters: Missing required keys - {missing_keys}")

    def tune(self, instances, post_processor=None, g=0.01):
        """Placeholder, as in the reference."""
        self._is_tuned = True

    @staticmethod
    def _roundtrip_time(mp, problem_size):
        """One optical round trip: FPGA cycles, N laser pulses and the buffer."""
        n = float(problem_size)
        return (mp["FPGA_fixed"] + mp["FPGA_var_fac"] * n) * mp["FPGA_clock"] + n * mp["laser_clock"] + mp["buffer_time"]

    def _optics_machine_energy(self, machine_parameters=None):
        """Energy model of the optical MF-CCVM (reference mf_solver.py:342-428): per round trip
        (FPGA power + laser power x (pump + 1 + j)) x round-trip time, minus the FPGA idling
        through the buffer, plus post-processing."""
        if machine_parameters is None:
            machine_parameters = self._default_optics_machine_parameters
        else:
            self._is_valid_optics_machine_parameters(machine_parameters)

        def _optics_machine_energy_callable(dataframe, problem_size):
            self._validate_machine_energy_dataframe_columns(dataframe)
            try:
                pump = self.parameter_key[problem_size]["pump"]
                measure_strength = self.parameter_key[problem_size]["j"]
            except KeyError as e:
                raise KeyError(
                    f"The parameter '{e.args[0]}' for the given instance size: {problem_size} is not defined."
                ) from e
            mp = machine_parameters
            iterations = np.mean(dataframe["iterations"].values)
            postprocessing_time = np.mean(dataframe["pp_time"].values)
            fpga_power = mp["FPGA_power"][problem_size]
            optics_power = fpga_power + mp["laser_power"] * (pump + 1 + measure_strength)
            optics_energy = (self._roundtrip_time(mp, problem_size) * optics_power
                             - fpga_power * mp["buffer_time"]) * iterations
            return optics_energy + mp["postprocessing_power"][problem_size] * postprocessing_time

        return _optics_machine_energy_callable

    def _optics_machine_time(self, machine_parameters=None):
        """round-trip time x iterations + pp_time (reference mf_solver.py:430-491)."""
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
                    f"The given dataframe is missing the {e.args[0]} column. Required columns are: ['iterations', 'pp_time']."
                )
            return self._roundtrip_time(machine_parameters, problem_size) * iterations + postprocessing_time

        return _optics_machine_time_callable

    # --------------------------------------------------------------------- loops
    def _solve(self, problem_size, batch_size, device, S, pump, dt, iterations, j, feedback_scale,
               pump_rate_flag, g, evolution_step_size, samples_taken):
        """Original MF-CCVM loop -> (mu, mu_tilde, sigma); mu_tilde is the clamped measurement of
        the LAST iteration (reference 591)."""
        mu, mu_tilde, sigma = self._engine_solve(
            SOLVER_MF, ALG_ORIGINAL, batch_size, iterations, S, evolution_step_size, pump=pump, dt=dt, j=j,
            feedback_scale=feedback_scale, pump_rate_flag=pump_rate_flag, g=g)
        self._publish_samples(("mu_sample", "sigma_sample"))
        return mu, mu_tilde, sigma

    def _solve_adam(self, problem_size, batch_size, device, S, pump, dt, iterations, j, feedback_scale,
                    pump_rate_flag, g, evolution_step_size, samples_taken, hyperparameters):
        """MF-CCVM loop with Adam on the feedback term -> (mu, mu_tilde, sigma)."""
        mu, mu_tilde, sigma = self._engine_solve(
            SOLVER_MF, ALG_ADAM, batch_size, iterations, S, evolution_step_size, hyperparameters, pump=pump,
            dt=dt, j=j, feedback_scale=feedback_scale, pump_rate_flag=pump_rate_flag, g=g)
        self._publish_samples(("mu_sample", "sigma_sample"))
        return mu, mu_tilde, sigma

    def __call__(self, instance, post_processor=None, g=0.01, pump_rate_flag=True, evolution_step_size=None,
                 evolution_file=None, algorithm_parameters=None):
        """Solve ``instance``; returns a Solution with variables problem_variables / mu / sigma
        (reference mf_solver.py:766-989).  The objective is evaluated on the (post-processed)
        change of variables of the last measurement mu_tilde."""
        self._check_device(instance)
        problem_size = instance.problem_size
        self._bind_instance(instance)
        pump, dt, iterations, j, feedback_scale, S = self._read_parameters(
            problem_size, ("pump", "dt", "iterations", "j", "feedback_scale", "S"))
        S = self._normalise_s(S, problem_size)
        lower, upper = self.solution_bounds

        def solve_args(adam):
            return (problem_size, self.batch_size, self.device, S, pump, dt, iterations, j, feedback_scale,
                    pump_rate_flag, g, evolution_step_size, 0 if evolution_step_size else None)

        def finish(outs):
            mu, mu_tilde, sigma = outs

            def variables(pv):
                return {"problem_variables": pv, "mu": mu, "sigma": sigma}

            return mu_tilde, (_cov_scale(S, lower, upper), 0.5 * (upper + lower)), None, variables

        return self._run(SOLVER_MF, instance, post_processor, evolution_step_size, evolution_file,
                         algorithm_parameters, iterations, S, solve_args, finish)
