"""PumpedLangevinSolver -- Langevin dynamics with a ramped pump term.  API of the reference's
``solvers/pumped_langevin_solver.py``; the loops (_solve 232-309, _solve_adam 311-449) run as
one persistent sm_100a kernel."""
from .. import engine
from .._native import SOLVER_PUMPED_LANGEVIN, ALG_ORIGINAL, ALG_ADAM
from .ccvm_solver import CCVMSolver
from .langevin_solver import unit_box_map

LANGEVIN_SCALING_MULTIPLIER = 0.05
"""Multiplier used by PumpedLangevinSolver in get_scaling_factor()."""


class PumpedLangevinSolver(CCVMSolver):
    """Pumped Langevin dynamics as a system of SDEs.

    Args:
        device (str): "cuda" to solve ("cpu" is accepted for construction/validation only).
        problem_category (str): "boxqp".
        batch_size (int): trajectories per solve.  Default 1000.
    """

    _PARAMETER_KEYS = frozenset(["pump", "dt", "S", "iterations", "sigma", "feedback_scale"])

    def __init__(self, device, problem_category="boxqp", batch_size=1000):
        super().__init__(device)
        self.batch_size = batch_size
        self._scaling_multiplier = LANGEVIN_SCALING_MULTIPLIER
        self._method_selector(problem_category)

    # --------------------------------------------------------------------- hooks
    def _calculate_drift_boxqp(self, c, p, S, feedback_scale):
        """(-1 + p - c^2) c + feedback_scale * grads(c), using the bound instance's solution
        bounds (reference pumped_langevin_solver.py:95-114)."""
        lower, upper = self.solution_bounds
        return engine.eval_hook(SOLVER_PUMPED_LANGEVIN, "drift", self.q_matrix, self.v_vector, (c,), lower,
                                upper, S, pump=p, feedback_scale=feedback_scale)[0]

    def _calculate_grads_boxqp(self, c, lower_limit=0, upper_limit=1, S=1):
        """-((c a/(2S) + b/2) Q) a/(2S) - V a/(2S) (reference 116-147)."""
        return engine.eval_hook(SOLVER_PUMPED_LANGEVIN, "grads", self.q_matrix, self.v_vector, (c,),
                                lower_limit, upper_limit, S)[0]

    def _append_samples_to_file(self, c_sample, evolution_file_object, s_sample=None):
        """c rows, tab after every value (s_sample optional, see LangevinSolver)."""
        self._append_rows(c_sample, evolution_file_object)
        if s_sample is not None:
            self._append_rows(s_sample, evolution_file_object)

    def tune(self, instances, post_processor=None, pump_rate_flag=True, g=0.05):
        """Placeholder, as in the reference."""
        self._is_tuned = True

    # --------------------------------------------------------------------- loops
    def _solve(self, problem_size, batch_size, device, S, pump, dt, iterations, sigma, pump_rate_flag,
               feedback_scale, evolution_step_size, samples_taken):
        """Original pumped-Langevin loop -> c."""
        (c,) = self._engine_solve(SOLVER_PUMPED_LANGEVIN, ALG_ORIGINAL, batch_size, iterations, S,
                                  evolution_step_size, pump=pump, dt=dt, sigma=sigma,
                                  pump_rate_flag=pump_rate_flag, feedback_scale=feedback_scale)
        self._publish_samples(("c_sample",))
        return c

    def _solve_adam(self, problem_size, batch_size, device, S, pump, dt, iterations, sigma, pump_rate_flag,
                    feedback_scale, evolution_step_size, samples_taken, hyperparameters):
        """Pumped-Langevin loop with Adam on the gradient -> c."""
        (c,) = self._engine_solve(SOLVER_PUMPED_LANGEVIN, ALG_ADAM, batch_size, iterations, S,
                                  evolution_step_size, hyperparameters, pump=pump, dt=dt, sigma=sigma,
                                  pump_rate_flag=pump_rate_flag, feedback_scale=feedback_scale)
        self._publish_samples(("c_sample",))
        return c

    def __call__(self, instance, post_processor=None, pump_rate_flag=True, evolution_step_size=None,
                 evolution_file=None, algorithm_parameters=None):
        """Solve ``instance``; returns a Solution whose problem_variables are (c + S) / (2S),
        post-processed if asked (reference pumped_langevin_solver.py:451-658)."""
        self._check_device(instance)
        problem_size = instance.problem_size
        self._bind_instance(instance)
        pump, dt, S, iterations, sigma, feedback_scale = self._read_parameters(
            problem_size, ("pump", "dt", "S", "iterations", "sigma", "feedback_scale"))
        S = self._normalise_s(S, problem_size)

        def solve_args(adam):
            return (problem_size, self.batch_size, self.device, S, pump, dt, iterations, sigma, pump_rate_flag,
                    feedback_scale, evolution_step_size, 0 if evolution_step_size else None)

        def finish(c):
            return c, unit_box_map(S), None, lambda pv: {"problem_variables": pv}

        return self._run(SOLVER_PUMPED_LANGEVIN, instance, post_processor, evolution_step_size, evolution_file,
                         algorithm_parameters, iterations, S, solve_args, finish)
