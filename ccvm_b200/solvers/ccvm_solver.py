"""Base class of the four solvers: the reference's public surface (``parameter_key``,
``get_scaling_factor``, hook slots, machine time / energy bookkeeping) plus the shared
``__call__`` pipeline that drives the CUDA engine.

Mirrors ``ccvm_simulators/solvers/ccvm_solver.py`` (constructor 33-55, get_scaling_factor
134-150, _method_selector 152-170, machine energy/time 176-444).  The iteration loops are not
here and not in Python at all: ``_solve`` / ``_solve_adam`` of each subclass hand the whole loop
to one persistent sm_100a kernel through ``ccvm_b200.engine``.
"""
import enum
import time
from abc import ABC, abstractmethod

import numpy as np
import torch

from .. import engine
from ..solution import Solution
from .algorithms import AdamParameters


class DeviceType(enum.Enum):
    """Device strings the solvers accept."""

    CPU_DEVICE = "cpu"
    CUDA_DEVICE = "cuda"


class MachineType(enum.Enum):
    """Machines whose time / energy a solver can be asked to model."""

    CPU = "cpu"
    GPU = "gpu"
    FPGA = "fpga"
    DL_CCVM = "dl-ccvm"
    MF_CCVM = "mf-ccvm"


_CPU_POWER = {20: 4.93, 30: 5.19, 40: 5.0, 50: 5.01, 60: 5.0, 70: 5.22}
_GPU_POWER = {20: 28.93, 30: 29.8, 40: 31.09, 50: 31.29, 60: 31.49, 70: 32.28}
_HOOKS = ("calculate_drift", "calculate_grads", "change_variables", "fit_to_constraints")


class _PendingSolution:
    """An instance whose solve and tail were planned but not yet launched (``CCVMSolver.solve_many``)."""

    __slots__ = ("instance", "batch_size", "iterations", "plan", "epilogue", "make_solution")

    def __init__(self, instance, batch_size, iterations, plan, epilogue, make_solution):
        self.instance, self.batch_size, self.iterations = instance, batch_size, iterations
        self.plan, self.epilogue, self.make_solution = plan, epilogue, make_solution


class CCVMSolver(ABC):
    """Common behaviour of DLSolver, MFSolver, LangevinSolver and PumpedLangevinSolver.

    Args:
        device (str): "cpu" or "cuda".  Construction and validation work for both (as in the
            reference); solving requires "cuda" -- this package has no CPU path.
    """

    #: exact key set a ``parameter_key`` entry must have; set by subclasses
    _PARAMETER_KEYS = frozenset()

    def __init__(self, device):
        if device not in DeviceType._value2member_map_:
            raise ValueError("Given device is not available")
        self.device = device
        self._is_tuned = False
        self._scaling_multiplier = None
        self._parameter_key = None
        self._default_cpu_machine_parameters = {"cpu_power": dict(_CPU_POWER)}
        self._default_cuda_machine_parameters = {"gpu_power": dict(_GPU_POWER)}
        self.calculate_grads = None
        self.change_variables = None
        self.fit_to_constraints = None
        #: validation hook: a ``[iterations][K][N][batch]`` tensor of standard normals that the
        #: next solve consumes instead of Philox noise (K = 2 for DL, else 1).
        self.noise_source = None
        #: list collecting planned (not yet launched) solves while ``solve_many`` is gathering a batch
        self._deferred = None
        #: optional queue of (seed, offset) noise streams for the next solves instead of torch's CUDA
        #: generator state: sweeps key every instance by its GLOBAL index, so results do not depend on
        #: how instances are dealt to ranks or chunks (and no two ranks ever draw the same stream)
        self.noise_streams = None
        #: cumulative host seconds spent planning + enqueueing (launch_many) and building Solutions
        #: (collect_many, after its event wait): what a sweep has to hide behind the kernels
        self.host_seconds = {"launch": 0.0, "collect": 0.0}

    # ------------------------------------------------------------------ properties
    @property
    def is_tuned(self):
        """bool: True if the current parameters were set by ``tune()``."""
        return self._is_tuned

    @property
    def parameter_key(self):
        """dict: solver parameters keyed by problem size."""
        return self._parameter_key

    @parameter_key.setter
    def parameter_key(self, parameters):
        expected = set(self._PARAMETER_KEYS)
        for entry in parameters.values():
            if entry.keys() != expected:
                raise ValueError(
                    "The parameter key is not valid for this solver. Expected keys: "
                    + str(expected)
                    + " Given keys: "
                    + str(entry.keys())
                )
        self._parameter_key = parameters
        self._is_tuned = False

    # ------------------------------------------------------------ abstract interface
    @abstractmethod
    def tune(self):
        """Placeholder in the reference; kept for interface parity."""

    @abstractmethod
    def _solve(self):
        """Original algorithm: one persistent-kernel launch."""

    @abstractmethod
    def _solve_adam(self):
        """Adam-enhanced algorithm: one persistent-kernel launch."""

    @abstractmethod
    def _calculate_drift_boxqp(self, **kwargs):
        pass

    @abstractmethod
    def _calculate_grads_boxqp(self, **kwargs):
        pass

    # -------------------------------------------------------------- shared operators
    def _change_variables_boxqp(self, problem_variables, lower_limit=0, upper_limit=1, S=1):
        """x = 0.5 * y / S * (u - l) + 0.5 * (u + l)  (reference dl_solver.py:219-235), on the device."""
        return engine.change_variables(problem_variables, lower_limit, upper_limit, S)

    def _fit_to_constraints_boxqp(self, c, lower_clamp, upper_clamp):
        """Box clamp (reference dl_solver.py:237-250).  Bounds may be scalars or tensors."""
        return engine.clamp(c, lower_clamp, upper_clamp)

    def get_scaling_factor(self, q_matrix):
        """sqrt(sum |Q|) * solver multiplier (reference ccvm_solver.py:134-150), reduced on device."""
        return engine.scaling_factor(q_matrix, self._scaling_multiplier)

    def _method_selector(self, problem_category):
        """Bind the problem-specific hooks (reference ccvm_solver.py:152-170)."""
        if problem_category.lower() == "boxqp":
            self.calculate_drift = self._calculate_drift_boxqp
            self.calculate_grads = self._calculate_grads_boxqp
            self.change_variables = self._change_variables_boxqp
            self.fit_to_constraints = self._fit_to_constraints_boxqp
        else:
            raise ValueError(
                "The given instance is not a valid problem category."
                f" Given category: {problem_category}"
            )

    def _require_stock_hooks(self):
        """The fused kernel implements the stock BoxQP hooks; it cannot honour replacements
        (the reference's tests swap them for mocks, test_mf_solver.py:262-266).  Fail loudly
        instead of silently ignoring an override."""
        stock = {
            "calculate_drift": "_calculate_drift_boxqp",
            "calculate_grads": "_calculate_grads_boxqp",
            "change_variables": "_change_variables_boxqp",
            "fit_to_constraints": "_fit_to_constraints_boxqp",
        }
        for slot, name in stock.items():
            bound = getattr(self, slot, None)
            if getattr(bound, "__func__", None) is not getattr(type(self), name):
                raise RuntimeError(
                    f"{type(self).__name__}.{slot} was replaced; the fused CUDA loop only implements the "
                    "stock boxqp hooks and cannot call a Python override."
                )

    # --------------------------------------------------------------- call pipeline
    def _check_device(self, instance):
        if instance.device != self.device:
            raise ValueError(
                f"The device type of the instance ({instance.device}) and the solver"
                f" ({self.device}) must match."
            )

    def _bind_instance(self, instance):
        self.q_matrix = instance.q_matrix
        self.v_vector = instance.v_vector
        self.solution_bounds = instance.solution_bounds

    def _read_parameters(self, problem_size, names):
        try:
            entry = self.parameter_key[problem_size]
            return [entry[name] for name in names]
        except KeyError as e:
            raise KeyError(
                f"The parameter '{e.args[0]}' for the given instance size is not defined."
            ) from e

    @staticmethod
    def _normalise_s(S, problem_size):
        """A 1-D tensor S must have one entry per variable (reference dl_solver.py:843-848).  The
        reference broadcasts it to (B, N); the engine keeps the vector."""
        if torch.is_tensor(S) and S.ndim == 1:
            if S.size(dim=0) != problem_size:
                raise ValueError("Tensor S size should be equal to problem size.")
        return S

    @staticmethod
    def _evolution_plan(instance, iterations, evolution_step_size, evolution_file):
        """Number of snapshots and file name (reference dl_solver.py:856-887)."""
        if not evolution_step_size:
            return None, evolution_file
        if evolution_step_size < 1:
            raise ValueError("The evolution step size must be greater than or equal to 1.")
        if evolution_file is None:
            evolution_file = f"./{instance.name}_evolution.txt"
        num_samples = int(iterations / evolution_step_size) + 1
        if iterations % evolution_step_size != 0:
            num_samples += 1
        return num_samples, evolution_file

    def _engine_solve(self, solver_id, algorithm, batch_size, iterations, S, evolution_step_size,
                      hyperparameters=None, **scalars):
        """Common body of every ``_solve`` / ``_solve_adam``.  Called directly (the reference's
        function-level API) it launches the loop kernel and returns the state tensors.  Inside
        ``__call__`` / ``solve_many`` it only PLANS the launch (descriptor + output tensors) so that
        the loop and the tail of the call can go to the GPU as one fused launch."""
        if self.device != "cuda":
            raise engine.nat.NativeError(
                "ccvm_b200 solves on CUDA only (device='cuda'); there is no CPU implementation.")
        self._require_stock_hooks()
        s_vec, s_val = (S if S.ndim == 1 else S[0], 0.0) if torch.is_tensor(S) and S.numel() > 1 else (None, float(S))
        lower, upper = self.solution_bounds
        num_samples, _ = self._evolution_plan(None, iterations, evolution_step_size, "unused")
        kwargs = dict(lower=lower, upper=upper, s=s_val, s_vec=s_vec, hyperparameters=hyperparameters,
                      noise=self.noise_source, evolution_step=evolution_step_size or None,
                      num_samples=num_samples or 0, **scalars)
        if self.noise_streams and self.noise_source is None:
            kwargs["seed"], kwargs["offset"] = self.noise_streams.pop(0)
        if self._deferred is not None:
            plan = engine.plan_solve(solver_id, algorithm, self.q_matrix, self.v_vector, batch_size, iterations,
                                     **kwargs)
            self._deferred.append(plan)
            self._samples = plan.samples
            return plan.outputs
        outs, samples = engine.solve(solver_id, algorithm, self.q_matrix, self.v_vector, batch_size, iterations,
                                     **kwargs)
        self._samples = samples
        return outs

    def _run(self, solver_id, instance, post_processor, evolution_step_size, evolution_file,
             algorithm_parameters, iterations, S, solve_args, finish):
        """Shared ``__call__`` body.  ``solve_args(adam)`` yields the positional arguments of
        ``_solve`` / ``_solve_adam``; ``finish(outs)`` maps raw loop outputs to
        (state, map1, map2, variables-dict-builder).

        The whole call is ONE kernel launch (``ccvm_solve_fused``): schedules, iteration loop, change of
        variables, post-processor, energy and solution statistics, followed by one 56-byte read-back.
        ``solve_time`` / ``pp_time`` split the measured wall time of that launch (with a stream
        synchronise on both sides -- the reference stops its clock without one) in the
        ratio of the device-measured durations of the loop and of the tail; ``pp_time`` therefore
        covers change of variables + post-processor + energy + statistics, and is 0 without a
        post-processor, as in the reference."""
        batch_size = self.batch_size
        num_samples, evolution_file = self._evolution_plan(instance, iterations, evolution_step_size,
                                                           evolution_file)
        self._samples = None
        in_batch = self._deferred is not None          # solve_many is collecting plans
        if in_batch and (self.noise_source is not None or evolution_step_size):
            raise ValueError("solve_many supports neither noise replay nor evolution sampling.")
        if not in_batch:
            self._deferred = []
        try:
            if algorithm_parameters is None:
                outs = self._solve(*solve_args(False))
            elif isinstance(algorithm_parameters, AdamParameters):
                outs = self._solve_adam(*solve_args(True), algorithm_parameters.to_dict())
            else:
                raise ValueError(f"Solver option type {type(algorithm_parameters)} is not supported.")
            plan = self._deferred[-1]
        finally:
            if not in_batch:
                self._deferred = None

        state, map1, map2, make_variables = finish(outs)
        epi = engine.plan_epilogue(batch_size, instance.problem_size, plan.device, map1=map1,
                                   post_processor=post_processor, pp_iterations=10, map2=map2,
                                   scaled_by=instance.scaled_by)   # stays on the device if it is there

        def make_solution(pv, objval, solve_time, pp_time, stats=None):
            return Solution(
                problem_size=instance.problem_size,
                batch_size=batch_size,
                instance_name=instance.name,
                iterations=iterations,
                objective_values=objval,
                solve_time=solve_time,
                pp_time=pp_time,
                optimal_value=instance.optimal_sol,
                best_value=instance.best_sol,
                num_frac_values=instance.num_frac_values,
                solution_vector=instance.solution_vector,
                variables=make_variables(pv),
                device=self.device,
                precomputed_stats=stats,
            )

        if in_batch:
            return _PendingSolution(instance, batch_size, iterations, plan, epi, make_solution)

        stream = torch.cuda.current_stream(plan.device)
        stream.synchronize()  # stream-level: other streams may be solving too
        start = time.time()
        with engine.nvtx_range(f"ccvm_b200.{type(self).__name__}.__call__ n={instance.problem_size} B={batch_size}"):
            raw = engine.solve_fused(plan, epi, _as_float(instance.optimal_sol))
            res = engine.decode_fused_results(raw.cpu())[0]   # the copy synchronises the stream
        wall = time.time() - start
        device_ns = res["loop_ns"] + res["tail_ns"]
        loop_share = res["loop_ns"] / device_ns if device_ns else 1.0
        solve_time = wall * loop_share / batch_size
        pp_time = wall * (1.0 - loop_share) / batch_size if post_processor else 0.0

        if evolution_step_size:
            self._publish_samples(self._sample_names)   # the snapshots exist now that the launch has run
            self._write_evolution(evolution_file, epi.energy)

        solution = make_solution(epi.pv, epi.energy, solve_time, pp_time,
                                 (res["best"], res["arg_best"], res["counts"]))
        if evolution_step_size:
            solution.evolution_file = evolution_file
        return solution

    # ------------------------------------------------------------- many instances
    def solve_many(self, instances, post_processor=None, algorithm_parameters=None, **call_kwargs):
        """Solve a sequence of instances with ONE launch per kernel bucket (grid over instances x
        trajectory blocks, ``ccvm_solve_batch_fused``): every CTA also finishes its own trajectories
        (change of variables, post-processor, energy) and merges the statistics of its instance, so a
        chunk ends with a single device->host copy of 56 bytes per instance.

        Returns the list of ``Solution`` objects ``[self(instance=i, ...) for i in instances]``
        would return under the same generator state (the reference's user loop over instance
        files, examples/ccvm_boxqp_dl.py:27-52), except that ``solve_time`` / ``pp_time`` are the
        measured device times of the shared launches apportioned by each instance's share of the
        drift work (batch x iterations x N^2) -- a batch-1000 solve fills a fraction of the GPU, so
        concurrent instances are how the machine is kept busy.

        ``launch_many`` / ``collect_many`` are the two halves: a sweep launches chunk k+1 before it
        collects chunk k, so that planning and result handling on the host overlap the kernels."""
        return self.collect_many(self.launch_many(instances, post_processor, algorithm_parameters, **call_kwargs))

    def launch_many(self, instances, post_processor=None, algorithm_parameters=None, **call_kwargs):
        """Plan and enqueue one chunk (kernels + the asynchronous read-back of its result blocks into
        pinned memory); returns a handle for ``collect_many``.  Does not synchronise."""
        if self.device != "cuda":
            raise engine.nat.NativeError(
                "ccvm_b200 solves on CUDA only (device='cuda'); there is no CPU implementation.")
        t_host = time.perf_counter()
        instances = list(instances)
        if not instances:
            return None
        if call_kwargs.get("evolution_step_size"):
            raise ValueError("solve_many does not support evolution sampling.")
        self._deferred = []
        try:
            pending = [self(instance=inst, post_processor=post_processor,
                            algorithm_parameters=algorithm_parameters, **call_kwargs) for inst in instances]
            plans = self._deferred
        finally:
            self._deferred = None
        dev = plans[0].device
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        offsets = [0]
        for p in pending:
            offsets.append(offsets[-1] + p.batch_size)
        # one energy vector for the chunk: the per-instance epilogues write their slices of it
        energy = torch.empty(offsets[-1], dtype=torch.float32, device=dev)
        for i, p in enumerate(pending):
            sl = energy[offsets[i]:offsets[i + 1]]
            p.epilogue.energy, p.epilogue.desc.energy = sl, sl.data_ptr()
        stream = torch.cuda.current_stream(dev)
        ev[0].record(stream)
        with engine.nvtx_range(f"ccvm_b200.{type(self).__name__}.solve_many x{len(pending)}"):
            raw = engine.solve_batch_fused(plans, [p.epilogue for p in pending],
                                           [_as_float(p.instance.optimal_sol) for p in pending])
        ev[1].record(stream)
        host = torch.empty(raw.shape, dtype=torch.uint8).pin_memory()
        host.copy_(raw, non_blocking=True)           # ONE device->host copy for the whole chunk
        done = torch.cuda.Event()
        done.record(stream)
        self.host_seconds["launch"] += time.perf_counter() - t_host
        return (pending, plans, ev, host, done, raw, bool(post_processor))

    def collect_many(self, handle):
        """Wait for a chunk launched by ``launch_many`` and build its ``Solution`` objects."""
        if handle is None:
            return []
        pending, plans, ev, host, done, _raw, has_pp = handle
        done.synchronize()
        t_host = time.perf_counter()
        stats = engine.decode_fused_results(host)
        t_total = ev[0].elapsed_time(ev[1]) * 1e-3
        work = [p.batch_size * p.iterations * p.instance.problem_size ** 2 for p in pending]
        total = float(sum(work)) or 1.0
        out = []
        for i, p in enumerate(pending):
            r = stats[i]
            device_ns = r["loop_ns"] + r["tail_ns"]
            loop_share = r["loop_ns"] / device_ns if device_ns else 1.0
            t_inst = t_total * work[i] / total
            out.append(p.make_solution(p.epilogue.pv, p.epilogue.energy, t_inst * loop_share / p.batch_size,
                                       (t_inst * (1.0 - loop_share) / p.batch_size) if has_pp else 0.0,
                                       (r["best"], r["arg_best"], r["counts"])))
        self.host_seconds["collect"] += time.perf_counter() - t_host
        return out

    # ----------------------------------------------------------- evolution sampling
    _EVOLUTION_TRAILING_TAB = True

    def _publish_samples(self, names):
        """Expose the engine's [K][samples][B][N] snapshot buffer as the reference's per-array
        CPU tensors of shape (B, N, samples) (dl_solver.py:877-886)."""
        self._sample_names = names
        for k, name in enumerate(names):
            val = None
            if self._samples is not None:
                val = self._samples[k].permute(1, 2, 0).contiguous().cpu()
            setattr(self, name, val)

    def _write_evolution(self, evolution_file, objval):
        batch_index = int(torch.argmax(-objval).item())
        with open(evolution_file, "w") as fh:
            for k in range(self._samples.shape[0]):
                rows = self._samples[k, :, batch_index, :].transpose(0, 1).cpu()  # (N, samples)
                self._append_rows(rows, fh)

    def _append_rows(self, rows, fh):
        """One text line per variable, samples tab-separated and rounded to 4 decimals
        (reference dl_solver.py:252-281; MF omits the trailing tab, mf_solver.py:267-300)."""
        for nn in range(rows.shape[0]):
            cells = [str(round(rows[nn, ii].item(), 4)) for ii in range(rows.shape[1])]
            if self._EVOLUTION_TRAILING_TAB:
                fh.write("".join(cell + "\t" for cell in cells))
            else:
                fh.write("\t".join(cells))
            fh.write("\n")

    # ------------------------------------------------- machine energy / time models
    # Closed-form bookkeeping (no device work), reference ccvm_solver.py:176-444.
    def _validate_machine_energy_dataframe_columns(self, dataframe):
        missing_columns = [c for c in ("pp_time", "iterations") if c not in dataframe.columns]
        if missing_columns:
            raise ValueError(f"The given dataframe is missing the following columns: {missing_columns}")

    @staticmethod
    def _mean_solve_time(dataframe):
        if "solve_time" not in dataframe.columns:
            raise ValueError("The given dataframe does not contain the column 'solve_time'")
        return np.mean(dataframe["solve_time"].values)

    def _power_times_time(self, machine_parameters, defaults, key):
        if machine_parameters is None:
            machine_parameters = defaults
        elif key not in machine_parameters.keys():
            raise ValueError(
                "The given machine parameters are not valid. "
                f"The dictionary must contain the key '{key}'"
            )

        def energy_callable(dataframe, problem_size):
            return machine_parameters[key][problem_size] * self._mean_solve_time(dataframe)

        return energy_callable

    def _cpu_machine_energy(self, machine_parameters=None):
        return self._power_times_time(machine_parameters, self._default_cpu_machine_parameters, "cpu_power")

    def _cuda_machine_energy(self, machine_parameters=None):
        return self._power_times_time(machine_parameters, self._default_cuda_machine_parameters, "gpu_power")

    def _cpu_gpu_machine_time(self, **_):
        def time_callable(dataframe, **_):
            return self._mean_solve_time(dataframe)

        return time_callable

    def _machine_dispatch(self, machine, table, what):
        if machine not in table:
            raise ValueError(
                f"The given machine type is not valid. "
                f"The machine type must be one of {', '.join(table.keys())}"
            )
        method = table[machine]
        if not method:
            raise ValueError(
                f"Mismatch between the solver and the machine type. "
                f"Provided machine type: {machine}, solver type: {self.__class__.__name__}"
            )
        return method

    def _optional(self, cls_name, attr):
        return getattr(self, attr, None) if self.__class__.__name__ == cls_name else None

    def machine_energy(self, machine, machine_parameters=None):
        """Callable(dataframe, problem_size) -> mean energy on ``machine``."""
        table = {
            "cpu": self._cpu_machine_energy,
            "gpu": self._cuda_machine_energy,
            "dl-ccvm": self._optional("DLSolver", "_optics_machine_energy"),
            "mf-ccvm": self._optional("MFSolver", "_optics_machine_energy"),
            "fpga": self._optional("LangevinSolver", "_fpga_machine_energy"),
        }
        return self._machine_dispatch(machine, table, "energy")(machine_parameters)

    def machine_time(self, machine, machine_parameters=None):
        """Callable(dataframe, problem_size) -> mean time per instance on ``machine``."""
        table = {
            "cpu": self._cpu_gpu_machine_time,
            "gpu": self._cpu_gpu_machine_time,
            "dl-ccvm": self._optional("DLSolver", "_optics_machine_time"),
            "mf-ccvm": self._optional("MFSolver", "_optics_machine_time"),
            "fpga": self._optional("LangevinSolver", "_fpga_machine_time"),
        }
        return self._machine_dispatch(machine, table, "time")(machine_parameters=machine_parameters)


# ------------------------------------------------------------------------ helpers
def _as_float(x):
    return float(x.item()) if torch.is_tensor(x) else float(x)
