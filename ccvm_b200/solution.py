"""Solution of one solve: same fields, metadata and statistics as the reference's
``ccvm_simulators/solution.py``.  The statistics (best value, seven gap-threshold success
fractions, 65-146 there) are reduced on the device by one kernel and read back with a single
36-byte copy instead of eight ``.item()`` synchronisations."""
import os
import copy
from dataclasses import dataclass, field, fields

import torch

from . import engine

_PERF_KEYS = ("optimal", "one_percent", "two_percent", "three_percent", "four_percent", "five_percent",
              "ten_percent")


@dataclass
class Solution:
    """Result record of ``Solver.__call__``.

    Fields mirror the reference one for one: problem_size, batch_size, instance_name, iterations,
    objective_values (tensor, excluded from repr/metadata), solve_time, pp_time, optimal_value,
    best_value, num_frac_values, solution_vector, variables (dict of tensors, excluded),
    evolution_file, device, and the derived solution_performance / best_objective_value.
    ``best_index`` (extra) is the batch index of the best trajectory.
    """

    problem_size: int
    batch_size: int
    instance_name: str
    iterations: int
    objective_values: torch.Tensor = field(repr=False)
    solve_time: float
    pp_time: float
    optimal_value: float
    best_value: float
    num_frac_values: int
    solution_vector: list
    variables: dict = field(repr=False)
    evolution_file: str = None
    device: str = field(default="cpu", repr=False)
    solution_performance: dict = None
    best_objective_value: float = None
    best_index: int = field(default=None, repr=False)
    # (best, arg_best, counts[7]) already reduced on the device by a batched launch (solve_many)
    precomputed_stats: tuple = field(default=None, repr=False)

    def __post_init__(self):
        target = torch.device(self.device)
        for key, value in self.variables.items():
            if torch.is_tensor(value) and value.device.type != target.type:
                self.variables[key] = value.to(self.device)
        self.get_solution_stats()
        if torch.is_tensor(self.objective_values) and self.objective_values.device.type != target.type:
            self.objective_values = self.objective_values.to(self.device)

    def get_solution_stats(self):
        """best_objective_value = max(-E); fraction of trajectories whose gap
        (optimal - (-E)) * 100 / |E| is within 0.1, 1, 2, 3, 4, 5, 10 %, rounded to 4 digits."""
        obj = self.objective_values
        if not obj.is_cuda:
            obj = engine.to_engine_device(obj)
        if self.precomputed_stats is not None:
            best, arg, counts = self.precomputed_stats
        else:
            best, arg, counts = engine.solution_stats(obj, self.optimal_value)
        self.best_objective_value = best
        self.best_index = arg
        n = obj.numel()
        self.solution_performance = {k: round(c / n, 4) for k, c in zip(_PERF_KEYS, counts)}

    def get_metadata_dict(self) -> dict:
        """All fields that take part in repr (i.e. everything but the tensors)."""
        return {f.name: copy.deepcopy(getattr(self, f.name)) for f in fields(self) if f.repr}

    def save_tensor_to_file(self, tensor_name, file_dir=".", file_name=None):
        """``torch.save`` one entry of ``variables`` to ``<file_dir>/<file_name>.pt``."""
        try:
            if file_dir != "." and not os.path.isdir(file_dir):
                os.makedirs(file_dir)
                print("The folder to store doesn't exist yet. Creating: ", file_dir)
        except Exception as e:
            raise Exception(f"Failed to create the folder path: {e}")
        if tensor_name not in self.variables.keys():
            raise Exception(f"Cannot find the {tensor_name} in the variables dictionary.")
        if not file_name:
            file_name = tensor_name
        tensor_value = self.variables[tensor_name]
        if not torch.is_tensor(tensor_value):
            raise Exception(f"A tensor object cannot be obtained by the given tensor_name: {tensor_name}")
        torch.save(tensor_value, f"{file_dir}/{file_name}.pt")
        print("Successfully saved the tensor!")
