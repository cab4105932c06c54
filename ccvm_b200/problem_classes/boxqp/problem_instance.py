"""BoxQP problem instance: same attributes, file format and methods as the reference's
``problem_classes/boxqp/problem_instance.py``.

The loader parses the whole file on the host in bulk (the reference writes the tensor one
element at a time, 185-188) and moves Q, V to the device with one copy each;
``compute_energy`` and ``scale_coefs`` run on the device through the C ABI."""
import enum

import numpy as np
import torch

from ... import engine


class DeviceType(enum.Enum):
    CPU_DEVICE = "cpu"
    CUDA_DEVICE = "cuda"


class InstanceType(enum.Enum):
    TUNING = "tuning"
    TEST = "test"


class ProblemInstance:
    """A BoxQP instance: maximise the file's objective <=> minimise 1/2 xQx + Vx with the NEGATED
    coefficients stored here, subject to lower <= x <= upper.

    Args:
        device (str): where q_matrix / v_vector live ("cpu" or "cuda").
        instance_type (str): "tuning" or "test".
        file_path (str): instance file to load (optional).
        file_delimiter (str): field delimiter, default tab.
        name (str): instance name; defaults to the file's base name.
        solution_bounds (tuple): (lower, upper), default (0.0, 1.0).
    """

    def __init__(self, device="cpu", instance_type="tuning", file_path=None, file_delimiter="\t", name=None,
                 solution_bounds=(0.0, 1.0)):
        self.problem_size = None
        self.optimal_sol = None
        self.best_sol = None
        self.optimality = None
        self.sol_time_gb = None
        self.sol_time_bfgs = None
        self.num_frac_values = None
        self.q_matrix = None
        self.v_vector = None
        self.solution_vector = None
        self.scaled_by = 1
        self.device = device
        self._custom_name = False
        self.file_delimiter = file_delimiter
        if instance_type not in {item.value for item in InstanceType}:
            raise ValueError("instance_type must be tuning or test")
        self.instance_type = instance_type
        if name:
            self.name = name
            self._custom_name = True
        if file_path:
            self.file_path = file_path
            self.load_instance(device=device, instance_type=instance_type, file_path=file_path,
                               file_delimiter=file_delimiter)
        self.problem_category = "boxqp"
        self.solution_bounds = solution_bounds

    @property
    def solution_bounds(self):
        """tuple(float): inclusive (minimum, maximum) of every solution entry."""
        return self._solution_bounds

    @solution_bounds.setter
    def solution_bounds(self, bounds):
        if len(bounds) != 2:
            raise ValueError("solution_bounds must be a tuple of size 2, containing the minimum and maximum bounds (inclusive)")
        elif bounds[0] >= bounds[1]:
            raise ValueError("Minimum solution bound must be less than maximum solution bound")
        self._solution_bounds = bounds

    def load_instance(self, device="cpu", instance_type="tuning", file_path=None, file_delimiter=None):
        """Read an instance file: line 1 = size, optimum, best, optimality, Gurobi time, BFGS time,
        seed (ignored), number of fractional values; line 2 = V; next N lines = Q; optional last
        line = the solver's solution vector.  Q and V are negated on load."""
        if not file_path and not getattr(self, "file_path", None):
            raise Exception("No file path specified, cannot load instance.")
        if file_path:
            self.file_path = file_path
        file_path = self.file_path
        if file_delimiter:
            self.file_delimiter = file_delimiter
        delim = self.file_delimiter
        try:
            with open(file_path, "r") as stream:
                lines = stream.readlines()
            head = lines[0].split("\n")[0].split(delim)
            n = int(head[0])
            optimal_sol, best_sol = float(head[1]), float(head[2])
            optimality = head[3].lower() == "true"
            sol_time_gb, sol_time_bfgs = float(head[4]), float(head[5])
            num_frac_values = int(head[7])
            v_host = -np.array([float(t) for t in lines[1].split("\n")[0].split(delim)[:n]], dtype=np.float32)
            q_host = -np.array([[float(t) for t in ln.split("\n")[0].split(delim)[:n]] for ln in lines[2:n + 2]],
                               dtype=np.float32)
            if q_host.shape != (n, n) or v_host.shape != (n,):
                raise ValueError(f"expected a {n}x{n} matrix and a {n}-vector")
            solution_vector = []
            if len(lines) > n + 2:
                solution_vector = [float(t) for t in lines[n + 2].split("\n")[0].split(delim) if not t == ""]
        except Exception as e:
            raise Exception("Error reading instance file: " + str(e))

        self.device = device
        self.instance_type = instance_type
        self.problem_size = n
        self.optimal_sol = optimal_sol
        self.best_sol = best_sol
        self.optimality = optimality
        self.sol_time_gb = sol_time_gb
        self.sol_time_bfgs = sol_time_bfgs
        self.num_frac_values = num_frac_values
        self.q_matrix = torch.from_numpy(q_host).to(device)
        self.v_vector = torch.from_numpy(v_host).to(device)
        self.solution_vector = solution_vector
        self.scaled_by = 1
        if not self._custom_name:
            self.name = file_path.split("/")[-1].split(".")[0]

    def compute_energy(self, confs):
        """E_b = (1/2 x_b Q x_b + V . x_b) * scaled_by for every row of ``confs`` (device kernel)."""
        sb = self.scaled_by
        sb = float(sb.item()) if torch.is_tensor(sb) and sb.numel() == 1 else sb
        return engine.compute_energy(confs, self.q_matrix, self.v_vector, sb)

    def scale_coefs(self, scaling_factor):
        """Divide Q and V by ``scaling_factor`` (scalar or tensor) and accumulate it in
        ``scaled_by``; calls stack multiplicatively."""
        self.q_matrix, self.v_vector = engine.scale_coefs(self.q_matrix, self.v_vector, scaling_factor)
        self.scaled_by = self.scaled_by * scaling_factor
