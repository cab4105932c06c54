"""Post-processor interface (reference post_processor/post_processor.py)."""
from abc import ABC, abstractmethod
from enum import Enum


class MethodType(str, Enum):
    BFGS = "bfgs"
    LBFGS = "lbfgs"
    Adam = "adam"
    ASGD = "asgd"
    GradDescent = "grad-descent"


class PostProcessor(ABC):
    """A post-processor refines a batch of solutions in place of the solver's raw output."""

    @abstractmethod
    def postprocess(self):
        pass


def require_tensors(c, q_matrix, v_vector):
    """Argument checks with the reference's messages (adam.py:47-53)."""
    import torch
    if not torch.is_tensor(c):
        raise TypeError("parameter c must be a tensor")
    if not torch.is_tensor(q_matrix):
        raise TypeError("parameter q_matrix must be a tensor")
    if not torch.is_tensor(v_vector):
        raise TypeError("parameter v_vector must be a tensor")
    if c.dim() != 2 or q_matrix.dim() != 2 or q_matrix.shape[0] != q_matrix.shape[1] \
            or c.shape[1] != q_matrix.shape[0] or v_vector.numel() != q_matrix.shape[0]:
        raise RuntimeError(
            f"shape mismatch: c {tuple(c.shape)}, q_matrix {tuple(q_matrix.shape)}, v_vector {tuple(v_vector.shape)}")
