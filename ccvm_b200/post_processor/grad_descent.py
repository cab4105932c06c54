"""Batched projected gradient descent (reference post_processor/grad_descent.py:13-68) as one
CUDA kernel: num_iter_pp x { x <- clamp(x - step_size (xQ + V), lower, upper) }."""
import time

import torch

from .. import engine
from .post_processor import PostProcessor, require_tensors


class PostProcessorGradDescent(PostProcessor):
    def __init__(self):
        self.pp_time = 0

    def postprocess(self, c, q_matrix, v_vector, lower_clamp=0.0, upper_clamp=1.0, num_iter_main=1000,
                    num_iter_pp=None, step_size=0.1):
        """Returns the refined (B, N) tensor.  ``num_iter_pp`` defaults to 1 % of ``num_iter_main``.
        Unlike the reference (whose first step mutates the caller's tensor) the input is left
        untouched."""
        start_time = time.time()
        require_tensors(c, q_matrix, v_vector)
        if num_iter_pp is None:
            num_iter_pp = int(num_iter_main * 0.01)
        x = engine.to_engine_device(c).to(torch.float32).clone()
        engine.postprocess_grad_descent(x, q_matrix.to(x.device), v_vector.to(x.device), num_iter_pp, step_size,
                                        lower_clamp, upper_clamp)
        torch.cuda.synchronize(x.device)
        self.pp_time = time.time() - start_time
        return x.to(c.device)
