"""Batched projected Adam step (reference post_processor/adam.py:15-69) as one CUDA kernel.

What the reference effectively does: ONE torch.optim.Adam(lr=0.01, betas=(0.9, 0.99)) step on
1/2 xQx + Vx, then the box clamp.  Its later iterations keep stepping the ORIGINAL Parameter
while returning a re-wrapped clamp of it, so ``num_iter > 1`` returns exactly the ``num_iter=1``
result (verified); ``num_iter`` is accepted and has the same (non-)effect here."""
import time

import torch

from .. import engine
from .post_processor import PostProcessor, MethodType, require_tensors


class PostProcessorAdam(PostProcessor):
    def __init__(self):
        self.pp_time = 0
        self.method_type = MethodType.Adam

    def postprocess(self, c, q_matrix, v_vector, lower_clamp=0.0, upper_clamp=1.0, num_iter=1, device="cpu"):
        start_time = time.time()
        require_tensors(c, q_matrix, v_vector)
        x = engine.to_engine_device(c).to(torch.float32).clone()
        if num_iter >= 1:
            engine.postprocess_adam(x, q_matrix.to(x.device), v_vector.to(x.device), 0.01, lower_clamp, upper_clamp)
        torch.cuda.synchronize(x.device)
        self.pp_time = time.time() - start_time
        return x.to(c.device)
