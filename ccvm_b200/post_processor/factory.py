"""Post-processor factory (reference post_processor/factory.py:12-35).

Only the two batched, device-friendly methods are on this package's hot path.  The reference's
asgd / bfgs / lbfgs post-processors are serial per-trajectory host loops over third-party
optimisers (scipy / torch.optim) and are out of scope; asking for them raises
NotImplementedError rather than silently running on the CPU."""
from .post_processor import MethodType
from .adam import PostProcessorAdam
from .grad_descent import PostProcessorGradDescent


class PostProcessorFactory:
    @staticmethod
    def create_postprocessor(method):
        key = method.lower()
        if key == MethodType.Adam.value:
            return PostProcessorAdam()
        if key == MethodType.GradDescent.value:
            return PostProcessorGradDescent()
        if key in (MethodType.BFGS.value, MethodType.LBFGS.value, MethodType.ASGD.value):
            raise NotImplementedError(
                f"post-processor '{method}' is a per-trajectory host optimiser in the reference and has no "
                "B200 implementation; use 'adam' or 'grad-descent'.")
        raise AssertionError(f"Method type is not valid. Provided: {method}")
