from .post_processor import PostProcessor, MethodType
from .adam import PostProcessorAdam
from .grad_descent import PostProcessorGradDescent
from .factory import PostProcessorFactory
