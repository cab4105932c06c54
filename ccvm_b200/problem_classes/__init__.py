from .boxqp import ProblemInstance, InstanceType, DeviceType
