from .problem_instance import ProblemInstance, InstanceType, DeviceType
