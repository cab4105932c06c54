from .ccvm_solver import CCVMSolver, MachineType, DeviceType
from .algorithms import AdamParameters
from .dl_solver import DLSolver
from .mf_solver import MFSolver
from .langevin_solver import LangevinSolver
from .pumped_langevin_solver import PumpedLangevinSolver
