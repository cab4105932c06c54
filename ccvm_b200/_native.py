"""ctypes binding of ``libccvm_b200.so`` (C ABI declared in ``include/ccvm_b200.h``).

This is the ONLY compute path of the package.  There is no CPU or eager-torch fallback: if the
shared library is missing, or no CUDA device is present, every compute call raises.
torch is used for device memory, streams and ``torch.distributed`` only.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CCVM_B200_LIB: alternative build of the same library (kernel experiments); default is the in-tree one
LIB_PATH = os.environ.get("CCVM_B200_LIB") or os.path.join(_HERE, "libccvm_b200.so")

SOLVER_DL, SOLVER_MF, SOLVER_LANGEVIN, SOLVER_PUMPED_LANGEVIN = 0, 1, 2, 3
ALG_ORIGINAL, ALG_ADAM = 0, 1
RNG_PHILOX, RNG_REPLAY = 0, 1
PP_NONE, PP_GRAD_DESCENT, PP_ADAM = 0, 1, 2
PP_IDS = {None: PP_NONE, "grad-descent": PP_GRAD_DESCENT, "adam": PP_ADAM}

EXPORTS = (
    "ccvm_solve", "ccvm_epilogue", "ccvm_compute_energy", "ccvm_postprocess_grad_descent",
    "ccvm_postprocess_adam", "ccvm_solution_stats", "ccvm_scaling_factor", "ccvm_solve_host",
    "ccvm_microbench_fp32", "ccvm_query_launch", "ccvm_abi_version", "ccvm_last_error",
    "ccvm_eval_hook", "ccvm_change_variables", "ccvm_fit_to_constraints", "ccvm_scale_coefs",
    "ccvm_solve_batch", "ccvm_solution_stats_batch", "ccvm_generate_boxqp", "ccvm_microbench_tf32",
    "ccvm_pack_record", "ccvm_merge_records", "ccvm_solve_fused", "ccvm_solve_batch_fused", "ccvm_dump_noise",
)
ABI_VERSION = 2          # CCVM_ABI_VERSION of include/ccvm_b200.h this binding was written against
RECORD_HEADER = 10       # CCVM_RECORD_HEADER
FUSED_RESULT_BYTES = 56  # { float best; int32 arg_best; int32 counts[7]; uint32 ctas; uint64 loop_ns, tail_ns }

_fp = C.c_void_p  # device / host pointers travel as plain addresses


class SolveDesc(C.Structure):
    _fields_ = [
        ("solver", C.c_int32), ("algorithm", C.c_int32), ("n", C.c_int32), ("batch", C.c_int32),
        ("iterations", C.c_int32), ("pump_rate_flag", C.c_int32),
        ("q", _fp), ("v", _fp),
        ("lower", C.c_double), ("upper", C.c_double),
        ("s", C.c_double), ("s_vec", _fp),
        ("pump", C.c_double), ("dt", C.c_double), ("noise_ratio", C.c_double), ("j", C.c_double),
        ("sigma", C.c_double), ("feedback_scale", C.c_double), ("g", C.c_double),
        ("alpha", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("add_assign", C.c_int32),
        ("rng_mode", C.c_int32), ("noise", _fp), ("noise_batch", C.c_int64),
        ("seed", C.c_uint64), ("offset", C.c_uint64), ("traj_base", C.c_int64),
        ("out0", _fp), ("out1", _fp), ("out2", _fp),
        ("evolution_step", C.c_int32), ("num_samples", C.c_int32), ("samples", _fp),
    ]


class EpilogueDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("batch", C.c_int32),
        ("q", _fp), ("v", _fp), ("state", _fp),
        ("apply_map1", C.c_int32), ("map1_scale", C.c_double), ("map1_shift", C.c_double),
        ("map1_scale_vec", _fp),
        ("post_processor", C.c_int32), ("pp_iterations", C.c_int32), ("pp_step", C.c_double),
        ("pp_lower", C.c_double), ("pp_upper", C.c_double),
        ("apply_map2", C.c_int32), ("map2_scale", C.c_double), ("map2_shift", C.c_double),
        ("map2_scale_vec", _fp),
        ("scaled_by", C.c_double),
        ("problem_variables", _fp), ("energy", _fp), ("scaled_by_dev", _fp),
    ]


class HookDesc(C.Structure):
    _fields_ = [
        ("solver", C.c_int32), ("kind", C.c_int32), ("n", C.c_int32), ("batch", C.c_int32),
        ("q", _fp), ("v", _fp), ("in0", _fp), ("in1", _fp), ("in2", _fp),
        ("lower", C.c_double), ("upper", C.c_double), ("s", C.c_double), ("s_vec", _fp),
        ("pump", C.c_double), ("rate", C.c_double), ("feedback_scale", C.c_double), ("j", C.c_double),
        ("g", C.c_double), ("out0", _fp), ("out1", _fp),
    ]


_lib = None


class NativeError(RuntimeError):
    """Raised when the CUDA engine reports an error (or cannot be loaded)."""


def load():
    """Load the shared library (once).  Raises NativeError with build instructions if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  ccvm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.ccvm_last_error.restype = C.c_char_p
    lib.ccvm_abi_version.restype = C.c_int
    if lib.ccvm_abi_version() != ABI_VERSION:
        raise NativeError(
            f"{LIB_PATH} has ABI version {lib.ccvm_abi_version()}, this package expects {ABI_VERSION}: the descriptor "
            "layouts differ -- rebuild it with `python -c 'import __graft_entry__ as g; g.build(force=True)'`.")
    lib.ccvm_solve.argtypes = [C.POINTER(SolveDesc), _fp]
    lib.ccvm_solve_batch.argtypes = [C.POINTER(SolveDesc), C.c_int32, _fp]
    lib.ccvm_solve_fused.argtypes = [C.POINTER(SolveDesc), C.POINTER(EpilogueDesc), C.c_double, _fp, _fp]
    lib.ccvm_solve_batch_fused.argtypes = [C.POINTER(SolveDesc), C.POINTER(EpilogueDesc), C.POINTER(C.c_double),
                                           C.c_int32, _fp, _fp]
    lib.ccvm_dump_noise.argtypes = [C.POINTER(SolveDesc), _fp, _fp]
    lib.ccvm_solution_stats_batch.argtypes = [_fp, _fp, _fp, C.c_int32, _fp, _fp]
    lib.ccvm_epilogue.argtypes = [C.POINTER(EpilogueDesc), _fp]
    lib.ccvm_compute_energy.argtypes = [_fp, _fp, _fp, C.c_double, C.c_int32, C.c_int32, _fp, _fp]
    lib.ccvm_postprocess_grad_descent.argtypes = [_fp, _fp, _fp, C.c_int32, C.c_int32, C.c_int32,
                                                  C.c_double, C.c_double, C.c_double, _fp]
    lib.ccvm_postprocess_adam.argtypes = [_fp, _fp, _fp, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                          C.c_double, _fp]
    lib.ccvm_solution_stats.argtypes = [_fp, C.c_int32, C.c_double, _fp, _fp]
    lib.ccvm_scaling_factor.argtypes = [_fp, C.c_int32, C.c_double, _fp, _fp]
    lib.ccvm_solve_host.argtypes = [C.POINTER(SolveDesc), C.POINTER(EpilogueDesc), _fp, _fp, C.c_double,
                                    _fp, _fp, _fp]
    lib.ccvm_microbench_fp32.argtypes = [C.c_int32, C.POINTER(C.c_double), _fp]
    lib.ccvm_microbench_tf32.argtypes = [C.c_int32, C.POINTER(C.c_double), _fp]
    lib.ccvm_pack_record.argtypes = [_fp, _fp, C.c_int32, C.c_int64, _fp, _fp]
    lib.ccvm_merge_records.argtypes = [_fp, C.c_int32, C.c_int32, _fp, _fp]
    lib.ccvm_query_launch.argtypes = [C.POINTER(SolveDesc), C.POINTER(C.c_int32)]
    lib.ccvm_eval_hook.argtypes = [C.POINTER(HookDesc), _fp]
    lib.ccvm_change_variables.argtypes = [_fp, _fp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double,
                                          _fp, _fp]
    lib.ccvm_fit_to_constraints.argtypes = [_fp, _fp, C.c_int32, C.c_int32, C.c_double, C.c_double, _fp, _fp,
                                            C.c_int64, _fp]
    lib.ccvm_scale_coefs.argtypes = [_fp, _fp, C.c_int32, _fp, C.c_int64, _fp, _fp, _fp]
    lib.ccvm_generate_boxqp.argtypes = [_fp, _fp, C.c_int32, C.c_uint64, C.c_double, C.c_double, _fp]
    for name in EXPORTS:
        if name not in ("ccvm_last_error",):
            getattr(lib, name).restype = C.c_int if name != "ccvm_last_error" else C.c_char_p
    lib.ccvm_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().ccvm_last_error().decode("utf-8", "replace")
        raise NativeError(f"ccvm_b200 native call failed ({rc}): {msg}")


def require_cuda():
    if not torch.cuda.is_available():
        raise NativeError("ccvm_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")


def ptr(t):
    """Device/host address of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def as_f32(t, device):
    """Contiguous fp32 view/copy of ``t`` on ``device`` (plumbing only)."""
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.float32).contiguous()
