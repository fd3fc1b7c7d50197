"""ctypes binding of the C ABI declared in include/rectipy_b200.h.

The shared library is built in-tree (rectipy_b200/csrc/librectipy_b200.so, see __graft_entry__.build or
`make -C rectipy_b200/csrc`).  There is no CPU or eager-PyTorch fallback: if the library is missing, or no
sm_100 device is present when a plan is created, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

RP_ABI_VERSION = 8
RP_MAX_IN, RP_MAX_OUT, RP_MAX_SV, RP_MAX_REC = 8, 8, 4, 4
RP_LI_TANH, RP_LI_SIGMOID, RP_QIF, RP_QIF_SFA, RP_LIF, RP_IK, RP_IKU, RP_IK_BIEXP, RP_JIT = range(9)
(RP_P_TAU, RP_P_K, RP_P_ETA, RP_P_TAU_S, RP_P_TAU_X, RP_P_ALPHA, RP_P_RMAX, RP_P_SIG_S, RP_P_V0,
 RP_P_C, RP_P_VR, RP_P_VTH, RP_P_G, RP_P_ER, RP_P_B, RP_P_TAU_U, RP_P_KAPPA, RP_NUM_PARAMS) = range(18)
RP_IN_NONE, RP_IN_DENSE, RP_IN_PROJ = range(3)
RP_OUT_DENSE, RP_OUT_READOUT = range(2)
RP_VAR_V, RP_VAR_S, RP_VAR_X, RP_VAR_R = range(4)
RP_PREC_FP32, RP_PREC_3XTF32, RP_PREC_3XF16 = range(3)
RP_NUM_STAGES = 5
STAGE_NAMES = ("fwd_fused", "dgrad", "wgrad", "adjoint_elementwise", "other")

_fp = C.c_void_p   # device pointers travel as plain integers


class rp_desc(C.Structure):
    _fields_ = [("model", C.c_int), ("n", C.c_int), ("batch", C.c_int), ("in_mode", C.c_int), ("n_in", C.c_int),
                ("in_target", C.c_int), ("out_mode", C.c_int), ("n_out", C.c_int), ("out_var", C.c_int),
                ("precision", C.c_int), ("dt", C.c_float), ("theta", C.c_float), ("v_reset", C.c_float),
                ("slope", C.c_float), ("param_per_neuron", C.c_int * RP_NUM_PARAMS),
                ("jit_nsv", C.c_int), ("jit_spiking", C.c_int), ("jit_post_out", C.c_int), ("jit_src_plane", C.c_int)]


class rp_fwd_args(C.Structure):
    _fields_ = [("T", C.c_int), ("sampling_steps", C.c_int), ("cutoff", C.c_int),
                ("x", _fp), ("W", _fp), ("W_in", _fp), ("W_out", _fp), ("params", _fp * RP_NUM_PARAMS),
                ("y0", _fp), ("yT", _fp), ("out_rec", _fp),
                ("n_rec_vars", C.c_int), ("rec_var", C.c_int * RP_MAX_REC), ("rec_reduce", C.c_int * RP_MAX_REC),
                ("rec_buf", _fp * RP_MAX_REC), ("history", _fp), ("t_offset", C.c_int), ("T_total", C.c_int)]


class rp_bwd_args(C.Structure):
    _fields_ = [("T", C.c_int), ("sampling_steps", C.c_int), ("cutoff", C.c_int), ("truncate_steps", C.c_int),
                ("x", _fp), ("W", _fp), ("W_in", _fp), ("W_out", _fp), ("params", _fp * RP_NUM_PARAMS),
                ("history", _fp), ("g_out_rec", _fp), ("g_yT", _fp),
                ("dW", _fp), ("dW_in", _fp), ("dW_out", _fp), ("dparams", _fp * RP_NUM_PARAMS),
                ("g_y0", _fp), ("g_x", _fp), ("t_offset", C.c_int), ("T_total", C.c_int)]


#: every symbol include/rectipy_b200.h declares (checked by tests/test_cabi.py)
EXPORTS = ["rp_abi_version", "rp_last_error", "rp_num_state_vars", "rp_num_history_planes", "rp_num_records", "rp_plan_create",
           "rp_plan_destroy", "rp_plan_workspace_bytes", "rp_plan_launch_count", "rp_forward", "rp_backward", "rp_plan_status",
           "rp_rls_run", "rp_gemm_tn", "rp_plan_time_contraction", "rp_trace_enable", "rp_trace_read", "rp_plan_stage_timing",
           "rp_plan_stage_times", "rp_plan_set_jit_module", "rp_plan_path"]

_LIB = None


def library_path() -> str:
    if os.environ.get("RECTIPY_B200_LIB"):        # A/B builds of the same ABI (kernel tuning experiments)
        return os.environ["RECTIPY_B200_LIB"]
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "librectipy_b200.so")


def load():
    """Load the engine library; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"rectipy_b200: CUDA library not built ({path} missing). Run `python -c 'import "
                           f"__graft_entry__ as g; g.build()'` or `make -C rectipy_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.rp_abi_version.restype = C.c_int
    lib.rp_last_error.restype = C.c_char_p
    lib.rp_num_state_vars.argtypes = [C.c_int]
    lib.rp_num_state_vars.restype = C.c_int
    lib.rp_num_history_planes.argtypes = [C.c_int]
    lib.rp_num_history_planes.restype = C.c_int
    lib.rp_num_records.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.rp_num_records.restype = C.c_int
    lib.rp_plan_create.argtypes = [C.POINTER(rp_desc), C.POINTER(C.c_void_p)]
    lib.rp_plan_create.restype = C.c_int
    lib.rp_plan_path.argtypes = [C.POINTER(rp_desc)]
    lib.rp_plan_path.restype = C.c_int
    lib.rp_plan_destroy.argtypes = [C.c_void_p]
    lib.rp_plan_destroy.restype = None
    lib.rp_plan_workspace_bytes.argtypes = [C.c_void_p]
    lib.rp_plan_workspace_bytes.restype = C.c_longlong
    lib.rp_plan_launch_count.argtypes = [C.c_void_p]
    lib.rp_plan_launch_count.restype = C.c_longlong
    lib.rp_forward.argtypes = [C.c_void_p, C.POINTER(rp_fwd_args), C.c_void_p]
    lib.rp_forward.restype = C.c_int
    lib.rp_backward.argtypes = [C.c_void_p, C.POINTER(rp_bwd_args), C.c_void_p]
    lib.rp_backward.restype = C.c_int
    lib.rp_plan_status.argtypes = [C.c_void_p, C.c_void_p]
    lib.rp_plan_status.restype = C.c_int
    lib.rp_trace_enable.argtypes = [C.c_int]
    lib.rp_trace_enable.restype = C.c_int
    lib.rp_trace_read.argtypes = [C.c_void_p, C.c_int]
    lib.rp_trace_read.restype = C.c_int
    lib.rp_rls_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_void_p]
    lib.rp_rls_run.restype = C.c_int
    lib.rp_plan_time_contraction.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_void_p]
    lib.rp_plan_time_contraction.restype = C.c_int
    lib.rp_plan_stage_timing.argtypes = [C.c_void_p, C.c_int]
    lib.rp_plan_stage_timing.restype = C.c_int
    lib.rp_plan_stage_times.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p]
    lib.rp_plan_stage_times.restype = C.c_int
    lib.rp_plan_set_jit_module.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
    lib.rp_plan_set_jit_module.restype = C.c_int
    lib.rp_gemm_tn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, C.c_int, C.c_int, C.c_void_p]
    lib.rp_gemm_tn.restype = C.c_int
    if lib.rp_abi_version() != RP_ABI_VERSION:
        raise RuntimeError("rectipy_b200: ABI version mismatch between _cabi.py and librectipy_b200.so; rebuild")
    _LIB = lib
    return lib


def last_error() -> str:
    return load().rp_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what}: {last_error()}")
