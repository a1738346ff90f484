"""ctypes binding of libmppi_b200.so (include/mppi_b200.h).

There is no Python or CPU implementation behind this module: if the CUDA
library is missing or cannot be loaded, importing the engine fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmppi_b200.so")

ABI_VERSION = 1
MAX_DU, MAX_SG, MAX_PARAMS = 4, 33, 32

(MODEL_PENDULUM, MODEL_CARTPOLE, MODEL_MOUNTAINCAR, MODEL_NAVIGATION2D, MODEL_RACING, MODEL_CARTPOLE_CONTINUOUS,
 MODEL_GOAL_IN_DANGER_ZONE) = range(7)
LAMBDA_FIXED, LAMBDA_MPO, LAMBDA_LBPS, LAMBDA_ESSPS = range(4)
CFG_FORCE_PAIRED, CFG_FORCE_SINGLE = 1, 2  # MppiConfig.flags
NAV2D_NUM_PARAMS, RACING_NUM_PARAMS, GOAL_ZONE_NUM_PARAMS = 12, 17, 11


class MppiConfig(C.Structure):
    """struct MppiConfig of include/mppi_b200.h, field for field."""

    _fields_ = [
        ("abi_version", C.c_int32), ("model", C.c_int32), ("horizon", C.c_int32), ("num_samples", C.c_int32),
        ("dim_state", C.c_int32), ("dim_control", C.c_int32),
        ("u_min", C.c_float * MAX_DU), ("u_max", C.c_float * MAX_DU), ("sigmas", C.c_float * MAX_DU),
        ("lambda_mode", C.c_int32), ("lambda_", C.c_double), ("lbps_delta", C.c_double),
        ("essps_target_ess", C.c_double), ("lambda_min", C.c_double), ("lambda_max", C.c_double),
        ("exploration", C.c_double),
        ("use_sg_filter", C.c_int32), ("sg_window_size", C.c_int32), ("sg_poly_order", C.c_int32),
        ("sg_coeffs_given", C.c_int32), ("sg_coeffs", C.c_float * MAX_SG),
        ("seed", C.c_uint64), ("device", C.c_int32),
        ("sample_offset", C.c_int64), ("total_samples", C.c_int64),
        ("num_model_params", C.c_int32), ("model_params", C.c_float * MAX_PARAMS),
        ("block_size", C.c_int32), ("flags", C.c_int32),
    ]


class MppiFp32Report(C.Structure):
    """struct MppiFp32Report of include/mppi_b200.h (csrc/mppi_microbench.cu)."""

    _fields_ = [(n, C.c_double) for n in ("ffma_tflops", "ffma2_tflops", "fmul_fadd_tflops", "fmul2_fadd2_tflops",
                                          "lat_ffma", "lat_ffma2", "lat_fadd2", "lat_fmnmx", "lat_fadd", "lat_fmul2",
                                          "sm_clock_mhz")] + [("sms", C.c_int32), ("reserved", C.c_int32)]


_P = C.c_void_p
_FP = C.c_void_p  # float* passed as raw address
RASTER_OBSTACLE, RASTER_LANE = 0, 1
TOP_MAX = 1024  # largest top_n of the select path (mppi_step_epilogue / mppi_top_candidates)


class MppiStepEpilogue(C.Structure):
    """struct MppiStepEpilogue of include/mppi_b200.h, field for field."""

    _fields_ = [
        ("d_state", _FP), ("d_action_seq", _FP), ("d_state_seq", _FP),
        ("goal_x", C.c_float), ("goal_y", C.c_float), ("goal_threshold", C.c_float),
        ("top_n", C.c_int32),
        ("d_cand_cost", _FP), ("d_cand_id", _FP), ("n_cand", C.c_int32),
        ("d_noise_global", _FP),
        ("d_next_state", _FP), ("d_flags", _FP), ("d_top_traj", _FP), ("d_top_w", _FP), ("d_top_cost", _FP),
        ("d_top_id", _FP),
    ]


# name -> (restype, argtypes); every symbol include/mppi_b200.h declares
PROTOTYPES = {
    "mppi_create": (C.c_int, [C.POINTER(MppiConfig), C.POINTER(_P)]),
    "mppi_destroy": (None, [_P]),
    "mppi_reset": (C.c_int, [_P, _P]),
    "mppi_last_error": (C.c_char_p, []),
    "mppi_abi_version": (C.c_int, []),
    "mppi_set_model_params": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int32]),
    "mppi_set_map": (C.c_int, [_P, C.c_int32, _FP, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                               C.c_float]),
    "mppi_raster_map": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float,
                                  C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32), C.c_int32, _FP]),
    "mppi_solve": (C.c_int, [_P, _FP, _FP, _FP, _FP, _FP, _P]),
    "mppi_solve_host": (C.c_int, [_P, _FP, _FP, _FP, _FP]),
    "mppi_shard_rollout": (C.c_int, [_P, _FP, _FP, _FP, _P]),
    "mppi_shard_lambda": (C.c_int, [_P, _FP, _P]),
    "mppi_shard_finish": (C.c_int, [_P, _FP, C.c_int32, _FP, _FP, _FP, _P]),
    "mppi_partial_floats": (C.c_int32, [_P]),
    "mppi_p2p_export": (C.c_int, [_P, C.POINTER(C.c_uint8)]),
    "mppi_p2p_connect": (C.c_int, [_P, C.POINTER(C.c_uint8), C.c_int32, C.c_int32]),
    "mppi_p2p_status": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "mppi_p2p_barrier": (C.c_int, [_P, _P]),
    "mppi_p2p_mailbox_ptr": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "mppi_p2p_connect_local": (C.c_int, [_P, C.POINTER(C.c_uint64), C.c_int32, C.c_int32]),
    "mppi_costs_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "mppi_partial_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "mppi_weights": (C.c_int, [_P, _FP, _P]),
    "mppi_top_samples": (C.c_int, [_P, C.c_int32, _FP, _FP, _P]),
    "mppi_step_epilogue": (C.c_int, [_P, C.POINTER(MppiStepEpilogue), _P]),
    "mppi_top_candidates": (C.c_int, [_P, C.c_int32, _FP, _FP, _P]),
    "mppi_last_epilogue_launches": (C.c_int32, [_P]),
    "mppi_rollout_actions": (C.c_int, [_P, _FP, _FP, C.c_int32, _FP, _P]),
    "mppi_get_lambda": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), _P]),
    "mppi_get_carry": (C.c_int, [_P, _FP, _FP, _P]),
    "mppi_set_carry": (C.c_int, [_P, _FP, _FP, _P]),
    "mppi_prev_action_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "mppi_last_launch_count": (C.c_int32, [_P]),
    "mppi_launch_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mppi_map_info": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]),
    "mppi_block_trace": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_uint64), C.c_int32]),
    "mppi_selftest": (C.c_int, [C.c_int32, C.POINTER(C.c_uint64)]),
    "mppi_refpath_create": (C.c_int, [C.c_int32, _FP, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_float,
                                      C.POINTER(_P)]),
    "mppi_refpath_destroy": (None, [_P]),
    "mppi_refpath_update": (C.c_int, [_P, _FP, _FP, _P]),
    "mppi_refpath_index": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), _P]),
    "mppi_kernel_timing": (C.c_int, [_P, C.c_int32]),
    "mppi_kernel_time_ms": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "mppi_fp32_microbench": (C.c_int, [C.c_int32, C.POINTER(MppiFp32Report)]),
    "mppi_philox4x32_10": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "mppi_solve_index": (C.c_uint64, [_P]),
    "mppi_sample_noise": (C.c_int, [_P, C.c_uint64, _FP, _P]),
}

_lib = None


class EngineError(RuntimeError):
    pass


def load():
    """Load libmppi_b200.so once. Raises (never falls back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C mppi_playground_b200/csrc`. mppi_playground_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.mppi_abi_version() != ABI_VERSION:
        raise EngineError(f"libmppi_b200.so ABI {lib.mppi_abi_version()} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().mppi_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise EngineError(f"libmppi_b200 error {rc}: {msg}")
