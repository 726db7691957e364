"""ctypes view of include/wgrt.h and the loader of libwgrt.so.

There is no CPU fallback: if the CUDA library cannot be loaded every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WGRT_LIB") or os.path.join(_HERE, "libwgrt.so")   # WGRT_LIB: an experiment build
CHECKED_LIB_PATH = os.path.join(_HERE, "libwgrt_checked.so")   # -DWGRT_CHECKED build: bounds assertions in the walk

WGRT_OK = 0
WGRT_FLAG_STRICT = 0x1
WGRT_FLAG_COUNTERS = 0x2
WGRT_FLAG_BINS_ZERO = 0x4
WGRT_FLAG_BINS_DEVICE = 0x8
WGRT_FLAG_BINS_COLUMNS = 0x10
WGRT_NUM_COUNTERS = 16
COUNTER_NAMES = ("rays", "bounces", "draws", "draw2", "draw3", "efield", "iters", "deposits",
                 "poly_tests", "edge_visits", "straddle", "cross", "exact_fallback", "warp_steps", "warp_batches", "near_tie")

_f32p = C.c_void_p
_f64p = C.c_void_p


class WgrtProblem(C.Structure):
    """Mirror of ``wgrt_problem_t`` (include/wgrt.h)."""
    _fields_ = [
        ("x", _f32p), ("y", _f32p), ("gap_x", _f32p), ("gap_y", _f32p), ("pol", _f32p), ("azi", _f32p),
        ("m", _f32p), ("n", _f32p), ("lmd_num", _f32p), ("te", _f32p), ("tm", _f32p),
        ("delta_phase", _f32p), ("rng_states", C.c_void_p), ("num_rays", C.c_int64),
        ("IC", _f64p), ("IC_n", C.c_int64),
        ("FC", _f64p), ("FC_n", C.c_int64), ("FC_offset", C.c_void_p), ("n_FC", C.c_int64),
        ("OC", _f64p), ("OC_n", C.c_int64), ("OC_offset", C.c_void_p), ("n_OC", C.c_int64),
        ("n_g", C.c_double),
        ("eff_reg1", _f64p), ("eff_reg1_n", C.c_int64),
        ("eff_reg2", _f64p), ("eff_reg2_n", C.c_int64),
        ("eff_reg_FOV", _f64p), ("eff_reg_FOV_range", _f64p),
        ("lut_ic1", _f64p), ("lut_ic2", _f64p), ("lut_ic3", _f64p), ("lut_fc1", _f64p),
        ("lut_fc2", _f64p), ("lut_oc1", _f64p), ("lut_oc2", _f64p),
        ("C_ic", C.c_int32), ("C_fc", C.c_int32), ("C_oc", C.c_int32), ("reserved0", C.c_int32),
        ("lut_TIR", _f64p), ("lut_gap", _f64p),
        ("L", C.c_int64), ("X", C.c_int64), ("Y", C.c_int64),
        ("matrix_EB", C.c_void_p), ("EBy", C.c_int64), ("EBx", C.c_int64),
        ("flags", C.c_uint32), ("tile_hint", C.c_uint32),
        ("runner_points", C.c_int64), ("runner_first_cell", C.c_int64),
        ("threshold", C.c_double), ("ray_index_base", C.c_int64), ("rng_seed_offset", C.c_int64),
    ]


WGRT_EVAL_NUM = 8
EVAL_FIELDS = ("sum_de", "y_min", "y_max", "y_sum", "y_zeros", "v_max")


class WgrtEvalParams(C.Structure):
    """Mirror of ``wgrt_eval_params_t`` (include/wgrt.h)."""
    _fields_ = [("scale", C.c_double), ("white_rgb", C.c_double * 3), ("M", C.c_double * 9),
                ("M_xyz", C.c_double * 9), ("white_xyz", C.c_double * 3), ("lab_d65", C.c_double * 3)]


class WgrtLegacyProblem(C.Structure):
    """Mirror of ``wgrt_legacy_problem_t`` (include/wgrt.h): the legacy energy-splitting tracer."""
    _fields_ = [
        ("vectors", C.c_void_p), ("capacity", C.c_int64), ("useful_count_in", C.c_int64),
        ("total_ray_counter", C.c_void_p), ("max_steps", C.c_int64),
        ("IC", C.c_void_p), ("IC_n", C.c_int64),
        ("FC", C.c_void_p), ("FC_n", C.c_int64), ("FC_offset", C.c_void_p), ("n_FC", C.c_int64),
        ("OC", C.c_void_p), ("OC_n", C.c_int64), ("OC_offset", C.c_void_p), ("n_OC", C.c_int64),
        ("eff_reg1", C.c_void_p), ("eff_reg1_n", C.c_int64), ("eff_reg2", C.c_void_p), ("eff_reg2_n", C.c_int64),
        ("eff_reg_FOV", C.c_void_p), ("eff_reg_FOV_range", C.c_void_p),
        ("lut_ic1", C.c_void_p), ("lut_ic2", C.c_void_p), ("lut_fc1", C.c_void_p), ("lut_fc2", C.c_void_p),
        ("lut_oc", C.c_void_p),
        ("C_ic", C.c_int32), ("C_fc", C.c_int32), ("C_oc", C.c_int32), ("reserved0", C.c_int32),
        ("lut_TIR", C.c_void_p), ("lut_gap", C.c_void_p), ("X", C.c_int64), ("Y", C.c_int64),
        ("matrix_EB", C.c_void_p), ("EBy", C.c_int64), ("EBx", C.c_int64),
    ]


LEGACY_COLS = 13


class WgrtError(RuntimeError):
    pass


class WgrtInvalidArgument(WgrtError, ValueError):
    """WGRT_ERR_INVALID: an argument the library rejected (also a ``ValueError``, which is what the Python
    layer raises for the checks it can do itself)."""


_lib: Optional[C.CDLL] = None

# every symbol include/wgrt.h declares
EXPORTED_SYMBOLS = (
    "wgrt_version", "wgrt_problem_size", "wgrt_last_error", "wgrt_device_count", "wgrt_release",
    "wgrt_trace_fullcolor", "wgrt_trace_fullcolor_host", "wgrt_trace_evaluate_host",
    "wgrt_seed_rng", "wgrt_counters_read", "wgrt_counters_reset",
    "wgrt_debug_locate", "wgrt_debug_efield", "wgrt_debug_xorshift", "wgrt_debug_fma_peak",
    "wgrt_debug_deposit_inside", "wgrt_debug_set_tie_tolerance", "wgrt_debug_check_failures",
    "wgrt_eval_pupil_sums", "wgrt_eval_pupil_sums_host", "wgrt_eval_metrics", "wgrt_eval_metrics_host",
    "wgrt_trace_evaluate_metrics_host",
    "wgrt_legacy_problem_size", "wgrt_legacy_step", "wgrt_legacy_pack_active", "wgrt_legacy_trace_host", "wgrt_bins_pack_u8", "wgrt_bins_unpack_u8",
)


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load libwgrt.so (built in-tree by ``__graft_entry__.build()`` / ``csrc/build.py``)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise WgrtError(
            f"{p} not found: the CUDA engine is not built. Run `python -m "
            "gpu_ray_tracing_for_waveguide_based_ar_display_b200.csrc.build` "
            "(needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(p)
    lib.wgrt_version.restype = C.c_int
    lib.wgrt_problem_size.restype = C.c_int
    if lib.wgrt_problem_size() != C.sizeof(WgrtProblem):
        raise WgrtError("wgrt_problem_t layout mismatch between _capi.py and libwgrt.so")
    lib.wgrt_last_error.restype = C.c_char_p
    lib.wgrt_device_count.restype = C.c_int
    lib.wgrt_release.restype = C.c_int
    lib.wgrt_trace_fullcolor.restype = C.c_int
    lib.wgrt_trace_fullcolor.argtypes = [C.POINTER(WgrtProblem), C.c_void_p]
    lib.wgrt_trace_fullcolor_host.restype = C.c_int
    lib.wgrt_trace_fullcolor_host.argtypes = [C.POINTER(WgrtProblem), C.c_int, C.c_void_p]
    lib.wgrt_bins_pack_u8.restype = C.c_int
    lib.wgrt_bins_pack_u8.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    lib.wgrt_bins_unpack_u8.restype = C.c_int
    lib.wgrt_bins_unpack_u8.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.wgrt_trace_evaluate_host.restype = C.c_int
    lib.wgrt_trace_evaluate_host.argtypes = [C.POINTER(WgrtProblem), C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wgrt_seed_rng.restype = C.c_int
    lib.wgrt_seed_rng.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.wgrt_counters_read.restype = C.c_int
    lib.wgrt_counters_read.argtypes = [C.c_void_p, C.c_int]
    lib.wgrt_counters_reset.restype = C.c_int
    lib.wgrt_debug_locate.restype = C.c_int
    lib.wgrt_debug_locate.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    lib.wgrt_debug_efield.restype = C.c_int
    lib.wgrt_debug_efield.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_void_p]
    lib.wgrt_debug_xorshift.restype = C.c_int
    lib.wgrt_debug_xorshift.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    lib.wgrt_debug_deposit_inside.restype = C.c_int
    lib.wgrt_debug_deposit_inside.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    lib.wgrt_debug_check_failures.restype = C.c_int
    lib.wgrt_debug_check_failures.argtypes = [C.c_void_p, C.c_int]
    lib.wgrt_debug_set_tie_tolerance.restype = C.c_int
    lib.wgrt_debug_set_tie_tolerance.argtypes = [C.c_double]
    lib.wgrt_debug_fma_peak.restype = C.c_int
    lib.wgrt_debug_fma_peak.argtypes = [C.c_void_p, C.c_void_p]
    lib.wgrt_eval_pupil_sums.restype = C.c_int
    lib.wgrt_eval_pupil_sums.argtypes = [C.c_void_p] + [C.c_int64] * 5 + [C.c_int] * 3 + \
        [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wgrt_eval_pupil_sums_host.restype = C.c_int
    lib.wgrt_eval_pupil_sums_host.argtypes = [C.c_void_p] + [C.c_int64] * 5 + [C.c_int] * 3 + \
        [C.c_void_p, C.c_void_p]
    lib.wgrt_eval_metrics.restype = C.c_int
    lib.wgrt_eval_metrics.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(WgrtEvalParams),
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wgrt_eval_metrics_host.restype = C.c_int
    lib.wgrt_eval_metrics_host.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(WgrtEvalParams),
                                           C.c_void_p, C.c_void_p]
    lib.wgrt_trace_evaluate_metrics_host.restype = C.c_int
    lib.wgrt_trace_evaluate_metrics_host.argtypes = [C.POINTER(WgrtProblem), C.c_int, C.c_int, C.c_int, C.c_int,
                                                     C.POINTER(WgrtEvalParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                                     C.c_void_p, C.c_void_p]
    lib.wgrt_legacy_problem_size.restype = C.c_int
    if lib.wgrt_legacy_problem_size() != C.sizeof(WgrtLegacyProblem):
        raise WgrtError("wgrt_legacy_problem_t layout mismatch between _capi.py and libwgrt.so")
    lib.wgrt_legacy_step.restype = C.c_int
    lib.wgrt_legacy_step.argtypes = [C.POINTER(WgrtLegacyProblem), C.c_void_p]
    lib.wgrt_legacy_pack_active.restype = C.c_int
    lib.wgrt_legacy_pack_active.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.wgrt_legacy_trace_host.restype = C.c_int
    lib.wgrt_legacy_trace_host.argtypes = [C.POINTER(WgrtLegacyProblem), C.c_int, C.c_void_p]
    if path is None:
        _lib = lib
    return lib


def check(rc: int, lib: Optional[C.CDLL] = None) -> None:
    if rc != WGRT_OK:
        lib = lib or load_library()
        msg = lib.wgrt_last_error()
        cls = WgrtInvalidArgument if rc == -1 else WgrtError
        raise cls(f"libwgrt error {rc}: {msg.decode() if msg else '?'}")


def np_ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def read_counters() -> dict:
    lib = load_library()
    buf = np.zeros(WGRT_NUM_COUNTERS, dtype=np.uint64)
    check(lib.wgrt_counters_read(np_ptr(buf), WGRT_NUM_COUNTERS), lib)
    return {k: int(buf[i]) for i, k in enumerate(COUNTER_NAMES)}


def reset_counters() -> None:
    lib = load_library()
    check(lib.wgrt_counters_reset(), lib)
