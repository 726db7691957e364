"""Runner-level entry: the device section of gpu_ray_tracing_pro_fullColor.py in one call.

``trace_full_color`` does what the reference runner does between building its inputs and reading
the bins back (gpu_ray_tracing_pro_fullColor.py:59-185) -- but instead of materialising twelve
``num_rays``-long host arrays from ``num_rays_per_FoV/2`` start points (RUN:62-115) and copying
5.4 GB to the device (RUN:145-158), it hands the start points to the engine, which derives every ray
from the runner's layout rule on the fly (include/wgrt.h, "runner layout").  The bins are
bit-identical to launching on the materialised arrays (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np

from . import _capi
from .GPU_ray_tracing_functions import pack_problem

__all__ = ["trace_full_color"]


def trace_full_color(points: np.ndarray, geom: Dict[str, np.ndarray], n_g: float, luts: Dict[str, np.ndarray],
                     num_rays_per_FoV: int, num_iter: int = 4, eb: Tuple[int, int] = (80, 120),
                     first_cell: int = 0, num_cells: Optional[int] = None,
                     matrix_EB: Optional[np.ndarray] = None, rng_states: Optional[np.ndarray] = None,
                     flags: int = 0, timings: Optional[list] = None, bins_start_zero: Optional[bool] = None
                     ) -> np.ndarray:
    """Trace ``num_iter`` launches of the full-colour walk over the runner's ray layout.

    points      [num_rays_per_FoV/2, 2] start points (``generate_points_in_polygon`` output)
    geom, luts  the design arrays / RCWA tables under the runner's names (IC, FC, FC_offset, ...,
                lut_TIR, lut_gap; lut_ic1 ... lut_oc2)
    first_cell, num_cells   sub-range of the runner's cell sequence (multi-GPU sharding); default all
    matrix_EB   optional preallocated float32 [L, Y, X, EBy, EBx] host array.  When omitted a
                zeroed one is created and never uploaded (it is cleared on the device);
                ``bins_start_zero=True`` declares a caller-provided array to be all zero, too.
    rng_states  optional uint32 host array for this cell range; when omitted the states are seeded
                on the device as the runner seeds them (RUN:158) and discarded.
    Returns the bin tensor (host).
    """
    lib = _capi.load_library()
    P = num_rays_per_FoV // 2
    if points.shape != (P, 2) or 2 * P != num_rays_per_FoV:
        raise ValueError("points must be [num_rays_per_FoV/2, 2]")
    L, X, Y, _ = geom["lut_TIR"].shape
    if num_cells is None:
        num_cells = L * X * Y - first_cell
    N = num_cells * num_rays_per_FoV
    px = np.ascontiguousarray(points[:, 0], dtype=np.float32)   # the runner stores float32 (RUN:65-66, 88-89)
    py = np.ascontiguousarray(points[:, 1], dtype=np.float32)
    if matrix_EB is None:
        matrix_EB = np.zeros((L, Y, X, eb[0], eb[1]), dtype=np.float32)
        bins_start_zero = True
    if bins_start_zero:
        flags |= _capi.WGRT_FLAG_BINS_ZERO
    args = (px, py, None, None, None, None, None, None, None, None, None, None, rng_states,
            geom["IC"], geom["FC"], geom["FC_offset"], geom["OC"], geom["OC_offset"], float(n_g),
            geom["eff_reg1"], geom["eff_reg2"], geom["eff_reg_FOV"], geom["eff_reg_FOV_range"],
            luts["lut_ic1"], luts["lut_ic2"], luts["lut_ic3"], luts["lut_fc1"], luts["lut_fc2"],
            luts["lut_oc1"], luts["lut_oc2"], geom["lut_TIR"], geom["lut_gap"], matrix_EB)
    prob, keep = pack_problem(args, host=True, flags=flags, runner_points=P, runner_first_cell=first_cell,
                              num_rays=N)
    tms = (C.c_float * 3)()
    _capi.check(lib.wgrt_trace_fullcolor_host(C.byref(prob), int(num_iter), tms), lib)
    if timings is not None:
        timings[:] = list(tms)
    del keep
    return matrix_EB
