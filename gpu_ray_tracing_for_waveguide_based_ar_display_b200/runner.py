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
from .GPU_ray_tracing_functions import _as_complex128, pack_problem

__all__ = ["trace_full_color", "trace_and_evaluate"]


def trace_full_color(points: np.ndarray, geom: Dict[str, np.ndarray], n_g: float, luts: Dict[str, np.ndarray],
                     num_rays_per_FoV: int, num_iter: int = 4, eb: Tuple[int, int] = (80, 120),
                     first_cell: int = 0, num_cells: Optional[int] = None,
                     matrix_EB: Optional[np.ndarray] = None, rng_states: Optional[np.ndarray] = None,
                     flags: int = 0, timings: Optional[list] = None, bins_start_zero: Optional[bool] = None,
                     rng_seed_offset: int = 0) -> np.ndarray:
    """Trace ``num_iter`` launches of the full-colour walk over the runner's ray layout.

    points      [num_rays_per_FoV/2, 2] start points (``generate_points_in_polygon`` output)
    geom, luts  the design arrays / RCWA tables under the runner's names (IC, FC, FC_offset, ...,
                lut_TIR, lut_gap; lut_ic1 ... lut_oc2)
    first_cell, num_cells   sub-range of the runner's cell sequence (multi-GPU sharding); default all
    matrix_EB   optional preallocated float32 [L, Y, X, EBy, EBx] host array.  When omitted a
                zeroed one is created and never uploaded (it is cleared on the device);
                ``bins_start_zero=True`` declares a caller-provided array to be all zero, too.
                A DEVICE buffer (anything with ``__cuda_array_interface__``) is used in place: the
                launches accumulate into it and nothing is downloaded (multi-GPU jobs reduce the bins
                over NVLink before anybody reads them).
    rng_states  optional uint32 host array for this cell range; when omitted the states are seeded
                on the device as the runner seeds them (RUN:158) and discarded.  ``rng_seed_offset``
                shifts the ray index of that seeding rule (independent streams for replicated jobs).
    Returns the bin tensor (host).
    """
    lib = _capi.load_library()
    luts = {k: _as_complex128(v, k) for k, v in luts.items()}
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
    if not isinstance(matrix_EB, np.ndarray) and hasattr(matrix_EB, "__cuda_array_interface__"):
        flags |= _capi.WGRT_FLAG_BINS_DEVICE
    args = (px, py, None, None, None, None, None, None, None, None, None, None, rng_states,
            geom["IC"], geom["FC"], geom["FC_offset"], geom["OC"], geom["OC_offset"], float(n_g),
            geom["eff_reg1"], geom["eff_reg2"], geom["eff_reg_FOV"], geom["eff_reg_FOV_range"],
            luts["lut_ic1"], luts["lut_ic2"], luts["lut_ic3"], luts["lut_fc1"], luts["lut_fc2"],
            luts["lut_oc1"], luts["lut_oc2"], geom["lut_TIR"], geom["lut_gap"], matrix_EB)
    # ray 0 of this call is ray first_cell * num_rays_per_FoV of the runner's whole job (only the
    # reference's zero-state reseed rule, GRTF:28-29, reads a ray's index)
    prob, keep = pack_problem(args, host=True, flags=flags, runner_points=P, runner_first_cell=first_cell,
                              num_rays=N, ray_index_base=first_cell * num_rays_per_FoV,
                              rng_seed_offset=rng_seed_offset)
    tms = (C.c_float * 3)()
    _capi.check(lib.wgrt_trace_fullcolor_host(C.byref(prob), int(num_iter), tms), lib)
    if timings is not None:
        timings[:] = list(tms)
    del keep
    return matrix_EB


def trace_and_evaluate(points: np.ndarray, geom: Dict[str, np.ndarray], n_g: float, luts: Dict[str, np.ndarray],
                       num_rays_per_FoV: int, num_iter: int = 4, eb: Tuple[int, int] = (80, 120),
                       mask_size: int = 30, step_y: int = 8, step_x: int = 12, flags: int = 0,
                       timings: Optional[list] = None, full_evaluation: bool = True, return_image: bool = False,
                       return_perceive: bool = False, host_evaluation: bool = False) -> dict:
    """The runner from its inputs to its evaluation (gpu_ray_tracing_pro_fullColor.py:59-198) with the
    bin tensor never leaving the device.

    Walks ``num_iter`` launches over the runner's ray layout, reduces the pupil-mask sums
    (AR_system_evaluation_functions.py:68-109), the per-cell totals (RUN:186) and -- with ``full_evaluation``
    -- the rest of ``evaluation()`` (reference lines 110-160: display model, Lab / CIEDE2000, luminance
    statistics per eye position) on the device, and downloads only the per-cell totals (90 KB at the default
    size) and 8 doubles per eye position (instead of the 864 MB bin tensor).  Returns a dict with ``efficiency``
    (RUN:186-192, per wavelength), ``cell_sums`` and ``delta_e``, ``U_fov``, ``U_EB`` as the reference's
    ``evaluation()`` returns them; ``output_image`` (5 MB) and ``matrix_eye_perceive`` (normalised as RUN:197
    does, 5 MB) only when ``return_image`` / ``return_perceive`` ask for them.  ``host_evaluation=True`` finishes
    ``evaluation()`` with the NumPy mirror on the host instead (the round-1 path; needs the pupil sums).
    """
    from . import AR_system_evaluation_functions as EV
    lib = _capi.load_library()
    luts = {k: _as_complex128(v, k) for k, v in luts.items()}
    P = num_rays_per_FoV // 2
    if points.shape != (P, 2) or 2 * P != num_rays_per_FoV:
        raise ValueError("points must be [num_rays_per_FoV/2, 2]")
    L, X, Y, _ = geom["lut_TIR"].shape
    N = L * X * Y * num_rays_per_FoV
    px = np.ascontiguousarray(points[:, 0], dtype=np.float32)
    py = np.ascontiguousarray(points[:, 1], dtype=np.float32)
    n_epy = (eb[0] - mask_size) // step_y + 1 if eb[0] >= mask_size else 0
    n_epx = (eb[1] - mask_size) // step_x + 1 if eb[1] >= mask_size else 0
    device_eval = full_evaluation and not host_evaluation
    want_perceive = return_perceive or not device_eval
    perceive = np.zeros((L, Y, X, n_epy, n_epx), dtype=np.float32) if want_perceive else None
    cells = np.zeros((L, Y, X), dtype=np.float32)
    args = (px, py, None, None, None, None, None, None, None, None, None, None, None,
            geom["IC"], geom["FC"], geom["FC_offset"], geom["OC"], geom["OC_offset"], float(n_g),
            geom["eff_reg1"], geom["eff_reg2"], geom["eff_reg_FOV"], geom["eff_reg_FOV_range"],
            luts["lut_ic1"], luts["lut_ic2"], luts["lut_ic3"], luts["lut_fc1"], luts["lut_fc2"],
            luts["lut_oc1"], luts["lut_oc2"], geom["lut_TIR"], geom["lut_gap"], None)
    prob, keep = pack_problem(args, host=True, flags=flags, runner_points=P, runner_first_cell=0, num_rays=N, eb=eb)
    tms = (C.c_float * 3)()
    if device_eval:
        metrics = np.zeros((n_epy * n_epx, _capi.WGRT_EVAL_NUM), dtype=np.float64)
        image = np.zeros((Y, X, 3, n_epy, n_epx), dtype=np.float32) if return_image else None
        prm = EV.eval_params(1.0 / (float(num_rays_per_FoV) * float(num_iter)))
        _capi.check(lib.wgrt_trace_evaluate_metrics_host(
            C.byref(prob), int(num_iter), mask_size, step_y, step_x, C.byref(prm), metrics.ctypes.data,
            cells.ctypes.data, perceive.ctypes.data if want_perceive else None,
            image.ctypes.data if return_image else None, tms), lib)
    else:
        _capi.check(lib.wgrt_trace_evaluate_host(C.byref(prob), int(num_iter), mask_size, step_y, step_x,
                                                 perceive.ctypes.data, cells.ctypes.data, tms), lib)
    if timings is not None:
        timings[:] = list(tms)
    del keep
    # RUN:186-192 with num_rays = all rays of one launch
    out = {"efficiency": cells.astype(np.float64).sum(axis=(1, 2)) / N / num_iter * 3, "cell_sums": cells}
    if want_perceive:
        out["matrix_eye_perceive"] = perceive / np.float32(num_rays_per_FoV) / np.float32(num_iter)   # RUN:197 is linear
    if device_eval:
        out["delta_e"], out["U_fov"], out["U_EB"] = EV.finish_metrics(metrics, Y * X, n_epy, n_epx)
        out["eval_metrics"] = metrics
        if return_image:
            out["output_image"] = image
    elif full_evaluation:
        shape_only = np.broadcast_to(np.float32(0), (L, Y, X, eb[0], eb[1]))
        out["delta_e"], out["U_fov"], out["U_EB"], out["output_image"] = EV.evaluation(
            shape_only, matrix_eye_perceive=out["matrix_eye_perceive"])
    return out


def main(argv=None) -> int:
    """``python -m gpu_ray_tracing_for_waveguide_based_ar_display_b200.runner`` -- the reference runner script
    (gpu_ray_tracing_pro_fullColor.py:11-210) end to end on this engine: design -> LUTs -> start points ->
    ``num_iter`` launches -> efficiencies and evaluation, printed like the reference prints them.  The RCWA
    LUT files of the reference are not redistributable / reachable, so synthetic LUTs of the same shapes
    stand in unless ``--lut-dir`` points at a directory holding the reference's seven ``lut_*_fullColor.npy``."""
    import argparse
    import os
    import time
    from . import GPU_ray_tracing_functions as GRTF
    from . import synthetic_inputs as si
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("--fov", type=int, nargs=2, default=(100, 75), metavar=("X", "Y"))      # RUN:16-17
    ap.add_argument("--rays-per-fov", type=int, default=5000)                               # RUN:61
    ap.add_argument("--num-iter", type=int, default=4)                                      # RUN:60
    ap.add_argument("--eyebox-bins", type=int, nargs=2, default=(80, 120), metavar=("Y", "X"))  # RUN:37
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--lut-dir", default=None)
    a = ap.parse_args(argv)
    print("=" * 60 + f"\n{'GPU Ray Tracing Simulation (B200 engine)':^60}\n" + "=" * 60)
    scene = si.make_scene(a.fov[0], a.fov[1], 2, eb=tuple(a.eyebox_bins), seed=a.seed, build_rays=False)
    luts = scene.luts
    if a.lut_dir:
        luts = {k: np.ascontiguousarray(np.load(os.path.join(a.lut_dir, f"{k}_fullColor.npy")).astype(np.complex128))
                for k in ("lut_ic1", "lut_ic2", "lut_ic3", "lut_fc1", "lut_fc2", "lut_oc1", "lut_oc2")}
    points = GRTF.generate_points_in_polygon(scene.geom["IC"], a.rays_per_fov // 2, rng=np.random.default_rng(a.seed))
    num_rays = a.fov[0] * a.fov[1] * 3 * a.rays_per_fov
    print(f"Rays per launch : {num_rays:,}   launches : {a.num_iter}")
    t0 = time.perf_counter()
    res = trace_and_evaluate(points, scene.geom, scene.n_g, luts, a.rays_per_fov, num_iter=a.num_iter,
                             eb=tuple(a.eyebox_bins))
    dt = time.perf_counter() - t0
    print(f"Number of rays traced : {num_rays * a.num_iter:,}")
    print(f"Trace + evaluation time : {dt:.3f} s")
    eff = res["efficiency"]
    print("Efficiency (Red)     : {:8.3f} %".format(eff[2] * 100))
    print("Efficiency (Green)   : {:8.3f} %".format(eff[1] * 100))
    print("Efficiency (Blue)    : {:8.3f} %".format(eff[0] * 100))
    print("Color dispersion     : {:8.2f}".format(res["delta_e"]))
    print("FoV uniformity       : {:8.2f} %".format(res["U_fov"] * 100))
    print("Eyebox uniformity    : {:8.2f} %".format(res["U_EB"] * 100))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
