"""ctypes front-end of oracle/libwgrt_oracle.so (TEST INFRASTRUCTURE ONLY).

``trace(scene_args...)`` walks rays on the CPU exactly as the reference kernel does
(GPU_ray_tracing_functions.py:833-1246, restated in wgrt_oracle.c) and mutates ``rng_states`` and
``matrix_EB`` in place, like the kernel.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from gpu_ray_tracing_for_waveguide_based_ar_display_b200._capi import (
    COUNTER_NAMES, LEGACY_COLS, WGRT_NUM_COUNTERS, WgrtLegacyProblem, WgrtProblem)
from gpu_ray_tracing_for_waveguide_based_ar_display_b200.GPU_ray_tracing_functions import (
    pack_problem)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libwgrt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "wgrt_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libwgrt_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.wgrt_oracle_trace.restype = C.c_int
        L.wgrt_oracle_trace.argtypes = [C.POINTER(WgrtProblem), C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        L.wgrt_oracle_trace_events.restype = C.c_int64
        L.wgrt_oracle_trace_events.argtypes = [C.POINTER(WgrtProblem), C.c_int64, C.c_void_p, C.c_int64]
        L.wgrt_oracle_legacy_step.restype = C.c_int
        L.wgrt_oracle_legacy_step.argtypes = [C.POINTER(WgrtLegacyProblem), C.c_void_p]
        L.wgrt_oracle_legacy_pack.restype = C.c_int64
        L.wgrt_oracle_legacy_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.wgrt_oracle_locate.restype = C.c_int
        L.wgrt_oracle_locate.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_int64, C.c_void_p]
        L.wgrt_oracle_efield.restype = C.c_int
        L.wgrt_oracle_efield.argtypes = [C.c_void_p] * 4 + [C.c_int64, C.c_void_p]
        L.wgrt_oracle_xorshift.restype = C.c_int
        L.wgrt_oracle_xorshift.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def trace(*args, first: int = 0, count: Optional[int] = None, num_threads: int = 0,
          counters: bool = False, single_lambda: bool = False, threshold: float = 0.0, ray_index_base: int = 0):
    """Same 33 positional arguments as the reference kernel, all host NumPy arrays
    (32 with ``single_lambda=True``: the argument list of ``process_rays_kernel_pro``)."""
    if single_lambda:
        args = tuple(args[:8]) + (None,) + tuple(args[8:])
    prob, keep = pack_problem(args, host=True, single_lambda=single_lambda, threshold=threshold,
                              ray_index_base=ray_index_base)
    n = prob.num_rays if count is None else count
    cnt = np.zeros(WGRT_NUM_COUNTERS, dtype=np.uint64)
    nt = num_threads or (os.cpu_count() or 1)
    rc = lib().wgrt_oracle_trace(C.byref(prob), first, n, cnt.ctypes.data, nt)
    del keep
    if rc != 0:
        raise RuntimeError(f"oracle error {rc}")
    if counters:
        return {k: int(cnt[i]) for i, k in enumerate(COUNTER_NAMES)}
    return None


EVENT_FIELDS = ("state", "u", "e1", "e12", "e123", "x", "y", "ener")


def trace_events(*args, idx: int, cap: int = 4096, single_lambda: bool = False, threshold: float = 0.0,
                 ray_index_base: int = 0) -> np.ndarray:
    """Walk ray ``idx`` alone and return its decisions as an [n_events, 8] array (EVENT_FIELDS): every draw
    ``u`` with the cumulative efficiencies it was compared with (state -1 = in-coupling from air; e123 is
    NaN for two-order events).  Mutates that ray's RNG state and the bins like a one-ray launch."""
    if single_lambda:
        args = tuple(args[:8]) + (None,) + tuple(args[8:])
    prob, keep = pack_problem(args, host=True, single_lambda=single_lambda, threshold=threshold,
                              ray_index_base=ray_index_base)
    ev = np.zeros((cap, 8), dtype=np.float64)
    n = lib().wgrt_oracle_trace_events(C.byref(prob), int(idx), ev.ctypes.data, cap)
    del keep
    if n < 0:
        raise RuntimeError(f"oracle error {n}")
    return ev[:min(n, cap)]


def legacy_step(*args) -> int:
    """One launch of the legacy ``process_rays_kernel`` (GRTF:192-417) on host arrays: the 21 positional
    arguments of the reference kernel; ``vectors``, ``d_total_ray_counter`` and ``matrix_EB`` are mutated in
    place.  Rows are processed in index order.  Returns the number of children dropped for lack of rows."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200.legacy import pack_legacy_problem
    prob, keep = pack_legacy_problem(args, host=True)
    dropped = C.c_uint64(0)
    rc = lib().wgrt_oracle_legacy_step(C.byref(prob), C.byref(dropped))
    del keep
    if rc != 0:
        raise RuntimeError(f"oracle error {rc}")
    return int(dropped.value)


def legacy_pack(src: np.ndarray, dst: np.ndarray, src_len: int) -> int:
    """``pack_active_to_front`` (GRTF:178-190), rows visited in index order; returns the packed count."""
    assert src.dtype == np.float64 and dst.dtype == np.float64 and src.shape[1] == LEGACY_COLS
    return int(lib().wgrt_oracle_legacy_pack(src.ctypes.data, dst.ctypes.data, int(src_len)))


def legacy_trace(vectors, geom, luts, eb=(80, 120), max_steps=100000, max_generations=64, capacity=None):
    """The generation loop of ``legacy.trace`` with the oracle kernels: returns (matrix_EB, live rows, stats)."""
    n0 = vectors.shape[0]
    capacity = int(capacity) if capacity else max(4 * n0, 1024)
    a = np.zeros((capacity, LEGACY_COLS)); b = np.zeros_like(a)
    a[:n0] = vectors
    X, Y, _ = geom["lut_TIR"].shape
    EB = np.zeros((Y, X, eb[0], eb[1]), dtype=np.float32)
    count, st = n0, dict(generations=0, live_rows=0, rows_processed=0, children=0, children_dropped=0, max_live_rows=n0)
    while count > 0 and st["generations"] < max_generations:
        counter = np.array([count], dtype=np.int32)
        st["children_dropped"] += legacy_step(a, count, counter, int(max_steps), geom["IC"], geom["FC"], geom["FC_offset"],
                                              geom["OC"], geom["OC_offset"], geom["eff_reg1"], geom["eff_reg2"],
                                              geom["eff_reg_FOV"], geom["eff_reg_FOV_range"], luts["lut_ic1"], luts["lut_ic2"],
                                              luts["lut_fc1"], luts["lut_fc2"], luts["lut_oc"], geom["lut_TIR"],
                                              geom["lut_gap"], EB)
        total = min(int(counter[0]), capacity)
        st["rows_processed"] += count
        st["children"] += int(counter[0]) - count
        count = legacy_pack(a, b, total)
        a, b = b, a
        st["generations"] += 1
        st["max_live_rows"] = max(st["max_live_rows"], count)
    st["live_rows"] = count
    return EB, a[:count].copy(), st


def locate(verts, offsets, px, py):
    verts = np.ascontiguousarray(verts, dtype=np.float64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    px = np.ascontiguousarray(px, dtype=np.float64)
    py = np.ascontiguousarray(py, dtype=np.float64)
    out = np.empty(len(px), dtype=np.int32)
    lib().wgrt_oracle_locate(verts.ctypes.data, len(verts), offsets.ctypes.data, len(offsets) - 1,
                             px.ctypes.data, py.ctypes.data, len(px), out.ctypes.data)
    return out


def efield(ete, etm, delta, jones):
    ete = np.ascontiguousarray(ete, dtype=np.float64)
    etm = np.ascontiguousarray(etm, dtype=np.float64)
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    jones = np.ascontiguousarray(jones, dtype=np.complex128)
    out = np.empty((len(ete), 3), dtype=np.float64)
    lib().wgrt_oracle_efield(ete.ctypes.data, etm.ctypes.data, delta.ctypes.data, jones.ctypes.data,
                             len(ete), out.ctypes.data)
    return out


def xorshift(states, draws):
    states = np.ascontiguousarray(states, dtype=np.uint32).copy()
    last = np.empty(len(states), dtype=np.float64)
    lib().wgrt_oracle_xorshift(states.ctypes.data, len(states), draws, last.ctypes.data)
    return states, last
