"""The legacy deterministic energy-splitting tracer.

Drop-in surface for ``GRTF.process_rays_kernel`` (/root/reference/GPU_ray_tracing_functions.py:192-417) and its
support kernels ``pack_active_to_front`` / ``zero_out_kernel`` / ``reset_counter_kernel`` (GRTF:167-190), plus the
generation loop the reference never shipped (``trace``).  Unlike the Monte-Carlo kernel of the runner, a ray that
hits a fold-coupler slice splits into the zero order (kept in its own row) and the diffracted order (appended as a
new row at an atomically incremented index); in the out-coupler zone every hit deposits the out-coupled energy
``|E|^2`` into the eyebox bin and the ray carries on with the zero order.  One launch advances every live row to
its next split (or its end); the launches of a job form "per-bounce ray queues", compacted in between.

Row layout (``vectors[N, 13]`` float64): x, y, gap_x, gap_y, theta, phi, m, n, Ete, Etm, delta_phase,
region_state, flag.  Tables are single-wavelength: ``lut_ic1/2 [X, Y, C]``, ``lut_fc1/2 [nFC, X, Y, C]``,
``lut_oc [nOC, X, Y, C]`` complex128 with C = 26 (3 orders per group), ``lut_TIR [X, Y, 4]``, ``lut_gap [X, Y, 8]``,
``matrix_EB [Y, X, EBy, EBx]`` float32.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import LEGACY_COLS, WgrtLegacyProblem
from .GPU_ray_tracing_functions import _Buf, _describe, _stream_handle, _want

__all__ = ["process_rays_kernel", "pack_active_to_front", "zero_out_kernel", "reset_counter_kernel", "trace",
           "pack_legacy_problem", "make_legacy_luts", "initial_rows"]

_ARGS = ("vectors", "useful_count_in", "d_total_ray_counter", "MAX_STEPS", "IC", "FC", "FC_offset", "OC", "OC_offset",
         "eff_reg1", "eff_reg2", "eff_reg_FOV", "eff_reg_FOV_range", "lut_ic1", "lut_ic2", "lut_fc1", "lut_fc2", "lut_oc",
         "lut_TIR", "lut_gap", "matrix_EB")
_SCALARS = {"useful_count_in", "MAX_STEPS"}


def pack_legacy_problem(args: Sequence[Any], host: bool) -> Tuple[WgrtLegacyProblem, list]:
    """Validate the 21 positional arguments of ``process_rays_kernel`` (GRTF:193-201) and fill a
    ``wgrt_legacy_problem_t``.  ``host=True`` wants NumPy arrays, ``host=False`` device buffers."""
    if len(args) != len(_ARGS):
        raise TypeError(f"process_rays_kernel takes {len(_ARGS)} positional arguments ({len(args)} given)")
    B: Dict[str, Optional[_Buf]] = {}
    for a, nm in zip(args, _ARGS):
        if nm in _SCALARS:
            continue
        b = _describe(a, nm)
        if b.is_host != host:
            raise TypeError(f"{nm}: expected a {'host' if host else 'device'} buffer")
        B[nm] = b
    _want(B["vectors"], "vectors", np.float64, 2)
    if B["vectors"].shape[1] != LEGACY_COLS:
        raise ValueError(f"vectors: expected shape [N, {LEGACY_COLS}]")
    _want(B["d_total_ray_counter"], "d_total_ray_counter", np.int32, 1)
    for nm in ("IC", "FC", "OC", "eff_reg1", "eff_reg2"):
        _want(B[nm], nm, np.float64, 2)
        if B[nm].shape[1] != 2:
            raise ValueError(f"{nm}: expected shape [V, 2]")
    for nm in ("FC_offset", "OC_offset"):
        _want(B[nm], nm, np.int64, 1)
        if B[nm].shape[0] < 1:
            raise ValueError(f"{nm}: needs at least one entry")
        if host:
            off = B[nm].host_array
            if off[0] != 0 or np.any(np.diff(off) < 0) or off[-1] > B[nm[:2]].shape[0]:
                raise ValueError(f"{nm}: must start at 0, be non-decreasing and end within the vertex array")
    n_FC, n_OC = B["FC_offset"].shape[0] - 1, B["OC_offset"].shape[0] - 1
    _want(B["lut_TIR"], "lut_TIR", np.float64, 3)
    _want(B["lut_gap"], "lut_gap", np.float64, 3)
    X, Y, k = B["lut_TIR"].shape
    if k != 4 or B["lut_gap"].shape != (X, Y, 8):
        raise ValueError("lut_TIR must be [X,Y,4] and lut_gap [X,Y,8]")
    _want(B["eff_reg_FOV"], "eff_reg_FOV", np.float64, 4)
    _want(B["eff_reg_FOV_range"], "eff_reg_FOV_range", np.float64, 3)
    if B["eff_reg_FOV"].shape != (X, Y, 4, 2) or B["eff_reg_FOV_range"].shape != (X, Y, 4):
        raise ValueError("eff_reg_FOV must be [X,Y,4,2] and eff_reg_FOV_range [X,Y,4] with the X,Y of lut_TIR")
    ch = {}
    for nm, lead, cmin in (("lut_ic1", (), 24), ("lut_ic2", (), 24), ("lut_fc1", (n_FC,), 20), ("lut_fc2", (n_FC,), 20),
                           ("lut_oc", (n_OC,), 26)):
        _want(B[nm], nm, np.complex128, 3 + len(lead))
        shp = B[nm].shape
        if shp[:-1] != lead + (X, Y) or shp[-1] < cmin:
            raise ValueError(f"{nm}: shape {shp} where {lead + (X, Y)} + (>= {cmin},) is required")
        ch[nm] = shp[-1]
    if ch["lut_ic1"] != ch["lut_ic2"] or ch["lut_fc1"] != ch["lut_fc2"]:
        raise ValueError("LUTs of one coupler family must have the same channel count")
    _want(B["matrix_EB"], "matrix_EB", np.float32, 4)
    if B["matrix_EB"].shape[:2] != (Y, X):
        raise ValueError(f"matrix_EB: leading shape {B['matrix_EB'].shape[:2]} where (Y, X) = {(Y, X)} is required")
    count = int(args[1])
    if count < 0 or count > B["vectors"].shape[0]:
        raise ValueError("useful_count_in must be within [0, len(vectors)]")
    p = WgrtLegacyProblem()
    p.vectors, p.capacity, p.useful_count_in = B["vectors"].ptr, B["vectors"].shape[0], count
    p.total_ray_counter, p.max_steps = B["d_total_ray_counter"].ptr, int(args[3])
    p.IC, p.IC_n = B["IC"].ptr, B["IC"].shape[0]
    p.FC, p.FC_n, p.FC_offset, p.n_FC = B["FC"].ptr, B["FC"].shape[0], B["FC_offset"].ptr, n_FC
    p.OC, p.OC_n, p.OC_offset, p.n_OC = B["OC"].ptr, B["OC"].shape[0], B["OC_offset"].ptr, n_OC
    p.eff_reg1, p.eff_reg1_n = B["eff_reg1"].ptr, B["eff_reg1"].shape[0]
    p.eff_reg2, p.eff_reg2_n = B["eff_reg2"].ptr, B["eff_reg2"].shape[0]
    p.eff_reg_FOV, p.eff_reg_FOV_range = B["eff_reg_FOV"].ptr, B["eff_reg_FOV_range"].ptr
    for nm in ("lut_ic1", "lut_ic2", "lut_fc1", "lut_fc2", "lut_oc"):
        setattr(p, nm, B[nm].ptr)
    p.C_ic, p.C_fc, p.C_oc = ch["lut_ic1"], ch["lut_fc1"], ch["lut_oc"]
    p.lut_TIR, p.lut_gap, p.X, p.Y = B["lut_TIR"].ptr, B["lut_gap"].ptr, X, Y
    p.matrix_EB, p.EBy, p.EBx = B["matrix_EB"].ptr, B["matrix_EB"].shape[2], B["matrix_EB"].shape[3]
    return p, [b.owner for b in B.values()]


class _Kernel:
    """``obj[grid, block(, stream)](*args)`` like a Numba dispatcher; the launch shape is the engine's own."""

    def __init__(self, fn):
        self._fn = fn

    def __getitem__(self, config):
        if not isinstance(config, tuple):
            config = (config,)
        stream = config[2] if len(config) > 2 else None
        return lambda *args: self._fn(args, stream)


def _stage(args, names):
    """Host NumPy arguments are copied to the device for the launch (as Numba does); returns the device
    argument list and the (host, tensor) pairs to copy back."""
    import torch
    from .GPU_ray_tracing_functions import _TorchAlias
    dev, back = list(args), []
    for i, (a, nm) in enumerate(zip(args, names)):
        if isinstance(a, np.ndarray):
            if not torch.cuda.is_available():
                raise _capi.WgrtError("no CUDA device: the engine has no CPU fallback")
            if not a.flags.c_contiguous:
                raise ValueError(f"{nm}: host array must be C-contiguous")
            view = a.view(np.float64) if a.dtype == np.complex128 else a
            t = torch.from_numpy(view).cuda()
            dev[i] = _TorchAlias(t, a.shape, a.dtype)
            back.append((a, t))
    return dev, back


def _launch_step(args, stream):
    lib = _capi.load_library()
    if len(args) != len(_ARGS):
        raise TypeError(f"process_rays_kernel takes {len(_ARGS)} positional arguments ({len(args)} given)")
    dev, back = _stage(args, _ARGS)
    prob, keep = pack_legacy_problem(dev, host=False)
    if back:
        import torch
        torch.cuda.current_stream().synchronize()
    _capi.check(lib.wgrt_legacy_step(C.byref(prob), C.c_void_p(_stream_handle(stream))), lib)
    if back:
        import torch
        torch.cuda.synchronize()
        for host, t in back:      # Numba copies every host argument back; only these three can have changed
            if host is args[0] or host is args[2] or host is args[20]:
                host[...] = t.cpu().numpy().view(host.dtype).reshape(host.shape)
    del keep


def _launch_pack(args, stream):
    lib = _capi.load_library()
    if len(args) != 4:
        raise TypeError(f"pack_active_to_front takes 4 positional arguments ({len(args)} given)")
    dev, back = _stage(args, ("src", "dst", "src_len", "out_count"))
    src, dst, cnt = _describe(dev[0], "src"), _describe(dev[1], "dst"), _describe(dev[3], "out_count")
    for b, nm in ((src, "src"), (dst, "dst")):
        _want(b, nm, np.float64, 2)
        if b.shape[1] != LEGACY_COLS:
            raise ValueError(f"{nm}: expected shape [N, {LEGACY_COLS}]")
    _want(cnt, "out_count", np.int32, 1)
    n = int(args[2])
    if n < 0 or n > src.shape[0] or n > dst.shape[0]:
        raise ValueError("src_len exceeds src / dst")
    if back:
        import torch
        torch.cuda.current_stream().synchronize()
    _capi.check(lib.wgrt_legacy_pack_active(C.c_void_p(src.ptr), C.c_void_p(dst.ptr), n, C.c_void_p(cnt.ptr),
                                            C.c_void_p(_stream_handle(stream))), lib)
    if back:
        import torch
        torch.cuda.synchronize()
        for host, t in back:
            if host is args[1] or host is args[3]:
                host[...] = t.cpu().numpy().view(host.dtype).reshape(host.shape)


def _launch_zero(args, stream):
    """zero_out_kernel(array) (GRTF:167-171): a memset on the caller's stream."""
    import torch
    (a,) = args
    if isinstance(a, np.ndarray):
        a[...] = 0
        return
    b = _describe(a, "array")
    n = int(np.prod(b.shape)) * b.dtype.itemsize
    from ctypes import CDLL
    rt = CDLL("libcudart.so.12")
    rt.cudaMemsetAsync.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
    if rt.cudaMemsetAsync(C.c_void_p(b.ptr), 0, n, C.c_void_p(_stream_handle(stream))) != 0:
        raise _capi.WgrtError("cudaMemsetAsync failed")


process_rays_kernel = _Kernel(_launch_step)
pack_active_to_front = _Kernel(_launch_pack)
zero_out_kernel = _Kernel(_launch_zero)
reset_counter_kernel = _Kernel(_launch_zero)          # GRTF:173-176: counter[0] = 0 on a one-element array


def trace(vectors: np.ndarray, geom: Dict[str, np.ndarray], luts: Dict[str, np.ndarray], eb: Tuple[int, int] = (80, 120),
          max_steps: int = 100000, max_generations: int = 64, capacity: Optional[int] = None,
          matrix_EB: Optional[np.ndarray] = None):
    """The whole deterministic job on the GPU (``wgrt_legacy_trace_host``): generation after generation, one
    ``process_rays_kernel`` launch over the live rows and a ballot / prefix-sum compaction of the survivors and
    their children into the other queue buffer, until no row is live.

    ``vectors`` [N0, 13] are the initial rows (see ``initial_rows``); ``geom`` holds IC, FC, FC_offset, OC,
    OC_offset, eff_reg1, eff_reg2, eff_reg_FOV, eff_reg_FOV_range and single-wavelength lut_TIR [X,Y,4] /
    lut_gap [X,Y,8]; ``luts`` holds lut_ic1, lut_ic2, lut_fc1, lut_fc2, lut_oc.  Returns
    ``(matrix_EB [Y, X, EBy, EBx], live_rows, stats)``.
    """
    lib = _capi.load_library()
    n0 = vectors.shape[0]
    capacity = int(capacity) if capacity else max(4 * n0, 1024)
    rows = np.zeros((capacity, LEGACY_COLS), dtype=np.float64)
    rows[:n0] = vectors
    X, Y, _ = geom["lut_TIR"].shape
    if matrix_EB is None:
        matrix_EB = np.zeros((Y, X, eb[0], eb[1]), dtype=np.float32)
    counter = np.array([n0], dtype=np.int32)
    args = (rows, n0, counter, int(max_steps), geom["IC"], geom["FC"], geom["FC_offset"], geom["OC"], geom["OC_offset"],
            geom["eff_reg1"], geom["eff_reg2"], geom["eff_reg_FOV"], geom["eff_reg_FOV_range"], luts["lut_ic1"],
            luts["lut_ic2"], luts["lut_fc1"], luts["lut_fc2"], luts["lut_oc"], geom["lut_TIR"], geom["lut_gap"], matrix_EB)
    prob, keep = pack_legacy_problem(args, host=True)
    stats = np.zeros(8, dtype=np.uint64)
    _capi.check(lib.wgrt_legacy_trace_host(C.byref(prob), int(max_generations), stats.ctypes.data), lib)
    del keep
    names = ("generations", "live_rows", "rows_processed", "children", "children_dropped", "max_live_rows")
    st = {k: int(stats[i]) for i, k in enumerate(names)}
    return matrix_EB, rows[:st["live_rows"]].copy(), st


# ------------------------------------------------------------------------------------------------------------
# synthetic inputs in the legacy layout (the reference ships no LUT files for this kernel)
# ------------------------------------------------------------------------------------------------------------
def make_legacy_luts(scene, lam: int = 1, seed: int = 0) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    """Single-wavelength tables of the legacy shapes (3 orders per group, C = 26) from a full-colour scene
    (``synthetic_inputs.make_scene``): geometry unchanged, ``lut_TIR`` / ``lut_gap`` sliced at wavelength ``lam``,
    Jones entries drawn so that the zero orders carry most of the energy and the split / out-coupled orders a
    few percent (no energy is created: every 2x2 block has spectral norm < 1)."""
    rs = np.random.default_rng(seed)
    g = {k: v for k, v in scene.geom.items() if k not in ("lut_TIR", "lut_gap")}
    g["lut_TIR"] = np.ascontiguousarray(scene.geom["lut_TIR"][lam])
    g["lut_gap"] = np.ascontiguousarray(scene.geom["lut_gap"][lam])
    X, Y, _ = g["lut_TIR"].shape
    nFC, nOC = len(g["FC_offset"]) - 1, len(g["OC_offset"]) - 1

    def table(lead, quartets):
        shape = lead + (X, Y, 26)
        t = (rs.uniform(0.005, 0.02, shape) * np.exp(1j * rs.uniform(-np.pi, np.pi, shape))).astype(np.complex128)
        t[..., 0] = rs.uniform(0.6, 1.0, lead + (X, Y))     # theta, phi of the outgoing direction (pass-through values)
        t[..., 1] = rs.uniform(-3.0, 3.0, lead + (X, Y))
        for (a, b, c, d), amp in quartets:
            ph = rs.uniform(-np.pi, np.pi, (4,) + lead + (X, Y))
            t[..., a] = amp * rs.uniform(0.95, 1.0, lead + (X, Y)) * np.exp(1j * ph[0])
            t[..., d] = 0.93 * amp * rs.uniform(0.95, 1.0, lead + (X, Y)) * np.exp(1j * ph[1])
            t[..., b] = 0.05 * amp * np.exp(1j * ph[2])
            t[..., c] = 0.05 * amp * np.exp(1j * ph[3])
        return t

    luts = {"lut_ic1": table((), [((8, 11, 20, 23), 0.6)]),
            "lut_ic2": table((), [((3, 6, 15, 18), 0.9)]),
            "lut_fc1": table((nFC,), [((3, 6, 15, 18), 0.92), ((4, 7, 16, 19), 0.3)]),
            "lut_fc2": table((nFC,), [((3, 6, 15, 18), 0.92), ((2, 5, 14, 17), 0.3)]),
            "lut_oc": table((nOC,), [((3, 6, 15, 18), 0.9), ((10, 13, 22, 25), 0.3)])}
    return g, luts


def initial_rows(points: np.ndarray, X: int, Y: int) -> np.ndarray:
    """One TE and one TM ray per start point and FoV cell, in region state 0 (about to be in-coupled)."""
    P = len(points)
    mm, nn, pol, pp = np.meshgrid(np.arange(X), np.arange(Y), np.arange(2), np.arange(P), indexing="ij")
    n = mm.size
    rows = np.zeros((n, LEGACY_COLS), dtype=np.float64)
    rows[:, 0] = points[pp.ravel(), 0].astype(np.float32)
    rows[:, 1] = points[pp.ravel(), 1].astype(np.float32)
    rows[:, 6], rows[:, 7] = mm.ravel(), nn.ravel()
    rows[:, 8] = 1.0 - pol.ravel()
    rows[:, 9] = pol.ravel()
    rows[:, 12] = 1.0
    return rows
