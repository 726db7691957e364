"""Drop-in surface of the reference module ``GPU_ray_tracing_functions`` for its hot path.

The reference runner does (gpu_ray_tracing_pro_fullColor.py:7, 79, 170-177)::

    import GPU_ray_tracing_functions as GRTF
    points = GRTF.generate_points_in_polygon(IC, n)
    GRTF.process_rays_kernel_pro_fullColor[blocks_per_grid, threads_per_block](...33 args...)

This module offers the same two names with the same positional arguments.  The kernel object is
not a Numba dispatcher: ``obj[grid, block]`` returns a callable that packs the 33 arguments into a
``wgrt_problem_t`` (include/wgrt.h) and calls ``wgrt_trace_fullcolor`` in libwgrt.so, the
hand-written sm_100a engine.  ``grid`` / ``block`` are accepted and ignored (the engine chooses
its own launch shape); an optional third element is a stream, as with Numba.

Buffers: anything exposing ``__cuda_array_interface__`` (Numba device arrays as the runner makes
them with ``cuda.to_device``, torch CUDA tensors, CuPy) is used in place, zero copy.  Host NumPy
arrays are staged to the device and ``rng_states`` / ``matrix_EB`` are copied back after the
launch, which is what Numba does for host arguments.  Like the reference kernel, the launch
mutates only ``rng_states`` and ``matrix_EB`` (GPU_ray_tracing_functions.py:33, 164).

There is no CPU fallback: without libwgrt.so (or without a CUDA device) the launch raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import WgrtProblem

__all__ = ["generate_points_in_polygon", "process_rays_kernel_pro_fullColor", "process_rays_kernel_pro", "pack_problem",
           "RayWalkKernel"]


# ------------------------------------------------------------------------------------------
# host helper (GPU_ray_tracing_functions.py:12-23)
# ------------------------------------------------------------------------------------------
def _contains_points(poly: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """Vectorised even-odd containment of ``pts`` [K,2] in the ring ``poly`` [V,2]."""
    xi, yi = poly[:, 0], poly[:, 1]
    xj, yj = np.roll(xi, 1), np.roll(yi, 1)
    px, py = pts[:, 0:1], pts[:, 1:2]
    straddle = (yi[None, :] > py) != (yj[None, :] > py)
    with np.errstate(divide="ignore", invalid="ignore"):
        xint = (xj - xi)[None, :] * (py - yi[None, :]) / (yj - yi)[None, :] + xi[None, :]
    return (np.count_nonzero(straddle & (px < xint), axis=1) & 1).astype(bool)


def generate_points_in_polygon(polygon_vertices, num_points, rng: Optional[np.random.Generator] = None):
    """Uniform points inside a polygon by rejection sampling from its bounding box.

    Same signature and sampling scheme as the reference (batches of twice the number still
    missing, first ``num_points`` kept).  Without ``rng`` it draws from NumPy's global state,
    like the reference, so ``np.random.seed`` makes it reproducible; pass a ``Generator`` for an
    explicitly seeded stream.
    """
    polygon_vertices = np.asarray(polygon_vertices, dtype=np.float64)
    num_points = int(num_points)
    xmin, ymin = np.min(polygon_vertices, axis=0)
    xmax, ymax = np.max(polygon_vertices, axis=0)
    kept: List[np.ndarray] = []
    have = 0
    while have < num_points:
        k = (num_points - have) * 2
        if rng is None:
            cand = np.random.uniform(low=[xmin, ymin], high=[xmax, ymax], size=(k, 2))
        else:
            cand = rng.uniform(low=[xmin, ymin], high=[xmax, ymax], size=(k, 2))
        good = cand[_contains_points(polygon_vertices, cand)]
        kept.append(good)
        have += len(good)
    if not kept:
        return np.zeros((0, 2))
    return np.concatenate(kept, axis=0)[:num_points]


# ------------------------------------------------------------------------------------------
# argument packing
# ------------------------------------------------------------------------------------------
_ARG_NAMES = ("x_v", "y_v", "gap_x_v", "gap_y_v", "pol_v", "azi_v", "m_v", "n_v", "lmd_num", "te_v",
              "tm_v", "delta_phase_v", "rng_states", "IC", "FC", "FC_offset", "OC", "OC_offset", "n_g",
              "eff_reg1", "eff_reg2", "eff_reg_FOV", "eff_reg_FOV_range", "lut_ic1", "lut_ic2",
              "lut_ic3", "lut_fc1", "lut_fc2", "lut_oc1", "lut_oc2", "lut_TIR", "lut_gap", "matrix_EB")
_DEAD_ARGS = {2, 3, 4, 5}          # gap_x_v, gap_y_v, pol_v, azi_v: never read by the walk


class _Buf:
    """Pointer + shape + dtype of one array argument."""
    __slots__ = ("ptr", "shape", "dtype", "owner", "is_host", "host_array", "staged")

    def __init__(self, ptr, shape, dtype, owner, is_host, host_array=None):
        self.ptr, self.shape, self.dtype, self.owner = ptr, tuple(shape), np.dtype(dtype), owner
        self.is_host, self.host_array, self.staged = is_host, host_array, None


def _describe(obj: Any, name: str) -> _Buf:
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        strides = cai.get("strides")
        shape = tuple(cai["shape"])
        dtype = np.dtype(cai["typestr"])
        if strides is not None and all(s > 0 for s in shape):
            # dense C order; the stride of an axis of extent 1 is never used and may be anything
            acc = dtype.itemsize
            for extent, stride in zip(reversed(shape), reversed(tuple(strides))):
                if extent > 1 and stride != acc:
                    raise ValueError(f"{name}: device array must be C-contiguous")
                acc *= extent
        return _Buf(int(cai["data"][0] or 0), shape, dtype, obj, False)
    if isinstance(obj, np.ndarray):
        if not obj.flags.c_contiguous:
            raise ValueError(f"{name}: host array must be C-contiguous")
        return _Buf(obj.ctypes.data, obj.shape, obj.dtype, obj, True, obj)
    raise TypeError(f"{name}: expected a NumPy array or an object with __cuda_array_interface__, "
                    f"got {type(obj).__name__}")


def _want(b: _Buf, name: str, dtype, ndim: int):
    if b.dtype != np.dtype(dtype):
        raise TypeError(f"{name}: dtype {b.dtype} where {np.dtype(dtype)} is required")
    if len(b.shape) != ndim:
        raise ValueError(f"{name}: {len(b.shape)}-D array where {ndim}-D is required")


def pack_problem(args: Sequence[Any], host: bool, flags: int = 0, tile_hint: int = 0,
                 runner_points: int = 0, runner_first_cell: int = 0, num_rays: Optional[int] = None,
                 single_lambda: bool = False, threshold: float = 0.0,
                 ray_index_base: int = 0, eb: Optional[Tuple[int, int]] = None,
                 rng_seed_offset: int = 0) -> Tuple[WgrtProblem, list]:
    """Validate the 33 positional kernel arguments and fill a ``wgrt_problem_t``.

    ``host=True`` requires NumPy arrays (used for the host entry point and by the test oracle);
    ``host=False`` requires device buffers (host arrays must have been staged by the caller).
    Raises ``TypeError`` / ``ValueError`` on dtype / rank / shape mismatches -- the analogue of
    Numba's typing error at the first launch.

    Runner layout (``runner_points = P > 0``, see include/wgrt.h): ``x_v`` / ``y_v`` hold the P
    start points, the arrays m_v .. delta_phase_v may be ``None``, ``num_rays`` gives the launch
    size (a multiple of 2P) and ``rng_states`` may be ``None`` for the host entry point.

    ``single_lambda=True`` is the layout of ``process_rays_kernel_pro`` (GRTF:419-831): ``lmd_num`` is
    ``None`` and the tables / bins lack the wavelength axis (LUTs [X,Y,C] or [n,X,Y,C], lut_TIR
    [X,Y,4], lut_gap [X,Y,8], matrix_EB [Y,X,EBy,EBx]); they are the L = 1 case of the same memory
    layout.  ``threshold`` is the energy gate of the fold / out-coupler branches.
    """
    if len(args) != len(_ARG_NAMES):
        raise TypeError(f"process_rays_kernel_pro_fullColor takes {len(_ARG_NAMES)} positional "
                        f"arguments ({len(args)} given)")
    bufs: List[Optional[_Buf]] = []
    for i, (a, nm) in enumerate(zip(args, _ARG_NAMES)):
        if nm == "n_g":
            bufs.append(None)
            continue
        if a is None and (i in _DEAD_ARGS or (single_lambda and i == 8) or
                          (runner_points > 0 and (6 <= i <= 11 or (i == 12 and host))) or
                          (nm == "matrix_EB" and host and eb is not None)):   # wgrt_trace_evaluate_host: bins stay on the device
            bufs.append(None)
            continue
        b = _describe(a, nm)
        if nm == "matrix_EB" and host and not b.is_host and (flags & _capi.WGRT_FLAG_BINS_DEVICE):
            bufs.append(b)          # host entry accumulating into the caller's device tensor
            continue
        if b.is_host != host:
            raise TypeError(f"{nm}: expected a {'host' if host else 'device'} buffer")
        bufs.append(b)
    B = dict(zip(_ARG_NAMES, bufs))

    N = None
    if runner_points > 0:
        for nm in ("x_v", "y_v"):
            _want(B[nm], nm, np.float32, 1)
            if B[nm].shape[0] != runner_points:
                raise ValueError(f"{nm}: runner layout needs {runner_points} start points")
        if num_rays is None or num_rays < 0 or num_rays % (2 * runner_points):
            raise ValueError("runner layout: num_rays must be a multiple of 2 * runner_points")
        N = int(num_rays)
        for i in range(6, 12):
            B[_ARG_NAMES[i]] = None
    else:
        for i in range(12):
            nm = _ARG_NAMES[i]
            b = B[nm]
            if b is None:
                continue
            _want(b, nm, np.float32, 1)
            if N is None:
                N = b.shape[0]
            elif b.shape[0] != N:
                raise ValueError(f"{nm}: length {b.shape[0]} differs from x_v length {N}")
    if B["rng_states"] is not None:
        _want(B["rng_states"], "rng_states", np.uint32, 1)
        if B["rng_states"].shape[0] != N:
            raise ValueError("rng_states: length differs from the ray arrays")
    for nm in ("IC", "FC", "OC", "eff_reg1", "eff_reg2"):
        _want(B[nm], nm, np.float64, 2)
        if B[nm].shape[1] != 2:
            raise ValueError(f"{nm}: expected shape [V, 2]")
    for nm in ("FC_offset", "OC_offset"):
        _want(B[nm], nm, np.int64, 1)
        if B[nm].shape[0] < 1:
            raise ValueError(f"{nm}: needs at least one entry")
    n_FC = B["FC_offset"].shape[0] - 1
    n_OC = B["OC_offset"].shape[0] - 1
    if host:
        for nm, tot in (("FC_offset", B["FC"].shape[0]), ("OC_offset", B["OC"].shape[0])):
            off = B[nm].host_array
            if off[0] != 0 or np.any(np.diff(off) < 0) or off[-1] > tot:
                raise ValueError(f"{nm}: must start at 0, be non-decreasing and end within the vertex array")

    if single_lambda:
        # present the L-less arrays as L = 1 (same bytes)
        for nm, nd in (("lut_TIR", 3), ("lut_gap", 3), ("lut_ic1", 3), ("lut_ic2", 3), ("lut_ic3", 3),
                       ("lut_fc1", 4), ("lut_fc2", 4), ("lut_oc1", 4), ("lut_oc2", 4), ("matrix_EB", 4)):
            b = B[nm]
            if len(b.shape) != nd:
                raise ValueError(f"{nm}: {len(b.shape)}-D array where {nd}-D is required (single-wavelength layout)")
            lead = 1 if nm in ("lut_fc1", "lut_fc2", "lut_oc1", "lut_oc2") else 0
            b.shape = b.shape[:lead] + (1,) + b.shape[lead:]
        if B["lmd_num"] is not None:
            raise TypeError("lmd_num is not an argument of the single-wavelength kernel")
    _want(B["lut_TIR"], "lut_TIR", np.float64, 4)
    _want(B["lut_gap"], "lut_gap", np.float64, 4)
    L, X, Y, k = B["lut_TIR"].shape
    if k != 4 or B["lut_gap"].shape != (L, X, Y, 8):
        raise ValueError("lut_TIR must be [L,X,Y,4] and lut_gap [L,X,Y,8]")
    _want(B["eff_reg_FOV"], "eff_reg_FOV", np.float64, 4)
    _want(B["eff_reg_FOV_range"], "eff_reg_FOV_range", np.float64, 3)
    if B["eff_reg_FOV"].shape != (X, Y, 4, 2) or B["eff_reg_FOV_range"].shape != (X, Y, 4):
        raise ValueError("eff_reg_FOV must be [X,Y,4,2] and eff_reg_FOV_range [X,Y,4] with the "
                         "X,Y of lut_TIR")
    chans = {}
    for nm, lead, cmin in (("lut_ic1", (), 41), ("lut_ic2", (), 41), ("lut_ic3", (), 41),
                           ("lut_fc1", (n_FC,), 20), ("lut_fc2", (n_FC,), 20),
                           ("lut_oc1", (n_OC,), 41), ("lut_oc2", (n_OC,), 41)):
        _want(B[nm], nm, np.complex128, 4 + len(lead))
        shp = B[nm].shape
        if shp[:-1] != lead + (L, X, Y) or shp[-1] < cmin:
            raise ValueError(f"{nm}: shape {shp} where {lead + (L, X, Y)} + (>= {cmin},) is required")
        chans[nm] = shp[-1]
    if not (chans["lut_ic1"] == chans["lut_ic2"] == chans["lut_ic3"]) or \
            chans["lut_fc1"] != chans["lut_fc2"] or chans["lut_oc1"] != chans["lut_oc2"]:
        raise ValueError("LUTs of one coupler family must have the same channel count")
    if B["matrix_EB"] is None:
        eb = (L, Y, X, int(eb[0]), int(eb[1]))
        if eb[3] <= 0 or eb[4] <= 0:
            raise ValueError("eb: eyebox bin counts must be positive")
    else:
        _want(B["matrix_EB"], "matrix_EB", np.float32, 5)
        eb = B["matrix_EB"].shape
    if eb[:3] != (L, Y, X):
        raise ValueError(f"matrix_EB: leading shape {eb[:3]} where (L, Y, X) = {(L, Y, X)} is required")
    if n_FC > 250 or n_OC > 250:
        raise ValueError("at most 250 fold-coupler / out-coupler polygons are supported")
    if runner_points > 0 and runner_first_cell * 2 * runner_points + N > L * X * Y * 2 * runner_points:
        raise ValueError("runner layout: cell range exceeds L * X * Y")

    p = WgrtProblem()
    for i in range(12):
        nm = _ARG_NAMES[i]
        field = ("x", "y", "gap_x", "gap_y", "pol", "azi", "m", "n", "lmd_num", "te", "tm",
                 "delta_phase")[i]
        setattr(p, field, B[nm].ptr if B[nm] is not None else None)
    p.rng_states = B["rng_states"].ptr if B["rng_states"] is not None else None
    p.num_rays = N
    p.runner_points = runner_points
    p.runner_first_cell = runner_first_cell
    p.IC, p.IC_n = B["IC"].ptr, B["IC"].shape[0]
    p.FC, p.FC_n, p.FC_offset, p.n_FC = B["FC"].ptr, B["FC"].shape[0], B["FC_offset"].ptr, n_FC
    p.OC, p.OC_n, p.OC_offset, p.n_OC = B["OC"].ptr, B["OC"].shape[0], B["OC_offset"].ptr, n_OC
    p.n_g = float(args[18])
    p.eff_reg1, p.eff_reg1_n = B["eff_reg1"].ptr, B["eff_reg1"].shape[0]
    p.eff_reg2, p.eff_reg2_n = B["eff_reg2"].ptr, B["eff_reg2"].shape[0]
    p.eff_reg_FOV, p.eff_reg_FOV_range = B["eff_reg_FOV"].ptr, B["eff_reg_FOV_range"].ptr
    for nm in ("lut_ic1", "lut_ic2", "lut_ic3", "lut_fc1", "lut_fc2", "lut_oc1", "lut_oc2"):
        setattr(p, nm, B[nm].ptr)
    p.C_ic, p.C_fc, p.C_oc = chans["lut_ic1"], chans["lut_fc1"], chans["lut_oc1"]
    p.lut_TIR, p.lut_gap = B["lut_TIR"].ptr, B["lut_gap"].ptr
    p.L, p.X, p.Y = L, X, Y
    p.matrix_EB, p.EBy, p.EBx = (B["matrix_EB"].ptr if B["matrix_EB"] is not None else None), eb[3], eb[4]
    p.flags = flags
    p.tile_hint = tile_hint
    p.threshold = float(threshold)
    p.ray_index_base = int(ray_index_base)
    p.rng_seed_offset = int(rng_seed_offset)
    return p, [b.owner for b in bufs if b is not None]


_LUT_ARGS = range(23, 30)          # lut_ic1 .. lut_oc2
_warned_c64 = False


def _as_complex128(a: Any, name: str) -> Any:
    """The engine reads the RCWA tables as complex128.  A complex64 table (the dtype of the reference's
    ``.npy`` files is unknown: they are not reachable) is converted -- host arrays with NumPy, device
    buffers on the device -- instead of being rejected where Numba would have compiled a kernel for it.
    Note that the reference itself would then evaluate parts of ``E_field_cal`` in single precision;
    the engine computes in double on the widened values."""
    global _warned_c64
    cai = getattr(a, "__cuda_array_interface__", None)
    dt = np.dtype(cai["typestr"]) if cai is not None else getattr(a, "dtype", None)
    if dt != np.dtype(np.complex64):
        return a
    if not _warned_c64:
        import warnings
        warnings.warn(f"{name}: complex64 look-up table converted to complex128 for the launch", RuntimeWarning,
                      stacklevel=3)
        _warned_c64 = True
    if cai is None:
        return np.ascontiguousarray(a, dtype=np.complex128)
    import torch
    t = torch.as_tensor(a, device="cuda").to(torch.complex128).contiguous()
    return _TorchAlias(torch.view_as_real(t), tuple(cai["shape"]), np.complex128)


def _stream_handle(stream: Any) -> int:
    if stream is None or stream == 0:
        return 0
    if isinstance(stream, int):
        return stream
    for attr in ("cuda_stream", "ptr"):               # torch.cuda.Stream, cupy
        if hasattr(stream, attr):
            return int(getattr(stream, attr))
    h = getattr(stream, "handle", None)                # numba.cuda stream
    if h is not None:
        return int(getattr(h, "value", h) or 0)
    raise TypeError(f"cannot interpret {type(stream).__name__} as a CUDA stream")


class _Launcher:
    def __init__(self, kernel: "RayWalkKernel", stream: Any):
        self._kernel, self._stream = kernel, stream

    def __call__(self, *args):
        return self._kernel._launch(args, self._stream)


class RayWalkKernel:
    """Stand-in for the Numba dispatcher of ``process_rays_kernel_pro_fullColor``."""

    def __init__(self, flags: int = 0, tile_hint: int = 0, runner: Optional[Tuple[int, int, int]] = None,
                 single_lambda: bool = False, threshold: float = 0.0, ray_index_base: int = 0):
        self.flags = flags
        self.tile_hint = tile_hint
        self.runner = runner      # (points P, first cell, num_rays) or None
        self.single_lambda = single_lambda
        self.threshold = threshold
        self.ray_index_base = ray_index_base   # index of ray 0 of a launch in the whole job (shards)
        self._last_sig = self._last_prob = None

    def __getitem__(self, config) -> _Launcher:
        if not isinstance(config, tuple):
            config = (config,)
        if len(config) < 2 or len(config) > 4:
            raise ValueError("launch configuration is [griddim, blockdim(, stream(, sharedmem))]")
        return _Launcher(self, config[2] if len(config) > 2 else None)

    def configured(self, *, strict: Optional[bool] = None, counters: Optional[bool] = None,
                   tile_hint: Optional[int] = None, ray_index_base: Optional[int] = None) -> "RayWalkKernel":
        """A copy with engine options changed (strict = literal thread-per-ray walk;
        ray_index_base = position of this launch's ray 0 in the whole job, for shards)."""
        f = self.flags
        if strict is not None:
            f = (f | _capi.WGRT_FLAG_STRICT) if strict else (f & ~_capi.WGRT_FLAG_STRICT)
        if counters is not None:
            f = (f | _capi.WGRT_FLAG_COUNTERS) if counters else (f & ~_capi.WGRT_FLAG_COUNTERS)
        return RayWalkKernel(f, self.tile_hint if tile_hint is None else tile_hint, self.runner,
                             self.single_lambda, self.threshold,
                             self.ray_index_base if ray_index_base is None else int(ray_index_base))

    def runner_layout(self, points: int, num_rays: int, first_cell: int = 0) -> "RayWalkKernel":
        """Launch on the runner's implicit ray layout (include/wgrt.h): ``x_v`` / ``y_v`` are the
        ``points`` start points, m_v .. delta_phase_v may be ``None``."""
        return RayWalkKernel(self.flags, self.tile_hint, (int(points), int(first_cell), int(num_rays)),
                             self.single_lambda, self.threshold, self.ray_index_base)

    def _launch(self, args, stream):
        lib = _capi.load_library()
        # The runner launches the kernel num_iter times on the SAME device arrays (RUN:169-177): a repeat call whose
        # arguments have the same device addresses, shapes and dtypes IS the same packed problem (it holds nothing
        # else), so the validated one is reused.  No reference to the caller's buffers is kept.
        sig = self._signature(args)
        if sig is not None:
            sig = (self.threshold, self.ray_index_base, self.single_lambda) + sig
        if sig is not None and sig == self._last_sig:
            _capi.check(lib.wgrt_trace_fullcolor(C.byref(self._last_prob), C.c_void_p(_stream_handle(stream))), lib)
            return
        self._last_sig = None
        self._launch_slow(args, stream, lib)
        if sig is not None:
            self._last_sig = sig

    @staticmethod
    def _signature(args):
        """(device address, shape, dtype) of every argument, or None when an argument is a host array (staged anew
        at every launch) or not an array."""
        out = []
        for a in args:
            if a is None or isinstance(a, (int, float)):
                out.append(a)
                continue
            cai = getattr(a, "__cuda_array_interface__", None)
            if cai is None:
                return None
            if cai.get("strides") is not None:
                return None                      # (strided views take the checked path)
            out.append((cai["data"][0], tuple(cai["shape"]), cai["typestr"]))
        return tuple(out)

    def _launch_slow(self, args, stream, lib):
        if self.single_lambda:
            if len(args) != len(_ARG_NAMES) - 1:
                raise TypeError(f"process_rays_kernel_pro takes {len(_ARG_NAMES) - 1} positional "
                                f"arguments ({len(args)} given)")
            args = tuple(args[:8]) + (None,) + tuple(args[8:])      # no lmd_num argument (GRTF:420-427)
        if len(args) != len(_ARG_NAMES):
            raise TypeError(f"process_rays_kernel_pro_fullColor takes {len(_ARG_NAMES)} positional "
                            f"arguments ({len(args)} given)")
        args = tuple(_as_complex128(a, _ARG_NAMES[i]) if i in _LUT_ARGS and a is not None else a
                     for i, a in enumerate(args))
        staged = []        # (host ndarray, device tensor, copy_back)
        dev_args = list(args)
        any_host = False
        for i, (a, nm) in enumerate(zip(args, _ARG_NAMES)):
            if isinstance(a, np.ndarray):
                any_host = True
                if i in _DEAD_ARGS:
                    dev_args[i] = None
                    continue
                import torch
                if not torch.cuda.is_available():
                    raise _capi.WgrtError("no CUDA device: the engine has no CPU fallback")
                if not a.flags.c_contiguous:
                    raise ValueError(f"{nm}: host array must be C-contiguous")
                view = a.view(np.float64) if a.dtype == np.complex128 else a
                if view.dtype == np.uint32:
                    t = torch.from_numpy(view.view(np.int32)).cuda()
                else:
                    t = torch.from_numpy(view).cuda()
                dev_args[i] = _TorchAlias(t, a.shape, a.dtype)
                staged.append((a, t, nm in ("rng_states", "matrix_EB")))
        rp, rc, rn = self.runner if self.runner else (0, 0, None)
        prob, keep = pack_problem(dev_args, host=False, flags=self.flags, tile_hint=self.tile_hint,
                                  runner_points=rp, runner_first_cell=rc, num_rays=rn,
                                  single_lambda=self.single_lambda, threshold=self.threshold,
                                  ray_index_base=self.ray_index_base)
        h = _stream_handle(stream)
        if any_host:
            import torch
            torch.cuda.current_stream().synchronize()      # staging copies ran on torch's stream
        _capi.check(lib.wgrt_trace_fullcolor(C.byref(prob), C.c_void_p(h)), lib)
        self._last_prob = prob
        if staged:
            import torch
            torch.cuda.synchronize()
            for host, t, back in staged:
                if back:
                    flat = t.cpu().numpy()
                    host[...] = flat.view(host.dtype).reshape(host.shape)
        del keep


class _TorchAlias:
    """A torch CUDA tensor presented with the dtype/shape of the host array it stages."""

    def __init__(self, t, shape, dtype):
        self._t = t
        self.__cuda_array_interface__ = {
            "data": (t.data_ptr(), False), "shape": tuple(shape),
            "typestr": np.dtype(dtype).str, "strides": None, "version": 3}


process_rays_kernel_pro_fullColor = RayWalkKernel()
# single-wavelength twin (GRTF:419-831): 32 arguments (no lmd_num), L-less tables, threshold 1e-15
process_rays_kernel_pro = RayWalkKernel(single_lambda=True, threshold=1e-15)
