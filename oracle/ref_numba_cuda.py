"""Launch the REFERENCE's own Numba-CUDA kernels on the GPU (test / baseline infrastructure only).

``oracle/build_ref_ptx.py`` compiles the unmodified reference kernels with Numba to PTX
(``oracle/_ref/*.ptx``).  This module gives that PTX to the CUDA driver (``cuModuleLoadData`` -- the
driver JITs it for the device, as it does when Numba launches the kernel) and launches it with
Numba's kernel ABI and the runner's launch shape (``threads_per_block = 256``,
``blocks = ceil(N / 256)``, gpu_ray_tracing_pro_fullColor.py:160, 167, 170-177).  It is the GPU-side
oracle for inputs far too large for the CPU simulator and the "reference Numba-CUDA on one B200"
baseline that BASELINE.json asks to be timed next to the engine.  Nothing in the product imports it.

Arguments are the kernel's positional arguments as device buffers (anything with
``__cuda_array_interface__``; C-contiguous) and the Python float ``n_g``.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_cuda = None
_modules = {}


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "manifest.json"))


def _driver():
    global _cuda
    if _cuda is None:
        _cuda = C.CDLL("libcuda.so.1")
        _check(_cuda.cuInit(0), "cuInit")
    return _cuda


def _check(rc, what):
    if rc != 0:
        msg = C.c_char_p()
        try:
            _cuda.cuGetErrorString(rc, C.byref(msg))
        except Exception:
            pass
        raise RuntimeError(f"{what} failed: CUDA driver error {rc} {msg.value.decode() if msg.value else ''}")


def manifest():
    with open(os.path.join(REF_DIR, "manifest.json")) as f:
        return json.load(f)


def _function(name: str):
    """cuModule + cuFunction of one reference kernel in the CURRENT context (the primary context
    torch / the runtime API made current)."""
    cu = _driver()
    ctx = C.c_void_p()
    _check(cu.cuCtxGetCurrent(C.byref(ctx)), "cuCtxGetCurrent")
    if not ctx.value:
        raise RuntimeError("no current CUDA context: touch the device first (torch.cuda.init())")
    key = (name, ctx.value)
    if key not in _modules:
        info = manifest()["kernels"][name]
        with open(os.path.join(REF_DIR, info["ptx"]), "rb") as f:
            ptx = f.read() + b"\0"
        mod = C.c_void_p()
        _check(cu.cuModuleLoadData(C.byref(mod), ptx), "cuModuleLoadData(reference PTX)")
        fn = C.c_void_p()
        _check(cu.cuModuleGetFunction(C.byref(fn), mod, info["entry"].encode()), "cuModuleGetFunction")
        _modules[key] = (mod, fn, info)
    return _modules[key]


def _array_params(obj, want):
    cai = obj.__cuda_array_interface__
    shape = tuple(int(s) for s in cai["shape"])
    dtype = np.dtype(cai["typestr"])
    if len(shape) != want["ndim"] or str(dtype) != want["dtype"]:
        raise TypeError(f"reference kernel expects {want['dtype']}[{want['ndim']}D], got {dtype}{shape}")
    strides, acc = [], dtype.itemsize
    for s in reversed(shape):
        strides.append(acc)
        acc *= s
    strides.reverse()
    if cai.get("strides") is not None and tuple(cai["strides"]) != tuple(strides) and all(s > 1 for s in shape):
        raise ValueError("reference kernel was compiled for C-contiguous arrays")
    nitems = int(np.prod(shape)) if shape else 1
    vals = [C.c_void_p(0), C.c_void_p(0), C.c_size_t(nitems), C.c_size_t(dtype.itemsize),
            C.c_void_p(int(cai["data"][0] or 0))]
    vals += [C.c_ssize_t(s) for s in shape] + [C.c_ssize_t(s) for s in strides]
    return vals


def launch(name: str, args: Sequence, stream: int = 0, threads_per_block: int = 256, n: int = None) -> None:
    """``GRTF.<name>[ceil(N/256), 256, stream](*args)`` with the reference's compiled kernel (N = length of the
    first array unless ``n`` is given)."""
    cu = _driver()
    _, fn, info = _function(name)
    if len(args) != len(info["params"]):
        raise TypeError(f"{name} takes {len(info['params'])} arguments, got {len(args)}")
    vals = []
    for a, want in zip(args, info["params"]):
        if want["kind"] == "array":
            vals += _array_params(a, want)
        elif want["dtype"].startswith("int"):
            vals.append(C.c_int64(int(a)))
        else:
            vals.append(C.c_double(float(a)))
    if n is None:
        n = int(args[0].__cuda_array_interface__["shape"][0])
    if n == 0:
        return
    params = (C.c_void_p * len(vals))(*[C.cast(C.pointer(v), C.c_void_p) for v in vals])
    blocks = (n + threads_per_block - 1) // threads_per_block
    _check(cu.cuLaunchKernel(fn, blocks, 1, 1, threads_per_block, 1, 1, 0, C.c_void_p(stream), params, None),
           f"cuLaunchKernel({name})")


def function_attributes(name: str) -> dict:
    """Registers / local memory the driver's JIT gave the reference kernel on this device."""
    cu = _driver()
    _, fn, _ = _function(name)
    out = {}
    for label, attr in (("max_threads_per_block", 0), ("shared_bytes", 1), ("local_bytes", 3), ("registers", 4),
                        ("ptx_version", 5), ("binary_version", 6)):
        v = C.c_int()
        _check(cu.cuFuncGetAttribute(C.byref(v), attr, fn), "cuFuncGetAttribute")
        out[label] = v.value
    return out
