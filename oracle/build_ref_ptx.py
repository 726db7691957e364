"""Compile the REFERENCE's own Numba-CUDA kernels to PTX (test / baseline infrastructure only).

The reference hot path is ``@cuda.jit`` Python (GPU_ray_tracing_functions.py:419-831, 833-1246); its
"build" is Numba's JIT: Python -> NVVM IR -> PTX, which the CUDA driver then JITs for the device.
This script runs exactly that first half here, on the unmodified file where it lies under
/root/reference, and writes ONLY compiler output (PTX text + a manifest) into ``oracle/_ref/``
(git-ignored, travels to the GPU box).  No reference source is copied.  On the B200 box
``oracle/ref_numba_cuda.py`` hands the PTX to the driver (``cuModuleLoadData``), which is what Numba
itself would do at the first launch, so the thing timed and compared there IS the reference kernel.

No GPU is needed: Numba asks the current device only for its compute capability when it compiles the
nested device functions (numba/cuda/dispatcher.py: ``compile_device``), so that one lookup is
answered with a stand-in object reporting cc 10.0 (B200).  Numba 0.65's NVVM then targets the
highest architecture it knows at or below that (``.target sm_90``), as it does on a real B200.

    python oracle/build_ref_ptx.py          # writes oracle/_ref/*.ptx + manifest.json (~20 s)
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_DIR = os.environ.get("WGRT_REFERENCE_DIR", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")


def _import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.path", "matplotlib.colors"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.Path = object
            mod.LogNorm = object
            sys.modules[name] = mod
    if os.environ.get("NUMBA_ENABLE_CUDASIM") == "1":
        raise RuntimeError("build_ref_ptx needs the real Numba CUDA target, not the simulator")
    sys.path.insert(0, REFERENCE_DIR)
    import GPU_ray_tracing_functions as GRTF
    return GRTF


def _signatures():
    from numba import types as nt

    def arr(dt, nd):
        return nt.Array(dt, nd, "C")

    f32 = arr(nt.float32, 1)
    c128 = nt.complex128
    geom = (arr(nt.float64, 2), arr(nt.float64, 2), arr(nt.int64, 1), arr(nt.float64, 2), arr(nt.int64, 1),
            nt.float64, arr(nt.float64, 2), arr(nt.float64, 2), arr(nt.float64, 4), arr(nt.float64, 3))
    # the argument types RUN:40-57, 145-159 produce (float32 ray arrays, uint32 RNG, float64 geometry,
    # int64 offsets, complex128 LUTs, float32 bins)
    full = (f32,) * 12 + (arr(nt.uint32, 1),) + geom + (
        arr(c128, 4), arr(c128, 4), arr(c128, 4), arr(c128, 5), arr(c128, 5), arr(c128, 5), arr(c128, 5),
        arr(nt.float64, 4), arr(nt.float64, 4), arr(nt.float32, 5))
    # single-wavelength twin (GRTF:419-428): no lmd_num array, LUTs and tables lose the wavelength axis
    pro = (f32,) * 11 + (arr(nt.uint32, 1),) + geom + (
        arr(c128, 3), arr(c128, 3), arr(c128, 3), arr(c128, 4), arr(c128, 4), arr(c128, 4), arr(c128, 4),
        arr(nt.float64, 3), arr(nt.float64, 3), arr(nt.float32, 4))
    # the legacy energy-splitting kernel and its compaction kernel (GRTF:178-417): float64 ray rows, int32 child
    # counter, single-wavelength tables with 3 orders per group (the reference ships neither a driver nor LUT
    # files for them; these are the argument types its signature implies)
    legacy = (arr(nt.float64, 2), nt.int64, arr(nt.int32, 1), nt.int64,
              arr(nt.float64, 2), arr(nt.float64, 2), arr(nt.int64, 1), arr(nt.float64, 2), arr(nt.int64, 1),
              arr(nt.float64, 2), arr(nt.float64, 2), arr(nt.float64, 4), arr(nt.float64, 3),
              arr(c128, 3), arr(c128, 3), arr(c128, 4), arr(c128, 4), arr(c128, 4),
              arr(nt.float64, 3), arr(nt.float64, 3), arr(nt.float32, 4))
    pack = (arr(nt.float64, 2), arr(nt.float64, 2), nt.int64, arr(nt.int32, 1))
    return {"process_rays_kernel_pro_fullColor": full, "process_rays_kernel_pro": pro,
            "process_rays_kernel": legacy, "pack_active_to_front": pack}


def _param_layout(sig):
    """Numba kernel ABI: an array is (meminfo, parent, nitems, itemsize, data, shape[nd], strides[nd])."""
    from numba import types as nt
    out = []
    for t in sig:
        if isinstance(t, nt.Array):
            out.append({"kind": "array", "ndim": t.ndim, "dtype": str(t.dtype)})
        else:
            out.append({"kind": "scalar", "dtype": str(t)})
    return out


def build(force: bool = False) -> str:
    manifest_path = os.path.join(OUT_DIR, "manifest.json")
    ref_file = os.path.join(REFERENCE_DIR, "GPU_ray_tracing_functions.py")
    if not os.path.exists(ref_file):
        raise FileNotFoundError(ref_file)
    with open(ref_file, "rb") as f:
        ref_sha = hashlib.sha256(f.read()).hexdigest()
    if not force and os.path.exists(manifest_path):
        with open(manifest_path) as f:
            if json.load(f).get("reference_sha256") == ref_sha:
                return manifest_path
    import numba
    import numba.cuda.dispatcher as dispatcher
    from numba import cuda

    class _B200:
        compute_capability = (10, 0)

    dispatcher.get_current_device = lambda: _B200()
    GRTF = _import_reference()
    os.makedirs(OUT_DIR, exist_ok=True)
    manifest = {"reference_sha256": ref_sha, "numba": numba.__version__, "kernels": {}}
    for name, sig in _signatures().items():
        ptx, _ = cuda.compile_ptx(getattr(GRTF, name).py_func, sig, cc=(10, 0))
        entry = [ln.split()[2].rstrip("(") for ln in ptx.splitlines() if ln.startswith(".visible .entry")]
        assert len(entry) == 1, entry
        out = os.path.join(OUT_DIR, name + ".ptx")
        with open(out, "w") as f:
            f.write(ptx)
        target = [ln for ln in ptx.splitlines() if ln.startswith(".target")][0]
        manifest["kernels"][name] = {"ptx": os.path.basename(out), "entry": entry[0], "target": target,
                                     "params": _param_layout(sig)}
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest_path


if __name__ == "__main__":
    print(build(force=True))
