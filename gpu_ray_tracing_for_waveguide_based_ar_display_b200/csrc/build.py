"""Build libwgrt.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree).

    python -m gpu_ray_tracing_for_waveguide_based_ar_display_b200.csrc.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libwgrt.so")
OUT_CHECKED = os.path.join(PKG, "libwgrt_checked.so")
OBJ_DIR = os.path.join(HERE, "build")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# translation unit -> extra flags.  The strict walk must not contract multiply-adds: it is the
# literal restatement that is compared with the CPU oracle.
UNITS = {
    "wgrt_strict.cu": ["-fmad=false"],
    "wgrt_legacy.cu": ["-fmad=false"],
    "wgrt_index.cu": [f"-D{k}={os.environ[k]}" for k in ("WGRT_ZONE_REFINE",) if k in os.environ],
    "wgrt_eval.cu": [],
    "wgrt_walk.cu": [f"-D{k}={os.environ[k]}" for k in ("WGRT_QUEUE_CAP", "WGRT_ZONE_REFINE", "WGRT_JSM_WARPS") if k in os.environ],
    "wgrt_api.cu": [f"-D{k}={os.environ[k]}" for k in ("WGRT_ZONE_REFINE",) if k in os.environ],
}
HEADERS = ["wgrt_device.cuh", "wgrt_region.cuh", os.path.join("..", "..", "include", "wgrt.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_compiler_args():
    # /opt/gcc wrappers in this image lack some spec files; the system g++ is complete.
    for cand in ("/usr/bin/g++",):
        if os.path.exists(cand):
            return ["-ccbin", cand]
    return []


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, checked: bool = False) -> str:
    """``checked=True`` builds libwgrt_checked.so: the same sources with -DWGRT_CHECKED (bounds assertions
    in the production walk, see wgrt_device.cuh), used by tests/test_gpu_checked_build.py."""
    obj_dir = os.path.join(OBJ_DIR, "checked") if checked else OBJ_DIR
    out = OUT_CHECKED if checked else OUT
    if os.environ.get("WGRT_BUILD_TAG"):     # experiment builds: libwgrt_<tag>.so next to the product library
        tag = os.environ["WGRT_BUILD_TAG"]
        obj_dir = os.path.join(OBJ_DIR, "exp_" + tag)
        out = os.path.join(PKG, f"libwgrt_{tag}.so")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    hdrs = [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    for src, extra in UNITS.items():
        s = os.path.join(HERE, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc, *ARCH, *COMMON, *_host_compiler_args(), *extra, *(["-DWGRT_CHECKED"] if checked else []),
                   "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
    if force or _stale(out, objs):
        cmd = [nvcc, *ARCH, "-shared", *_host_compiler_args(), "-o", out, *objs, "-lcudart"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build(force="--force" in sys.argv, verbose=False, checked=True))
