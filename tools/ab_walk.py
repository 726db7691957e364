"""A/B timing of several builds of libwgrt on ONE scene in ONE process (experiment builds: WGRT_BUILD_TAG=<tag>
python -m ...csrc.build -> libwgrt_<tag>.so).  Prints ms per C2 launch (best / median of --iters) per library and
checks that every build deposits the same number of rays.

    python tools/ab_walk.py [--rays 5000] [--iters 5] lib1.so lib2.so ...
"""
import argparse, ctypes as C, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi, synthetic_inputs as si

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=100); ap.add_argument("--ny", type=int, default=75)
ap.add_argument("--rays", type=int, default=5000); ap.add_argument("--iters", type=int, default=5)
ap.add_argument("libs", nargs="+")
a = ap.parse_args()
scene = si.make_scene(a.nx, a.ny, a.rays, seed=1)

def to_dev(x):
    if isinstance(x, np.ndarray):
        v = x.view(np.float64) if x.dtype == np.complex128 else x
        t = torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).cuda()
        return GRTF._TorchAlias(t, x.shape, x.dtype)
    return x
dev = [to_dev(x) for x in scene.kernel_args(scene.new_matrix_EB())]
rng0 = dev[12]._t.clone()
prob, keep = GRTF.pack_problem(dev, host=False)
ref = None
for path in a.libs:
    lib = C.CDLL(os.path.abspath(path))
    lib.wgrt_trace_fullcolor.restype = C.c_int
    lib.wgrt_trace_fullcolor.argtypes = [C.c_void_p, C.c_void_p]
    lib.wgrt_last_error.restype = C.c_char_p
    if lib.wgrt_problem_size() != C.sizeof(prob):
        print(f"{os.path.basename(path):28s} struct size mismatch, skipped"); continue
    times = []
    for it in range(a.iters + 2):
        dev[12]._t.copy_(rng0); dev[32]._t.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        rc = lib.wgrt_trace_fullcolor(C.byref(prob), None)
        e1.record(); torch.cuda.synchronize()
        if rc != 0:
            print(path, "error", lib.wgrt_last_error()); break
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    dep = float(dev[32]._t.sum(dtype=torch.float64).item())
    ref = dep if ref is None else ref
    print(f"{os.path.basename(path):28s} best {min(times):7.3f}  median {float(np.median(times)):7.3f} ms   deposits {dep:.0f}"
          f"{'' if dep == ref else '  <-- DIFFERS'}", flush=True)
    lib.wgrt_release()
