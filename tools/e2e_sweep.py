"""Wall time of the runner-level host job (runner.trace_full_color, C2, pinned host buffers) for a few pipeline
chunk counts: `WGRT_LIB=<build> python tools/e2e_sweep.py [--iters 10] [--chunks 0,4,8,12]` (0 = the library's choice)."""
import argparse, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import runner, synthetic_inputs as si

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10); ap.add_argument("--chunks", default="0,4,8,12")
ap.add_argument("--rays", type=int, default=5000)
a = ap.parse_args()
scene = si.make_scene(100, 75, a.rays, seed=1, build_rays=False)
pts = si.points_in_disc(scene.geom["IC"], a.rays // 2, 2025)
def pin(x):
    t = torch.from_numpy(np.ascontiguousarray(x).view(np.float64) if x.dtype == np.complex128 else np.ascontiguousarray(x)).pin_memory()
    v = t.numpy()
    return t, (v.view(np.complex128).reshape(x.shape) if x.dtype == np.complex128 else v)
keep, geom, luts = [], {}, {}
for src, dst in ((scene.geom, geom), (scene.luts, luts)):
    for k, v in src.items():
        t, view = pin(v); keep.append(t); dst[k] = view
t, eb = pin(scene.new_matrix_EB()); keep.append(t)
for ch in [int(c) for c in a.chunks.split(",")]:
    if ch: os.environ["WGRT_HOST_CHUNKS"] = str(ch)
    else: os.environ.pop("WGRT_HOST_CHUNKS", None)
    for n_it in (1, a.iters):
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            runner.trace_full_color(pts, geom, scene.n_g, luts, a.rays, num_iter=n_it, matrix_EB=eb, bins_start_zero=True)
            best = min(best, time.perf_counter() - t0)
        print(f"chunks {ch:3d}  num_iter {n_it:3d}  {best * 1e3:8.2f} ms  ({best * 1e3 / n_it:6.2f} ms per launch)", flush=True)
