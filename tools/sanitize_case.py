"""The launches compute-sanitizer is run on (tools/run_sanitizers.sh): BASELINE config 1 through the host
entry with explicit ray arrays, and a 3-chunk runner-layout job (two walk streams sharing the region index,
H2D / D2H streams) with two launches per chunk.  NumPy + ctypes only (no torch: its start-up under the
sanitizer takes minutes); results are checked against the CPU oracle."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi, runner, synthetic_inputs as si
from oracle import oracle

lib = _capi.load_library()
tie = float(os.environ.get("WGRT_SANITIZE_TIE_TOL", "-1"))
_capi.check(lib.wgrt_debug_set_tie_tolerance(tie), lib)   # > 0: also drives rays through the redo kernel

# (1) config 1: 5 x 5 FoV cells, 532 nm only, 64 rays per cell, explicit ray arrays
scene = si.make_scene(5, 5, 64, seed=11, lmd_subset=[1])
EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
prob, keep = GRTF.pack_problem(scene.kernel_args(EB, rng), host=True)
_capi.check(lib.wgrt_trace_fullcolor_host(C.byref(prob), 2, None), lib)
EB_o = scene.new_matrix_EB(); rng_o = scene.rays.rng_states.copy()
for _ in range(2):
    oracle.trace(*scene.kernel_args(EB_o, rng_o))
assert np.array_equal(EB, EB_o) and np.array_equal(rng, rng_o)
print("config 1 ok:", scene.rays.num_rays, "rays x 2 launches,", int(EB.sum()), "deposits")

# (2) runner layout, 3 pipeline chunks, 2 launches per chunk
os.environ["WGRT_HOST_CHUNKS"] = "3"
rpc = 200
scene = si.make_scene(9, 4, rpc, seed=41)
pts = si.points_in_disc(scene.geom["IC"], rpc // 2, 42)
scene.rays = si.build_ray_set(pts, 9, 4, 3, rpc)
EB = runner.trace_full_color(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=2)
EB_o = scene.new_matrix_EB(); rng_o = scene.rays.rng_states.copy()
for _ in range(2):
    oracle.trace(*scene.kernel_args(EB_o, rng_o))
assert np.array_equal(EB, EB_o)
print("3-chunk host pipeline ok:", scene.rays.num_rays, "rays x 2 launches,", int(EB.sum()), "deposits")
_capi.check(lib.wgrt_release(), lib)
