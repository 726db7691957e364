"""north_star: "the host functions keep their Python signatures, so gpu_ray_tracing_pro_fullColor.py runs
unchanged".  This test runs the DEVICE SECTION of the reference runner the way the runner runs it --
Numba ``DeviceNDArray``s made by ``cuda.to_device`` (RUN:40-57, 145-159), the launch
``GRTF.process_rays_kernel_pro_fullColor[blocks_per_grid, threads_per_block](...33 args...)`` repeated
``num_iter`` times (RUN:168-177), ``cuda.synchronize()`` (RUN:178) and ``copy_to_host()`` (RUN:185) --
with ``GRTF`` being this package's module, and compares with the CPU oracle.  No torch object is involved:
Numba owns every buffer, and the engine's launch on the legacy default stream is fenced by Numba's
``cuda.synchronize()`` exactly like Numba's own launch.

Runs in a fresh interpreter: other tests import oracle/make_golden.py, which switches Numba to its CPU
simulator for the whole process.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RUNNER_DEVICE_SECTION = r'''
import sys
import numpy as np
sys.path.insert(0, ROOT)
from numba import cuda
import gpu_ray_tracing_for_waveguide_based_ar_display_b200.GPU_ray_tracing_functions as GRTF
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import synthetic_inputs as si
from oracle import oracle

num_FOV_x, num_FOV_y, num_rays_per_FoV, num_iter = 8, 6, 1000, 4
scene = si.make_scene(num_FOV_x, num_FOV_y, num_rays_per_FoV, seed=91)
g, l, r = scene.geom, scene.luts, scene.rays
if COMPLEX64:
    l = {k: v.astype(np.complex64) for k, v in l.items()}
n_g = scene.n_g
matrix_EB = np.zeros(scene.eb_shape, dtype=np.float32)                       # RUN:37

# RUN:40-57
d_IC = cuda.to_device(g["IC"]); d_FC = cuda.to_device(g["FC"]); d_FC_offset = cuda.to_device(g["FC_offset"])
d_OC = cuda.to_device(g["OC"]); d_OC_offset = cuda.to_device(g["OC_offset"])
d_eff_reg1 = cuda.to_device(g["eff_reg1"]); d_eff_reg2 = cuda.to_device(g["eff_reg2"])
d_eff_reg_FOV = cuda.to_device(g["eff_reg_FOV"]); d_eff_reg_FOV_range = cuda.to_device(g["eff_reg_FOV_range"])
d_lut_ic1 = cuda.to_device(l["lut_ic1"]); d_lut_ic2 = cuda.to_device(l["lut_ic2"]); d_lut_ic3 = cuda.to_device(l["lut_ic3"])
d_lut_fc1 = cuda.to_device(l["lut_fc1"]); d_lut_fc2 = cuda.to_device(l["lut_fc2"])
d_lut_oc1 = cuda.to_device(l["lut_oc1"]); d_lut_oc2 = cuda.to_device(l["lut_oc2"])
d_lut_TIR = cuda.to_device(g["lut_TIR"]); d_lut_gap = cuda.to_device(g["lut_gap"])

# RUN:145-159
num_rays = r.num_rays
d_x = cuda.to_device(r.x); d_y = cuda.to_device(r.y)
d_gap_x = cuda.to_device(r.gap_x); d_gap_y = cuda.to_device(r.gap_y)
d_pol = cuda.to_device(r.pol); d_azi = cuda.to_device(r.azi)
d_m = cuda.to_device(r.m); d_n = cuda.to_device(r.n); d_lmd_num = cuda.to_device(r.lmd_num)
d_te = cuda.to_device(r.te); d_tm = cuda.to_device(r.tm); d_delta_phase = cuda.to_device(r.delta_phase)
rng_states = (np.uint32(0x9E3779B9) * (np.arange(num_rays, dtype=np.uint32) + np.uint32(1)))   # RUN:158
d_rng_states = cuda.to_device(rng_states)
d_matrix_EB = cuda.to_device(matrix_EB)
threads_per_block = 256
blocks_per_grid = (num_rays + threads_per_block - 1) // threads_per_block

# RUN:168-178
for _ in range(num_iter):
    GRTF.process_rays_kernel_pro_fullColor[blocks_per_grid, threads_per_block](
        d_x, d_y, d_gap_x, d_gap_y, d_pol, d_azi, d_m, d_n, d_lmd_num,
        d_te, d_tm, d_delta_phase, d_rng_states,
        d_IC, d_FC, d_FC_offset, d_OC, d_OC_offset, n_g,
        d_eff_reg1, d_eff_reg2, d_eff_reg_FOV, d_eff_reg_FOV_range,
        d_lut_ic1, d_lut_ic2, d_lut_ic3, d_lut_fc1, d_lut_fc2, d_lut_oc1, d_lut_oc2,
        d_lut_TIR, d_lut_gap, d_matrix_EB)
cuda.synchronize()
matrix_EB = d_matrix_EB.copy_to_host()                                        # RUN:185
rng_out = d_rng_states.copy_to_host()

# a second job on a Numba stream, fenced by that stream only
s = cuda.stream()
d_rng2 = cuda.to_device(rng_states, stream=s); d_EB2 = cuda.to_device(np.zeros_like(matrix_EB), stream=s)
GRTF.process_rays_kernel_pro_fullColor[blocks_per_grid, threads_per_block, s](
    d_x, d_y, d_gap_x, d_gap_y, d_pol, d_azi, d_m, d_n, d_lmd_num, d_te, d_tm, d_delta_phase, d_rng2,
    d_IC, d_FC, d_FC_offset, d_OC, d_OC_offset, n_g, d_eff_reg1, d_eff_reg2, d_eff_reg_FOV, d_eff_reg_FOV_range,
    d_lut_ic1, d_lut_ic2, d_lut_ic3, d_lut_fc1, d_lut_fc2, d_lut_oc1, d_lut_oc2, d_lut_TIR, d_lut_gap, d_EB2)
EB2 = d_EB2.copy_to_host(stream=s); rng2 = d_rng2.copy_to_host(stream=s)
s.synchronize()

# the oracle on the same host inputs (complex64 tables widened exactly as the engine widens them)
lw = {k: v.astype(np.complex128) for k, v in l.items()}
scene.luts = lw
EB_o = np.zeros(scene.eb_shape, dtype=np.float32); rng_o = rng_states.copy()
EB_1 = None
for it in range(num_iter):
    oracle.trace(*scene.kernel_args(EB_o, rng_o))
    if it == 0:
        EB_1, rng_1 = EB_o.copy(), rng_o.copy()
assert np.array_equal(rng_out, rng_o), "rng_states differ from the oracle"
assert np.array_equal(matrix_EB, EB_o), "matrix_EB differs from the oracle"
assert np.array_equal(rng2, rng_1) and np.array_equal(EB2, EB_1), "launch on a Numba stream differs"
assert matrix_EB.sum() > 100
print("numba runner ok", int(matrix_EB.sum()))
'''


@pytest.mark.parametrize("complex64", [False, True])
def test_runner_device_section_on_numba_device_arrays(complex64):
    code = f"ROOT = {ROOT!r}\nCOMPLEX64 = {complex64}\n" + RUNNER_DEVICE_SECTION
    env = {k: v for k, v in os.environ.items() if k != "NUMBA_ENABLE_CUDASIM"}
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], cwd=ROOT, env=env, capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0 and "numba runner ok" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
