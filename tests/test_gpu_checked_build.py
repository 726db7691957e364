"""The production walk under bounds assertions (libwgrt_checked.so, -DWGRT_CHECKED).

compute-sanitizer is not allowed on the GPU pool ("runs under it have left GPUs needing a reset"), so the
memcheck of SURVEY.md section 4 is replaced by assertions of our own: in the checked build every index the
walk forms into shared memory (event table rows, survivor stack slots, state info), the Jones scratch, the
atlas levels, the per-set region grids / vertex arrays and the ray / RNG arrays is asserted to be in range
and violations are counted instead of corrupting memory.  The parity cases below must produce the
reference's results with ZERO violations -- at the default tie tolerance and with a large share of the rays
pushed through the literal redo kernel, on explicit ray arrays and on the pipelined runner layout.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import golden_bins, load_golden_walk

pytestmark = pytest.mark.gpu

from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi, synthetic_inputs as si


@pytest.fixture(scope="module")
def checked():
    if not os.path.exists(_capi.CHECKED_LIB_PATH):
        pytest.fail("libwgrt_checked.so is not built (python -m ...csrc.build)")
    lib = _capi.load_library(_capi.CHECKED_LIB_PATH)
    n = C.c_uint64()
    _capi.check(lib.wgrt_debug_check_failures(C.byref(n), 1), lib)
    yield lib
    _capi.check(lib.wgrt_release(), lib)


def failures(lib):
    n = C.c_uint64()
    _capi.check(lib.wgrt_debug_check_failures(C.byref(n), 1), lib)
    return int(n.value)


def test_plain_build_has_no_assertions():
    lib = _capi.load_library()
    n = C.c_uint64()
    assert lib.wgrt_debug_check_failures(C.byref(n), 0) == -4        # WGRT_ERR_UNSUPPORTED


@pytest.mark.parametrize("tie", [-1.0, 0.05])
def test_golden_walks_without_violations(checked, tie):
    _capi.check(checked.wgrt_debug_set_tie_tolerance(tie), checked)
    try:
        for name in ("walk_c1", "walk_mid", "walk_deep", "walk_fine", "walk_thin", "walk_mix", "walk_pol"):
            scene, g = load_golden_walk(name)
            EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
            prob, keep = GRTF.pack_problem(scene.kernel_args(EB, rng), host=True)
            _capi.check(checked.wgrt_trace_fullcolor_host(C.byref(prob), int(g["num_iter"]), None), checked)
            assert np.array_equal(rng, g["rng_states"]) and np.array_equal(EB, golden_bins(g)), name
            assert failures(checked) == 0, name
    finally:
        _capi.check(checked.wgrt_debug_set_tie_tolerance(-1.0), checked)


@pytest.mark.parametrize("chunks", [1, 3])
def test_runner_layout_pipeline_without_violations(checked, chunks, oracle, monkeypatch):
    """Runner layout (implicit rays), shuffled explicit rays, ragged tiles and a 1.2 M-ray launch that fills
    every resident warp of the GPU."""
    monkeypatch.setenv("WGRT_HOST_CHUNKS", str(chunks))
    rpc = 200
    scene = si.make_scene(9, 4, rpc, seed=41)
    pts = si.points_in_disc(scene.geom["IC"], rpc // 2, 42)
    scene.rays = si.build_ray_set(pts, 9, 4, 3, rpc)
    EB_o = scene.new_matrix_EB(); rng_o = scene.rays.rng_states.copy()
    for _ in range(2):
        oracle.trace(*scene.kernel_args(EB_o, rng_o))
    px = np.ascontiguousarray(pts[:, 0], dtype=np.float32); py = np.ascontiguousarray(pts[:, 1], dtype=np.float32)
    EB = scene.new_matrix_EB()
    a = list(scene.kernel_args(EB, None))
    a[0], a[1] = px, py
    for i in range(2, 13):
        a[i] = None
    prob, keep = GRTF.pack_problem(a, host=True, flags=_capi.WGRT_FLAG_BINS_ZERO, runner_points=rpc // 2,
                                   num_rays=scene.rays.num_rays)
    _capi.check(checked.wgrt_trace_fullcolor_host(C.byref(prob), 2, None), checked)
    assert np.array_equal(EB, EB_o) and failures(checked) == 0
    # shuffled explicit arrays: runs of length 1, tiles cut anywhere
    perm = np.random.default_rng(0).permutation(scene.rays.num_rays)
    rays = scene.rays
    scene.rays = si.RaySet(*(x[perm].copy() for x in rays.arrays()), rays.rng_states[perm].copy())
    EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
    prob, keep = GRTF.pack_problem(scene.kernel_args(EB, rng), host=True, tile_hint=97)
    _capi.check(checked.wgrt_trace_fullcolor_host(C.byref(prob), 1, None), checked)
    EB1 = scene.new_matrix_EB(); rng1 = scene.rays.rng_states.copy()
    oracle.trace(*scene.kernel_args(EB1, rng1))
    assert np.array_equal(EB, EB1) and np.array_equal(rng, rng1) and failures(checked) == 0
    # a launch large enough to occupy every resident warp (fast vs strict, both in the checked library)
    big = si.make_scene(16, 12, 2000, seed=77)
    res = []
    for flags in (0, _capi.WGRT_FLAG_STRICT):
        EB = big.new_matrix_EB(); rng = big.rays.rng_states.copy()
        prob, keep = GRTF.pack_problem(big.kernel_args(EB, rng), host=True, flags=flags)
        _capi.check(checked.wgrt_trace_fullcolor_host(C.byref(prob), 1, None), checked)
        res.append((EB, rng))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert failures(checked) == 0 and res[0][0].sum() > 0
