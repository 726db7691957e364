"""The CPU oracle must reproduce what the REFERENCE file itself produced (oracle/make_golden.py ran
/root/reference/GPU_ray_tracing_functions.py under Numba's CUDA simulator) -- bit for bit."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_bins, input_digest, load_golden_walk

WALKS = ["walk_small", "walk_c1", "walk_mid", "walk_deep", "walk_fine", "walk_thin", "walk_mix", "walk_pol"]


@pytest.mark.parametrize("name", WALKS)
def test_walk_matches_reference(name, oracle):
    scene, g = load_golden_walk(name)
    assert input_digest(scene) == str(g["digest"]), "input generators drifted from the fixture"
    EB = scene.new_matrix_EB()
    rng = scene.rays.rng_states.copy()
    for _ in range(int(g["num_iter"])):
        oracle.trace(*scene.kernel_args(EB, rng))
    assert np.array_equal(rng, g["rng_states"]), "final RNG states differ (a ray drew a different number of times)"
    assert np.array_equal(EB, golden_bins(g)), "eyebox bins differ"
    assert EB.sum() > 0


def test_walk_small_from_stored_inputs(oracle):
    """Self-contained fixture: inputs are stored, so this pins the oracle independently of the generators."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200.GPU_ray_tracing_functions import _ARG_NAMES
    g = np.load(os.path.join(GOLDEN, "walk_small.npz"))
    args = []
    for nm in _ARG_NAMES[:-1]:
        a = g["in_" + nm]
        args.append(float(a) if nm == "n_g" else np.ascontiguousarray(a))
    EB = np.zeros(tuple(g["eb_shape"]), dtype=np.float32)
    rng = args[12].copy()
    args[12] = rng
    oracle.trace(*args, EB)
    assert np.array_equal(rng, g["rng_states"])
    assert np.array_equal(EB, golden_bins(g))


def test_thread_count_invariance(oracle, small_scene):
    res = []
    for nt in (1, 3, 8):
        EB = small_scene.new_matrix_EB(); rng = small_scene.rays.rng_states.copy()
        oracle.trace(*small_scene.kernel_args(EB, rng), num_threads=nt)
        res.append((EB, rng))
    for EB, rng in res[1:]:
        assert np.array_equal(EB, res[0][0]) and np.array_equal(rng, res[0][1])


def test_xorshift_unit(oracle):
    g = np.load(os.path.join(GOLDEN, "units.npz"))
    st, last = oracle.xorshift(g["xs_in"], 7)
    assert np.array_equal(st, g["xs_out"])
    assert np.array_equal(last, g["xs_last"])
    assert np.all((last > 0) & (last < 1))


def test_efield_unit(oracle):
    g = np.load(os.path.join(GOLDEN, "units.npz"))
    out = oracle.efield(g["ef_ete"], g["ef_etm"], g["ef_delta"], g["ef_jones"])
    assert np.array_equal(out, g["ef_out"])          # same libm, same operation order -> bit equal
    assert np.all(np.abs(out[:, 2]) <= np.pi)


@pytest.mark.parametrize("ring", ["IC", "FC", "OC", "eff_reg1", "eff_reg2"])
def test_polygon_unit(ring, oracle):
    g = np.load(os.path.join(GOLDEN, "units.npz"))
    pts = g[ring + "_pts"]
    hit = oracle.locate(g[ring + "_verts"], g[ring + "_off"], pts[:, 0], pts[:, 1])
    assert np.array_equal(hit, g[ring + "_hit"])
    assert (hit >= 0).any() and (hit < 0).any()


def test_eyebox_rectangle_unit(oracle):
    """is_inside_or_on_edge_4d (GRTF:100-108) evaluated by the reference on eyebox rectangles (the design's, a
    rotated one, another vertex order, the unit square) at and around the 1e-9 band of every side / corner."""
    g = np.load(os.path.join(GOLDEN, "units.npz"))
    for k in range(len(g["rect_verts"])):
        pts = g["rect_pts"][k]
        hit = oracle.locate(g["rect_verts"][k], np.array([0, 4]), pts[:, 0], pts[:, 1]) >= 0
        assert np.array_equal(hit, g["rect_hit"][k]), k
        assert hit.any() and not hit.all()


def test_event_trace_is_consistent_with_the_walk(oracle, small_scene):
    """oracle.trace_events (used by tools/triage_mismatch.py): one row per draw, and replaying a ray alone
    leaves the RNG state the launch leaves."""
    EB = small_scene.new_matrix_EB(); rng = small_scene.rays.rng_states.copy()
    c = oracle.trace(*small_scene.kernel_args(EB, rng), counters=True)
    draws = 0
    rng1 = small_scene.rays.rng_states.copy()
    for i in range(0, small_scene.rays.num_rays, 7):
        ev = oracle.trace_events(*small_scene.kernel_args(small_scene.new_matrix_EB(), rng1), idx=i)
        assert rng1[i] == rng[i] and len(ev) >= 1 and ev[0, 0] == -1
        assert np.all((ev[:, 1] > 0) & (ev[:, 1] < 1)) and np.all(ev[:, 3] >= ev[:, 2])
        draws += len(ev)
    assert 0 < draws <= c["draws"]


def test_counters_consistent(oracle, small_scene):
    EB = small_scene.new_matrix_EB(); rng = small_scene.rays.rng_states.copy()
    c = oracle.trace(*small_scene.kernel_args(EB, rng), counters=True)
    assert c["rays"] == small_scene.rays.num_rays
    assert c["draws"] == c["draw2"] + c["draw3"]
    assert c["efield"] == 2 * c["draw2"] + 3 * c["draw3"]
    assert c["deposits"] == int(EB.sum())


def _single_lambda_case():
    import ast
    from oracle import make_golden
    g = np.load(os.path.join(GOLDEN, "walk_single_lambda.npz"))
    recipe = ast.literal_eval(str(g["recipe"]))
    scene = make_golden.scene_from_recipe(recipe)
    assert input_digest(scene) == str(g["digest"])
    lam = int(g["lam"])
    EB = np.zeros(tuple(g["eb_shape"]), dtype=np.float32)
    rng = scene.rays.rng_states[scene.rays.lmd_num == lam].copy()
    args, _ = make_golden.single_lambda_args(scene, lam, EB, rng)
    want = np.zeros(EB.size, np.float32); want[g["eb_index"]] = g["eb_value"]
    return g, args, EB, rng, want.reshape(EB.shape)


def test_single_lambda_twin_matches_reference(oracle):
    """process_rays_kernel_pro (GRTF:419-831): 32 arguments, no wavelength axis, threshold 1e-15."""
    g, args, EB, rng, want = _single_lambda_case()
    oracle.trace(*args, single_lambda=True, threshold=1e-15)
    assert np.array_equal(rng, g["rng_states"])
    assert np.array_equal(EB, want) and EB.sum() > 0


@pytest.mark.skipif(not os.path.exists("/root/reference/GPU_ray_tracing_functions.py"),
                    reason="the reference sources only exist in the build container")
def test_reference_kernels_compile_to_ptx():
    """oracle/build_ref_ptx.py: the reference's own Numba kernels -> PTX (what the GPU box launches as
    the reference).  Checks the manifest: both kernels, Numba's kernel ABI laid out for 33 / 32 arguments,
    and that the PTX really is the reference's code (its module and function names are mangled into the
    entry point)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, "oracle", "build_ref_ptx.py")], check=True,
                   stdout=subprocess.DEVNULL, env={k: v for k, v in os.environ.items() if k != "NUMBA_ENABLE_CUDASIM"})
    with open(os.path.join(root, "oracle", "_ref", "manifest.json")) as f:
        man = json.load(f)
    full, pro = man["kernels"]["process_rays_kernel_pro_fullColor"], man["kernels"]["process_rays_kernel_pro"]
    assert len(full["params"]) == 33 and len(pro["params"]) == 32
    assert "GPU_ray_tracing_functions" in full["entry"] and "process_rays_kernel_pro_fullColor" in full["entry"]
    assert full["params"][18] == {"kind": "scalar", "dtype": "float64"}          # n_g
    assert full["params"][32] == {"kind": "array", "ndim": 5, "dtype": "float32"}  # matrix_EB
    for k in (full, pro):
        ptx = open(os.path.join(root, "oracle", "_ref", k["ptx"])).read()
        assert ".visible .entry " + k["entry"] in ptx and "fma.rn.f64" in ptx
