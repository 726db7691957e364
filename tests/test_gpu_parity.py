"""GPU parity tests: the CUDA engine (through the C ABI / the reference-shaped kernel object) against
the CPU oracle, the committed golden vectors, and size-independent invariants.  Bar: bit-exact
matrix_EB (integer counts stored in float32) and bit-exact final rng_states (=> identical number of
draws per ray)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_bins, load_golden_walk

pytestmark = pytest.mark.gpu

from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi, synthetic_inputs as si
from gpu_ray_tracing_for_waveguide_based_ar_display_b200.couplers_coor import WaveguideDesign

KERNEL = GRTF.process_rays_kernel_pro_fullColor
STRICT = KERNEL.configured(strict=True)


def run_engine(kernel, scene, num_iter=1, EB=None, rng=None):
    EB = scene.new_matrix_EB() if EB is None else EB
    rng = scene.rays.rng_states.copy() if rng is None else rng
    for _ in range(num_iter):
        kernel[(scene.rays.num_rays + 255) // 256, 256](*scene.kernel_args(EB, rng))
    return EB, rng


def run_oracle(oracle, scene, num_iter=1):
    EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
    for _ in range(num_iter):
        oracle.trace(*scene.kernel_args(EB, rng))
    return EB, rng


def assert_same(a, b, what):
    EBa, ra = a; EBb, rb = b
    bad = np.flatnonzero(ra != rb)
    assert bad.size == 0, f"{what}: rng_states differ for {bad.size} rays, first {bad[:8]}"
    assert np.array_equal(EBa, EBb), f"{what}: matrix_EB differs in {np.count_nonzero(EBa != EBb)} bins"


# ---------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", ["walk_small", "walk_c1", "walk_mid", "walk_deep", "walk_fine", "walk_thin", "walk_mix", "walk_pol"])
@pytest.mark.parametrize("mode", ["fast", "strict"])
def test_golden_fixture(name, mode):
    scene, g = load_golden_walk(name)
    got = run_engine(KERNEL if mode == "fast" else STRICT, scene, int(g["num_iter"]))
    assert_same(got, (golden_bins(g), g["rng_states"]), f"{name}/{mode}")
    assert got[0].sum() > 0


# ---------------------------------------------------------------------------- oracle, seeded inputs
@pytest.mark.parametrize("cfg", [
    dict(nx=5, ny=5, rays=64, seed=101, lmd=[1]),                 # BASELINE config 1 shape
    dict(nx=7, ny=4, rays=300, seed=102, lmd=None),
    dict(nx=3, ny=2, rays=2, seed=103, lmd=None),                 # ragged: one TE + one TM ray per cell
    dict(nx=2, ny=2, rays=5000, seed=104, lmd=None),              # runner-sized cells
])
def test_against_oracle(cfg, oracle):
    scene = si.make_scene(cfg["nx"], cfg["ny"], cfg["rays"], seed=cfg["seed"], lmd_subset=cfg["lmd"])
    want = run_oracle(oracle, scene, 2)
    assert_same(run_engine(KERNEL, scene, 2), want, "fast")
    assert_same(run_engine(STRICT, scene, 2), want, "strict")


def test_deep_walks_against_oracle(oracle):
    """Divergence stress: thin plate, wide FoV, 15 fold slices, strong turn orders (BASELINE config 5)."""
    d = WaveguideDesign(t=0.3, num_FC=15, fov_x_deg=24.0)
    eff = dict(incouple=0.9, incouple_m1=0.08, ic_zero=0.9, ic_cross=0.05, fc_zero=0.6, fc_turn=0.3,
               oc_zero=0.85, oc_cross=0.06, outcouple=0.05)
    scene = si.make_scene(4, 3, 600, seed=7, design=d, eff=eff)
    want = run_oracle(oracle, scene)
    assert_same(run_engine(KERNEL, scene), want, "fast/deep")
    assert_same(run_engine(STRICT, scene), want, "strict/deep")
    assert want[0].sum() > 50


def test_many_slices_take_the_fallback_layouts(oracle):
    """The walk keeps a cell's Jones rows in shared memory when 16 warps of them fit (up to ~105 event rows) and the zone
    tables of designs with up to 254 / 128 zones.  Designs beyond that -- 25 fold slices + 12 out-coupler slices: 178
    rows, Jones rows in the global scratch; 60 + 20 slices: 138 zones, transition table in global memory; 160 + 40
    slices: 886 rows, 4 warps per SM, more than 254 zones: zone ids from global memory as well -- run
    the same code on the fallback layouts and must give the same bins and RNG states."""
    eff = dict(incouple=0.9, ic_zero=0.9, fc_zero=0.7, fc_turn=0.2, oc_zero=0.8, outcouple=0.1)
    for nfc, noc in ((25, 12), (60, 20), (160, 40)):
        scene = si.make_scene(4, 3, 500, seed=90 + nfc, design=WaveguideDesign(num_FC=nfc, num_OC=noc), eff=eff)
        assert len(scene.geom["FC_offset"]) - 1 == nfc and len(scene.geom["OC_offset"]) - 1 == noc
        want = run_oracle(oracle, scene, 2)
        assert_same(run_engine(KERNEL, scene, 2), want, f"fast/{nfc}+{noc} slices")
        assert want[0].sum() > 0


def test_fine_eyebox_grid(oracle):
    """BASELINE config 4: high-resolution bins."""
    scene = si.make_scene(3, 3, 800, eb=(320, 480), seed=12,
                          eff=dict(incouple=0.9, ic_zero=0.95, fc_zero=0.8, fc_turn=0.15, outcouple=0.1))
    assert_same(run_engine(KERNEL, scene), run_oracle(oracle, scene), "fast/fine-eyebox")


def test_shuffled_ray_order(oracle):
    """Rays need not arrive cell by cell: any order must give the same per-ray results."""
    scene = si.make_scene(4, 3, 40, seed=31)
    perm = np.random.default_rng(0).permutation(scene.rays.num_rays)
    rays = scene.rays
    scene.rays = si.RaySet(*(a[perm].copy() for a in rays.arrays()), rays.rng_states[perm].copy())
    want = run_oracle(oracle, scene)
    assert_same(run_engine(KERNEL, scene), want, "fast/shuffled")


def test_empty_and_invalid_inputs(oracle):
    scene = si.make_scene(3, 2, 20, seed=8)
    empty = scene.rays.take(slice(0, 0))
    full = scene.rays
    scene.rays = empty
    EB, rng = run_engine(KERNEL, scene)
    assert EB.sum() == 0 and rng.size == 0
    scene.rays = full
    # rays that point outside the LUT grid are skipped and leave their RNG state untouched
    bad = full.take(slice(0, full.num_rays))
    bad.m[:7] = 99.0
    bad.lmd_num[7:11] = -1.0
    scene.rays = bad
    for k in (KERNEL, STRICT):
        EB, rng = run_engine(k, scene)
        assert np.array_equal(rng[:11], bad.rng_states[:11])
    scene.rays = full
    with pytest.raises(TypeError):
        args = list(scene.kernel_args(scene.new_matrix_EB())); args[0] = args[0].astype(np.float64)
        KERNEL[1, 256](*args)


# ---------------------------------------------------------------------------- invariants
def test_tile_size_and_launch_split_invariance():
    scene = si.make_scene(5, 4, 500, seed=44)
    base = run_engine(KERNEL, scene)
    # (tiles of 2048 rays and more: the last tiles of a launch are handed out in quarters, wgrt_walk.cu TAIL_SPLIT)
    for tile in (32, 97, 1000, 2048, 2500, 4099, 100000):
        assert_same(run_engine(KERNEL.configured(tile_hint=tile), scene), base, f"tile={tile}")
    # two half launches == one launch (per-ray RNG, additive bins)
    N = scene.rays.num_rays
    EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
    full = scene.rays
    for sl in (slice(0, N // 3), slice(N // 3, N)):
        scene.rays = full.take(sl)
        part_rng = scene.rays.rng_states
        run_engine(KERNEL, scene, EB=EB, rng=part_rng)
        rng[sl] = part_rng
    scene.rays = full
    assert_same((EB, rng), base, "split launches")


def test_fast_equals_strict_at_scale():
    """Too large for the CPU oracle in seconds: pin the fast engine on the literal GPU walk."""
    scene = si.make_scene(20, 15, 2000, seed=77)          # 1.8 M rays
    a = run_engine(KERNEL, scene, 2)
    b = run_engine(STRICT, scene, 2)
    assert_same(a, b, "fast vs strict, 1.8M rays x 2")
    # conservation: one deposit at most per ray per launch
    assert 0 < a[0].sum() <= 2 * scene.rays.num_rays
    assert np.all(a[0] == np.round(a[0]))


def test_counters_match_oracle(oracle):
    scene = si.make_scene(4, 3, 100, seed=19)
    EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
    want = oracle.trace(*scene.kernel_args(EB, rng), counters=True)
    for kern, keys in ((KERNEL.configured(counters=True),
                        ("rays", "bounces", "draws", "draw2", "draw3", "efield", "iters", "deposits")),
                       (STRICT.configured(counters=True),
                        ("rays", "bounces", "draws", "draw2", "draw3", "efield", "iters", "deposits",
                         "poly_tests", "edge_visits", "straddle", "cross"))):
        _capi.reset_counters()
        run_engine(kern, scene)
        got = _capi.read_counters()
        for k in keys:
            assert got[k] == want[k], (k, got[k], want[k])


# ---------------------------------------------------------------------------- buffers / boundary
def test_device_buffers_zero_copy_and_stream(oracle):
    import torch
    scene = si.make_scene(4, 3, 64, seed=3)
    want = run_oracle(oracle, scene)
    dev = []
    for a in scene.kernel_args(scene.new_matrix_EB()):
        if isinstance(a, np.ndarray):
            v = a.view(np.float64) if a.dtype == np.complex128 else a
            t = torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).cuda()
            dev.append(GRTF._TorchAlias(t, a.shape, a.dtype))
        else:
            dev.append(a)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        KERNEL[1, 256, s](*dev)
    s.synchronize()
    rng = dev[12]._t.cpu().numpy().view(np.uint32)
    EB = dev[32]._t.cpu().numpy()
    assert_same((EB, rng), want, "device buffers")


def test_host_entry_point(oracle):
    """wgrt_trace_fullcolor_host: the RUN:145-185 region for host buffers, num_iter launches."""
    scene = si.make_scene(4, 3, 64, seed=3)
    want = run_oracle(oracle, scene, 3)
    EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
    prob, keep = GRTF.pack_problem(scene.kernel_args(EB, rng), host=True)
    lib = _capi.load_library()
    tms = (C.c_float * 3)()
    _capi.check(lib.wgrt_trace_fullcolor_host(C.byref(prob), 3, tms), lib)
    assert_same((EB, rng), want, "host entry")
    assert all(t >= 0 for t in tms)


# ---------------------------------------------------------------------------- runner layout (row f2)
def test_runner_layout_equals_materialised_arrays(oracle):
    """Implicit rays (start points + layout rule) must give exactly what the runner's arrays give."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import runner
    scene = si.make_scene(5, 4, 300, seed=61)
    pts = si.points_in_disc(scene.geom["IC"], 150, 62)
    scene.rays = si.build_ray_set(pts, 5, 4, 3, 300)
    want = run_oracle(oracle, scene, 2)
    assert_same(run_engine(KERNEL, scene, 2), want, "materialised")
    # device launch on the implicit layout, fast and strict
    N = scene.rays.num_rays
    px = pts[:, 0].astype(np.float32); py = pts[:, 1].astype(np.float32)
    for kern in (KERNEL, STRICT):
        EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
        a = list(scene.kernel_args(EB, rng))
        a[0], a[1] = px, py
        for i in range(2, 12):
            a[i] = None
        for _ in range(2):
            kern.runner_layout(150, N)[1, 256](*a)
        assert_same((EB, rng), want, "runner layout / device")
    # host entry: seeds generated on the device, bins cleared on the device
    EB = runner.trace_full_color(pts, scene.geom, scene.n_g, scene.luts, 300, num_iter=2)
    assert np.array_equal(EB, want[0])
    # a cell sub-range with explicit RNG states (multi-GPU sharding)
    c0, c1 = 7, 31
    rng = si.initial_rng_states((c1 - c0) * 300, offset=c0 * 300)
    EBp = runner.trace_full_color(pts, scene.geom, scene.n_g, scene.luts, 300, num_iter=2, first_cell=c0,
                                  num_cells=c1 - c0, rng_states=rng)
    assert np.array_equal(rng, want[1][c0 * 300:c1 * 300])
    full = scene.rays
    scene.rays = full.take(slice(c0 * 300, c1 * 300))
    part = run_oracle(oracle, scene, 2)
    scene.rays = full
    assert np.array_equal(EBp, part[0])


@pytest.mark.parametrize("chunks", [1, 3, 7, 64])
def test_host_pipeline_chunking_is_invisible(chunks, oracle, monkeypatch):
    """The host entry pipelines H2D / walk / D2H over chunks (FoV-x column ranges with the runner
    layout, ray ranges with explicit arrays): any chunking, bins that do not start at zero, cell
    sub-ranges that cut through columns, and returned RNG states must all match the plain launches."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import runner
    monkeypatch.setenv("WGRT_HOST_CHUNKS", str(chunks))
    rpc = 120
    scene = si.make_scene(9, 4, rpc, seed=41)
    pts = si.points_in_disc(scene.geom["IC"], rpc // 2, 41 + 1)
    scene.rays = si.build_ray_set(pts, 9, 4, 3, rpc)
    want = run_oracle(oracle, scene, 2)
    # (a) runner layout, whole job, device-seeded RNG, bins cleared on the device
    EB = runner.trace_full_color(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=2)
    assert np.array_equal(EB, want[0])
    # (b) runner layout, bins accumulate onto what the caller passes in, RNG states round trip
    base = np.random.default_rng(5).integers(0, 3, size=scene.eb_shape).astype(np.float32)
    EB = base.copy(); rng = si.initial_rng_states(scene.rays.num_rays)
    runner.trace_full_color(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=2, matrix_EB=EB, rng_states=rng)
    assert np.array_equal(EB, base + want[0]) and np.array_equal(rng, want[1])
    # (c) a cell range that starts and ends inside FoV-x columns (multi-GPU shard); the rest of the
    #     caller's bins must come back untouched
    c0, c1 = 17, 83
    full = scene.rays
    scene.rays = full.take(slice(c0 * rpc, c1 * rpc))
    part = run_oracle(oracle, scene, 2)
    scene.rays = full
    EB = base.copy(); rng = si.initial_rng_states((c1 - c0) * rpc, offset=c0 * rpc)
    runner.trace_full_color(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=2, first_cell=c0, num_cells=c1 - c0,
                            matrix_EB=EB, rng_states=rng)
    assert np.array_equal(EB, base + part[0]) and np.array_equal(rng, part[1])
    EB = runner.trace_full_color(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=2, first_cell=c0,
                                 num_cells=c1 - c0)
    assert np.array_equal(EB, part[0])
    # (d) explicit ray arrays through the C ABI host entry
    EB = base.copy(); rng = scene.rays.rng_states.copy()
    prob, keep = GRTF.pack_problem(scene.kernel_args(EB, rng), host=True)
    lib = _capi.load_library()
    _capi.check(lib.wgrt_trace_fullcolor_host(C.byref(prob), 2, None), lib)
    assert np.array_equal(EB, base + want[0]) and np.array_equal(rng, want[1])
    # (e) zero RNG states reseed from the ray's index in the WHOLE job (GRTF:28-29), whatever chunk or
    #     shard the ray ends up in
    zeros = np.arange(7, scene.rays.num_rays, 131)
    rng0 = scene.rays.rng_states.copy(); rng0[zeros] = 0
    EB_o = scene.new_matrix_EB(); rng_o = rng0.copy()
    oracle.trace(*scene.kernel_args(EB_o, rng_o))
    EB = scene.new_matrix_EB(); rng = rng0.copy()
    prob, keep = GRTF.pack_problem(scene.kernel_args(EB, rng), host=True)
    _capi.check(lib.wgrt_trace_fullcolor_host(C.byref(prob), 1, None), lib)
    assert np.array_equal(rng, rng_o) and np.array_equal(EB, EB_o)
    lo, hi = 17 * rpc, 83 * rpc                      # a shard launched on its own
    scene.rays = full.take(slice(lo, hi))
    EB = scene.new_matrix_EB(); rng = rng0[lo:hi].copy()
    KERNEL.configured(ray_index_base=lo)[1, 256](*scene.kernel_args(EB, rng))
    scene.rays = full
    assert np.array_equal(rng, rng_o[lo:hi])


def test_host_entry_device_bins_and_seed_offset(oracle):
    """WGRT_FLAG_BINS_DEVICE: the host entry accumulates into the caller's device tensor and downloads
    nothing; rng_seed_offset shifts the device-side seeding rule (replicated jobs, independent streams)."""
    import torch
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import runner
    rpc = 200
    scene = si.make_scene(5, 4, rpc, seed=43)
    pts = si.points_in_disc(scene.geom["IC"], rpc // 2, 44)
    n = 5 * 4 * 3 * rpc
    for off in (0, 3 * n):
        rng = si.initial_rng_states(n, offset=off)
        want = runner.trace_full_color(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=2, rng_states=rng)
        dev = torch.full(scene.eb_shape, 2.0, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        alias = GRTF._TorchAlias(dev, scene.eb_shape, np.float32)
        out = runner.trace_full_color(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=2, matrix_EB=alias,
                                      rng_seed_offset=off)
        assert out is alias
        assert np.array_equal(dev.cpu().numpy(), want + 2.0)        # accumulated in place, device-seeded streams
        out = runner.trace_full_color(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=2, matrix_EB=alias,
                                      rng_seed_offset=off, bins_start_zero=True)
        assert np.array_equal(dev.cpu().numpy(), want)               # cleared on the device first
    assert want.sum() > 0


# ---------------------------------------------------------------------------- single-wavelength twin (row f3)
@pytest.mark.parametrize("mode", ["fast", "strict"])
def test_single_lambda_twin(mode):
    from test_oracle_golden import _single_lambda_case
    g, args, EB, rng, want = _single_lambda_case()
    kern = GRTF.process_rays_kernel_pro if mode == "fast" else GRTF.process_rays_kernel_pro.configured(strict=True)
    assert kern.threshold == 1e-15 and kern.single_lambda
    kern[(len(rng) + 255) // 256, 256](*args)
    assert_same((EB, rng), (want, g["rng_states"]), f"single lambda / {mode}")
    with pytest.raises(TypeError):
        kern[1, 256](*args, None)        # 33 arguments: that is the full-colour signature


# ---------------------------------------------------------------------------- unit-level parity
def test_xorshift_unit(oracle):
    g = np.load(os.path.join(GOLDEN, "units.npz"))
    lib = _capi.load_library()
    st = g["xs_in"].copy(); last = np.zeros(len(st))
    _capi.check(lib.wgrt_debug_xorshift(st.ctypes.data, len(st), 7, last.ctypes.data), lib)
    assert np.array_equal(st, g["xs_out"]) and np.array_equal(last, g["xs_last"])


def test_efield_unit():
    """libdevice vs the reference's CPU libm: same algebra, last-bit differences allowed.
    Tolerance 1e-12 absolute on amplitudes, 1e-9 on the wrapped phase away from the +-pi seam."""
    g = np.load(os.path.join(GOLDEN, "units.npz"))
    lib = _capi.load_library()
    ete, etm, delta = (np.ascontiguousarray(g[k]) for k in ("ef_ete", "ef_etm", "ef_delta"))
    n = len(ete)
    out = np.zeros((n, 3))
    jones = np.ascontiguousarray(g["ef_jones"])
    _capi.check(lib.wgrt_debug_efield(ete.ctypes.data, etm.ctypes.data, delta.ctypes.data,
                                      jones.ctypes.data, n, out.ctypes.data), lib)
    ref = g["ef_out"]
    np.testing.assert_allclose(out[:, :2], ref[:, :2], rtol=0, atol=1e-12)
    d = np.abs(out[:, 2] - ref[:, 2])
    d = np.minimum(d, 2 * np.pi - d)
    assert np.all(d < 1e-9)


@pytest.mark.parametrize("ring", ["IC", "FC", "OC", "eff_reg1", "eff_reg2"])
@pytest.mark.parametrize("mode", [0, 1])
def test_polygon_unit_golden(ring, mode):
    g = np.load(os.path.join(GOLDEN, "units.npz"))
    lib = _capi.load_library()
    verts = np.ascontiguousarray(g[ring + "_verts"]); off = np.ascontiguousarray(g[ring + "_off"])
    pts = g[ring + "_pts"]
    px = np.ascontiguousarray(pts[:, 0]); py = np.ascontiguousarray(pts[:, 1])
    out = np.zeros(len(px), dtype=np.int32)
    _capi.check(lib.wgrt_debug_locate(verts.ctypes.data, len(verts), off.ctypes.data, len(off) - 1,
                                      px.ctypes.data, py.ctypes.data, len(px), out.ctypes.data, mode), lib)
    assert np.array_equal(out, g[ring + "_hit"])
