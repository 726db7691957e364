"""Row f1: the evaluation mirror against outputs of the reference's own evaluation() (tests/golden/
eval.npz, produced by oracle/make_golden.py with only `colour` stubbed).  Tolerance: rel 1e-5 on the
maps (the reference sums float32 in NumPy's pairwise order; the GPU kernel in its own order)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import AR_system_evaluation_functions as EV


def load():
    g = np.load(os.path.join(GOLDEN, "eval.npz"))
    EB = np.zeros(int(np.prod(g["eb_shape"])), dtype=np.float32)
    EB[g["eb_index"]] = g["eb_value"]
    EB = EB.reshape(tuple(g["eb_shape"]))
    return g, EB, EB / int(g["rays_per_fov"]) / int(g["num_iter"])


def test_evaluation_post_processing_matches_reference():
    """CPU part (reference lines 112-160) fed with the reference's own pupil sums."""
    g, EB, EB2 = load()
    delta_e, U_fov, U_EB, img = EV.evaluation(EB2, matrix_eye_perceive=g["perceive"])
    assert U_fov == pytest.approx(float(g["U_fov"]), rel=1e-9)
    assert U_EB == pytest.approx(float(g["U_EB"]), rel=1e-9)
    np.testing.assert_allclose(img, g["output_image"], rtol=1e-5, atol=1e-6)
    assert delta_e == pytest.approx(float(g["delta_e_stubbed"]), rel=1e-9)   # same restated colour maths
    assert 0 < U_fov < 1 and 0 < U_EB < 1


def test_ciede2000_known_pairs():
    """Sharma et al. (2005) test data, pairs 1, 2, 17 and 25."""
    lab1 = np.array([[50.0, 2.6772, -79.7751], [50.0, 3.1571, -77.2803], [50.0, 2.5, 0.0], [60.2574, -34.0099, 36.2677]])
    lab2 = np.array([[50.0, 0.0, -82.7485], [50.0, 0.0, -82.7485], [73.0, 25.0, -18.0], [60.4626, -34.1751, 39.4387]])
    want = np.array([2.0425, 2.8615, 27.1492, 1.2644])
    np.testing.assert_allclose(EV._delta_e_2000(lab1, lab2), want, atol=1e-4)


@pytest.mark.gpu
def test_pupil_sums_kernel_matches_reference():
    g, EB, EB2 = load()
    perceive, cells = EV.pupil_sums(EB2)
    assert perceive.shape == g["perceive"].shape
    np.testing.assert_allclose(perceive, g["perceive"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(cells, EB2.sum(axis=(-1, -2), dtype=np.float64), rtol=1e-5)
    # on raw integer counts the sums are exact
    p_int, c_int = EV.pupil_sums(EB)
    assert np.array_equal(c_int, EB.sum(axis=(-1, -2), dtype=np.float64).astype(np.float32))
    assert np.all(p_int == np.round(p_int))


@pytest.mark.gpu
def test_evaluation_end_to_end_on_gpu():
    g, EB, EB2 = load()
    delta_e, U_fov, U_EB, img = EV.evaluation(EB2)
    assert U_fov == pytest.approx(float(g["U_fov"]), rel=1e-5)
    assert U_EB == pytest.approx(float(g["U_EB"]), rel=1e-5)
    np.testing.assert_allclose(img, g["output_image"], rtol=1e-4, atol=1e-5)
    eff = EV.efficiency_per_colour(EB, 6 * 5 * 3 * int(g["rays_per_fov"]), int(g["num_iter"]))
    want = EB.sum(axis=(1, 2, 3, 4), dtype=np.float64) / (6 * 5 * 3 * int(g["rays_per_fov"])) / int(g["num_iter"]) * 3
    np.testing.assert_allclose(eff, want, rtol=1e-12)


@pytest.mark.gpu
def test_pupil_sums_other_shapes():
    rs = np.random.default_rng(3)
    for shape, mask, sy, sx in (((1, 2, 3, 40, 50), 30, 8, 12), ((2, 1, 1, 64, 64), 16, 1, 1), ((1, 1, 2, 20, 20), 30, 8, 12),
                                ((1, 2, 2, 80, 120), 30, 1, 1),        # the full pupil convolution (EVAL:75-89)
                                ((1, 1, 2, 320, 480), 120, 1, 1),      # BASELINE config 4: 614 KB tiles, every pupil position
                                ((2, 1, 1, 320, 480), 120, 32, 48),    # ... sampled as the reference samples
                                ((1, 1, 1, 97, 131), 31, 3, 5)):       # ragged sizes, odd mask
        EB = rs.integers(0, 5, size=shape).astype(np.float32)
        lib = EV._capi.load_library()
        n_epy = (shape[3] - mask) // sy + 1 if shape[3] >= mask else 0
        n_epx = (shape[4] - mask) // sx + 1 if shape[4] >= mask else 0
        out = np.zeros(shape[:3] + (n_epy, n_epx), np.float32); cells = np.zeros(shape[:3], np.float32)
        EV._capi.check(lib.wgrt_eval_pupil_sums_host(EB.ctypes.data, *shape, mask, sy, sx, out.ctypes.data,
                                                     cells.ctypes.data), lib)
        yy, xx = np.ogrid[:mask, :mask]
        disc = (np.sqrt((xx - (mask / 2 - 0.5)) ** 2 + (yy - (mask / 2 - 0.5)) ** 2) <= mask / 2)
        if n_epy * n_epx > 4000:
            # dense sampling: check against FFT-free direct correlation on a subset of positions
            pos = [(int(a), int(b)) for a, b in zip(rs.integers(0, n_epy, 300), rs.integers(0, n_epx, 300))]
            pos += [(0, 0), (n_epy - 1, n_epx - 1), (0, n_epx - 1), (n_epy - 1, 0)]
        else:
            pos = [(iy, ix) for iy in range(n_epy) for ix in range(n_epx)]
        for iy, ix in pos:
            want = (EB[..., iy * sy:iy * sy + mask, ix * sx:ix * sx + mask] * disc).sum(axis=(-1, -2))
            assert np.array_equal(out[..., iy, ix], want), (shape, iy, ix)
        assert np.array_equal(cells, EB.sum(axis=(-1, -2)))


@pytest.mark.gpu
def test_trace_and_evaluate_keeps_bins_on_the_device():
    """runner.trace_and_evaluate == runner.trace_full_color followed by the evaluation mirror, with the
    pupil sums / cell totals bit-equal (same kernel, same bins) and the maps within 1e-5."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import runner, synthetic_inputs as si
    rpc, it = 400, 2
    scene = si.make_scene(6, 5, 2, seed=21, eff=dict(incouple=0.9, ic_zero=0.95, fc_zero=0.8, fc_turn=0.15, outcouple=0.2))
    pts = si.points_in_disc(scene.geom["IC"], rpc // 2, 22)
    EB = runner.trace_full_color(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=it)
    assert EB.sum() > 1000
    want_p, want_c = EV.pupil_sums(EB)
    got = runner.trace_and_evaluate(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=it, return_perceive=True, return_image=True)
    assert np.array_equal(got["cell_sums"], want_c)
    assert np.array_equal(got["matrix_eye_perceive"] * np.float32(rpc) * np.float32(it), want_p) or \
        np.allclose(got["matrix_eye_perceive"], want_p / rpc / it, rtol=1e-6, atol=0)
    d_e, U_fov, U_EB, img = EV.evaluation(EB / np.float32(rpc) / np.float32(it))
    assert got["U_fov"] == pytest.approx(U_fov, rel=1e-5) and got["U_EB"] == pytest.approx(U_EB, rel=1e-5)
    assert got["delta_e"] == pytest.approx(d_e, rel=1e-5)
    np.testing.assert_allclose(got["output_image"], img, rtol=1e-4, atol=1e-5)   # device: double -> float32; host: cv2 float32 HSV
    eff = EV.efficiency_per_colour(EB, scene.eb_shape[0] * 6 * 5 * rpc, it)
    np.testing.assert_allclose(got["efficiency"], eff, rtol=1e-12)


@pytest.mark.gpu
def test_runner_main_prints_the_reference_report(capsys):
    """The runner script equivalent (RUN:11-210) end to end at a small size."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import runner
    assert runner.main(["--fov", "6", "5", "--rays-per-fov", "400", "--num-iter", "2"]) == 0
    out = capsys.readouterr().out
    for key in ("Number of rays traced : 72,000", "Efficiency (Red)", "Color dispersion", "FoV uniformity", "Eyebox uniformity"):
        assert key in out, out


@pytest.mark.gpu
def test_bins_pack_unpack_u8():
    """wgrt_bins_pack_u8 / _unpack_u8 (the exact narrow all-reduce of multi_gpu.reduce_bins)."""
    import ctypes as C
    import torch
    lib = EV._capi.load_library()
    rs = np.random.default_rng(8)
    for make, want_bad in ((lambda: rs.integers(0, 256, 4096 * 7).astype(np.float32), 0),
                           (lambda: rs.integers(0, 9, 1024).astype(np.float32) + np.float32(0.5) * (np.arange(1024) == 77), 1),
                           (lambda: np.where(np.arange(2048) == 5, 256.0, 1.0).astype(np.float32), 1),
                           (lambda: np.where(np.arange(2048) == 9, -1.0, 2.0).astype(np.float32), 1)):
        a = make()
        t = torch.from_numpy(a).cuda()
        q = torch.empty(a.size, dtype=torch.uint8, device="cuda"); st = torch.empty(2, dtype=torch.int32, device="cuda")
        EV._capi.check(lib.wgrt_bins_pack_u8(C.c_void_p(t.data_ptr()), a.size, C.c_void_p(q.data_ptr()),
                                             C.c_void_p(st.data_ptr()), C.c_float(255.0), None), lib)
        torch.cuda.synchronize()
        vmax = st[0:1].view(torch.float32).item()
        assert int(st[1].item()) == want_bad
        if not want_bad:
            assert vmax == a.max() and np.array_equal(q.cpu().numpy(), a.astype(np.uint8))
            back = torch.zeros_like(t)
            EV._capi.check(lib.wgrt_bins_unpack_u8(C.c_void_p(q.data_ptr()), a.size, C.c_void_p(back.data_ptr()), None), lib)
            assert torch.equal(back, t)
            EV._capi.check(lib.wgrt_bins_pack_u8(C.c_void_p(t.data_ptr()), a.size, C.c_void_p(q.data_ptr()),
                                                 C.c_void_p(st.data_ptr()), C.c_float(31.0), None), lib)
            assert int(st[1].item()) == int(a.max() > 31)      # entries above the limit are flagged


@pytest.mark.gpu
def test_device_evaluation_matches_reference():
    """evaluation() lines 110-160 on the device (wgrt_eval_metrics): U_fov, U_EB and the sRGB view against the
    reference's own outputs (rel 1e-5), delta_e against the restated colour maths the fixture was made with."""
    g, EB, EB2 = load()
    raw, _ = EV.pupil_sums(EB)                                   # raw integer pupil sums, exact
    scale = 1.0 / (int(g["rays_per_fov"]) * int(g["num_iter"]))
    delta_e, U_fov, U_EB, img = EV.evaluation_device(raw, scale)
    assert U_fov == pytest.approx(float(g["U_fov"]), rel=1e-5)
    assert U_EB == pytest.approx(float(g["U_EB"]), rel=1e-5)
    assert delta_e == pytest.approx(float(g["delta_e_stubbed"]), rel=1e-6)
    np.testing.assert_allclose(img, g["output_image"], rtol=1e-4, atol=1e-5)
    # and against the host mirror on the same raw sums (same formulas, double precision both sides)
    d2, uf2, ue2, img2 = EV.evaluation(EB2, matrix_eye_perceive=raw * np.float32(scale))
    assert (delta_e, U_fov, U_EB) == pytest.approx((d2, uf2, ue2), rel=1e-6)
    # eye positions that see a dark pixel count as zero uniformity on both sides
    raw0 = raw.copy(); raw0[:, 2, 3, 1, 4] = 0
    d3, uf3, ue3, _ = EV.evaluation_device(raw0, scale, return_image=False)
    d4, uf4, ue4, _ = EV.evaluation(EB2, matrix_eye_perceive=raw0 * np.float32(scale))
    assert (d3, uf3, ue3) == pytest.approx((d4, uf4, ue4), rel=1e-6) and ue3 == 0


@pytest.mark.gpu
def test_trace_and_evaluate_finishes_on_the_device():
    """runner.trace_and_evaluate: K launches + pupil sums + evaluation() on the device, < 100 KB downloaded;
    equal to the host-finished path on the same job."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import runner, synthetic_inputs as si
    eff = dict(incouple=0.9, incouple_m1=0.05, ic_zero=0.95, ic_cross=0.02, fc_zero=0.8, fc_turn=0.18,
               oc_zero=0.85, oc_cross=0.02, outcouple=0.12)
    rpc = 4000
    scene = si.make_scene(6, 5, rpc, seed=41, eff=eff, build_rays=False)
    pts = si.points_in_disc(scene.geom["IC"], rpc // 2, 42)
    dev = runner.trace_and_evaluate(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=2, return_image=True,
                                    return_perceive=True)
    host = runner.trace_and_evaluate(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=2, host_evaluation=True)
    assert np.array_equal(dev["cell_sums"], host["cell_sums"]) and dev["cell_sums"].sum() > 0
    assert np.array_equal(dev["matrix_eye_perceive"], host["matrix_eye_perceive"])
    for k in ("delta_e", "U_fov", "U_EB"):
        assert dev[k] == pytest.approx(host[k], rel=1e-6), k
    np.testing.assert_allclose(dev["output_image"], host["output_image"], rtol=1e-4, atol=1e-5)
    lean = runner.trace_and_evaluate(pts, scene.geom, scene.n_g, scene.luts, rpc, num_iter=2)
    assert "output_image" not in lean and "matrix_eye_perceive" not in lean
    assert (lean["delta_e"], lean["U_fov"], lean["U_EB"]) == (dev["delta_e"], dev["U_fov"], dev["U_EB"])
