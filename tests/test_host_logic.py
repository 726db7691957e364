"""CPU-only tests of the host side: generators, argument validation, the C-ABI library surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi, synthetic_inputs as si
from gpu_ray_tracing_for_waveguide_based_ar_display_b200.couplers_coor import (
    WaveguideDesign, couplers_coor_full_color)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_geometry_sanity_targets():
    """values obtained from the reference's couplers_coor_full_color."""
    out = couplers_coor_full_color(100, 75)
    assert len(out) == 37
    IC, FC, FC_off, OC, OC_off, r1, r2, rect, rng_, TIR, gap = out[:11]
    assert IC.shape == (100, 2) and np.array_equal(IC[0], IC[-1])
    assert list(FC_off) == [0, 5, 11, 17, 22, 28, 72, 135] and FC.shape == (135, 2)
    assert list(OC_off) == [0, 4, 9, 15, 21, 26, 30] and OC.shape == (30, 2)
    assert r1.shape == (83, 2) and r2.shape == (111, 2)
    assert rect.shape == (100, 75, 4, 2) and rng_.shape == (100, 75, 4)
    assert TIR.shape == (3, 100, 75, 4) and gap.shape == (3, 100, 75, 8)
    assert abs(out[14] - 315.108) < 1e-3 and abs(out[15] - np.pi / 2) < 1e-12
    np.testing.assert_allclose(r1.min(0), [-30.00, -6.14], atol=6e-3)
    np.testing.assert_allclose(r1.max(0), [9.17, 21.37], atol=6e-3)
    np.testing.assert_allclose(OC.min(0), [-9.17, 8.63], atol=6e-3)
    assert not np.isnan(TIR).any() and not np.isnan(gap).any()
    for k in range(7):          # slices are closed rings
        assert np.allclose(FC[FC_off[k]], FC[FC_off[k + 1] - 1])
    # eyebox rectangle corner order and range vector (couplers_coor.py:514-532)
    xmin, xmax, ymin, ymax = rng_[3, 4]
    assert np.allclose(rect[3, 4], [[xmin, ymax], [xmin, ymin], [xmax, ymin], [xmax, ymax]])


def test_geometry_is_parameterised():
    d = WaveguideDesign(t=0.3, num_FC=15, fov_x_deg=24.0)
    out = couplers_coor_full_color(6, 5, design=d)
    assert len(out[2]) - 1 == 15
    assert out[10].shape == (3, 6, 5, 8)


def test_generate_points_in_polygon():
    IC = couplers_coor_full_color(3, 3)[0]
    pts = GRTF.generate_points_in_polygon(IC, 777, rng=np.random.default_rng(4))
    assert pts.shape == (777, 2)
    assert np.all(np.hypot(pts[:, 0] + 28, pts[:, 1] - 15) <= 2.0 + 1e-9)
    again = GRTF.generate_points_in_polygon(IC, 777, rng=np.random.default_rng(4))
    assert np.array_equal(pts, again)
    np.random.seed(9)
    a = GRTF.generate_points_in_polygon(IC, 50)
    np.random.seed(9)
    assert np.array_equal(a, GRTF.generate_points_in_polygon(IC, 50))
    assert GRTF.generate_points_in_polygon(IC, 0).shape == (0, 2)


def test_ray_set_layout_follows_runner():
    pts = np.arange(10, dtype=np.float64).reshape(5, 2)
    rs = si.build_ray_set(pts, 3, 2, 3, 10)
    assert rs.num_rays == 3 * 2 * 3 * 10
    # runner order: FoV-x outer, FoV-y, wavelength inner; TE half then TM half (RUN:82-115)
    blk = lambda k: slice(10 * k, 10 * k + 10)
    assert np.all(rs.m[blk(0)] == 0) and np.all(rs.n[blk(0)] == 0) and np.all(rs.lmd_num[blk(0)] == 0)
    assert np.all(rs.lmd_num[blk(1)] == 1) and np.all(rs.n[blk(3)] == 1) and np.all(rs.m[blk(6)] == 1)
    assert np.array_equal(rs.te[blk(4)], [1] * 5 + [0] * 5) and np.array_equal(rs.tm[blk(4)], [0] * 5 + [1] * 5)
    assert np.array_equal(rs.x[blk(7)], np.tile(pts[:, 0], 2).astype(np.float32))
    assert rs.rng_states.dtype == np.uint32
    assert rs.rng_states[0] == 0x9E3779B9 and rs.rng_states[1] == (2 * 0x9E3779B9) % 2 ** 32
    sub = si.build_ray_set(pts, 3, 2, 3, 10, lmd_subset=[1])
    assert sub.num_rays == 60 and np.all(sub.lmd_num == 1)


def test_lut_shapes(small_scene):
    l = small_scene.luts
    assert l["lut_ic1"].shape == (3, 4, 3, 42) and l["lut_fc1"].shape == (7, 3, 4, 3, 26)
    assert l["lut_oc2"].shape == (6, 3, 4, 3, 42)
    assert all(v.dtype == np.complex128 for v in l.values())


def test_pack_problem_accepts_reference_arguments(small_scene):
    EB = small_scene.new_matrix_EB()
    prob, keep = GRTF.pack_problem(small_scene.kernel_args(EB), host=True)
    assert prob.num_rays == small_scene.rays.num_rays
    assert (prob.L, prob.X, prob.Y, prob.EBy, prob.EBx) == (3, 4, 3, 80, 120)
    assert (prob.n_FC, prob.n_OC, prob.C_ic, prob.C_fc, prob.C_oc) == (7, 6, 42, 26, 42)
    assert prob.matrix_EB == EB.ctypes.data


def _args(scene):
    return list(scene.kernel_args(scene.new_matrix_EB()))


def test_pack_problem_rejects_bad_arguments(small_scene):
    a = _args(small_scene)
    with pytest.raises(TypeError):
        GRTF.pack_problem(a[:-1], host=True)                      # 32 arguments
    b = list(a); b[0] = b[0].astype(np.float64)
    with pytest.raises(TypeError):
        GRTF.pack_problem(b, host=True)                            # x_v dtype
    b = list(a); b[1] = b[1][:-1].copy()
    with pytest.raises(ValueError):
        GRTF.pack_problem(b, host=True)                            # ragged ray arrays
    b = list(a); b[12] = b[12].astype(np.int64)
    with pytest.raises(TypeError):
        GRTF.pack_problem(b, host=True)                            # rng dtype
    b = list(a); b[23] = b[23].astype(np.complex64)
    with pytest.raises(TypeError):
        GRTF.pack_problem(b, host=True)                            # complex64 LUT
    b = list(a); b[26] = b[26][:-1].copy()
    with pytest.raises(ValueError):
        GRTF.pack_problem(b, host=True)                            # nFC mismatch between offsets and LUT
    b = list(a); b[32] = np.zeros((3, 4, 3, 80, 120), np.float32)
    with pytest.raises(ValueError):
        GRTF.pack_problem(b, host=True)                            # EB must be [L, Y, X, ...]
    b = list(a); b[15] = np.array([0, 9, 5, 135, 135, 135, 135, 135], dtype=np.int64)
    with pytest.raises(ValueError):
        GRTF.pack_problem(b, host=True)                            # non-monotone offsets
    b = list(a); b[13] = np.asfortranarray(np.tile(b[13], (1, 1)))[:, ::-1]
    with pytest.raises(ValueError):
        GRTF.pack_problem(b, host=True)                            # non-contiguous
    b = list(a); b[2] = None; b[3] = None; b[4] = None; b[5] = None
    GRTF.pack_problem(b, host=True)                                # dead arrays may be omitted
    with pytest.raises(TypeError):
        GRTF.pack_problem(a, host=False)                           # host arrays where device buffers are required


def test_pack_problem_runner_layout(small_scene):
    a = _args(small_scene)
    pts = np.zeros(24, np.float32)
    b = [pts, pts] + [None] * 10 + [None] + a[13:]
    prob, _ = GRTF.pack_problem(b, host=True, runner_points=24, num_rays=4 * 3 * 3 * 48)
    assert prob.runner_points == 24 and prob.num_rays == 1728 and not prob.rng_states
    with pytest.raises(ValueError):
        GRTF.pack_problem(b, host=True, runner_points=24, num_rays=1729)
    with pytest.raises(ValueError):
        GRTF.pack_problem(b, host=True, runner_points=24, num_rays=1728, runner_first_cell=1)
    with pytest.raises(ValueError):
        GRTF.pack_problem([pts[:5], pts] + b[2:], host=True, runner_points=24, num_rays=1728)


def test_pack_problem_shard_and_evaluate_fields(small_scene):
    """ray_index_base / rng_seed_offset reach the struct; the evaluate entry may omit matrix_EB (eb=...);
    a device matrix_EB is accepted by the host entry only with WGRT_FLAG_BINS_DEVICE."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi
    a = _args(small_scene)
    prob, _ = GRTF.pack_problem(a, host=True, ray_index_base=12345, rng_seed_offset=777)
    assert prob.ray_index_base == 12345 and prob.rng_seed_offset == 777
    prob, _ = GRTF.pack_problem(a, host=True)
    assert prob.ray_index_base == 0 and prob.rng_seed_offset == 0
    pts = np.zeros(24, np.float32)
    b = [pts, pts] + [None] * 10 + [None] + a[13:32] + [None]
    prob, _ = GRTF.pack_problem(b, host=True, runner_points=24, num_rays=1728, eb=(40, 60))
    assert not prob.matrix_EB and (prob.EBy, prob.EBx) == (40, 60)
    with pytest.raises((TypeError, ValueError)):
        GRTF.pack_problem(b, host=True, runner_points=24, num_rays=1728)          # no bins and no eb
    with pytest.raises(ValueError):
        GRTF.pack_problem(b, host=True, runner_points=24, num_rays=1728, eb=(0, 60))

    class FakeDevice:                                                             # quacks like a device tensor
        __cuda_array_interface__ = {"data": (0xdead0000, False), "shape": (3, 3, 4, 80, 120), "typestr": "<f4",
                                    "strides": None, "version": 3}
    c = list(a); c[32] = FakeDevice()
    with pytest.raises(TypeError):
        GRTF.pack_problem(c, host=True)                                           # device bins without the flag
    prob, _ = GRTF.pack_problem(c, host=True, flags=_capi.WGRT_FLAG_BINS_DEVICE)
    assert prob.matrix_EB == 0xdead0000 and prob.flags & _capi.WGRT_FLAG_BINS_DEVICE
    k = GRTF.process_rays_kernel_pro_fullColor.configured(ray_index_base=99)
    assert k.ray_index_base == 99 and k.runner_layout(4, 16).ray_index_base == 99
    assert GRTF.process_rays_kernel_pro_fullColor.ray_index_base == 0


def test_kernel_object_surface():
    k = GRTF.process_rays_kernel_pro_fullColor
    launcher = k[1024, 256]
    assert callable(launcher)
    with pytest.raises(ValueError):
        k[1]
    strict = k.configured(strict=True, counters=True)
    assert strict.flags == (_capi.WGRT_FLAG_STRICT | _capi.WGRT_FLAG_COUNTERS) and k.flags == 0
    with pytest.raises(TypeError):
        launcher(1, 2, 3)


def test_capi_library_exports_every_declared_symbol():
    """libwgrt.so loads without a GPU and exports exactly what include/wgrt.h declares."""
    hdr = open(os.path.join(ROOT, "include", "wgrt.h")).read()
    declared = set(re.findall(r"\b(wgrt_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"wgrt_problem_t"}
    assert {"wgrt_trace_fullcolor", "wgrt_trace_fullcolor_host", "wgrt_version"} <= declared
    lib = _capi.load_library()
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in wgrt.h but not exported"
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    assert lib.wgrt_version() == 100
    assert ctypes.sizeof(_capi.WgrtProblem) == lib.wgrt_problem_size()      # layout guard
    assert isinstance(lib.wgrt_last_error(), bytes)


def test_problem_struct_matches_header_order():
    hdr = open(os.path.join(ROOT, "include", "wgrt.h")).read()
    body = hdr[hdr.index("typedef struct wgrt_problem {"):hdr.index("} wgrt_problem_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b([A-Za-z_0-9]+);", body)
    assert names == [f[0] for f in _capi.WgrtProblem._fields_]


def test_no_cpu_fallback(small_scene):
    """Without a CUDA device a launch must raise, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    EB = small_scene.new_matrix_EB()
    with pytest.raises(Exception):
        GRTF.process_rays_kernel_pro_fullColor[1, 256](*small_scene.kernel_args(EB))
    assert EB.sum() == 0


def test_strided_device_views_are_rejected():
    """the C-contiguity check of __cuda_array_interface__ buffers must ignore only the strides of
    axes of extent 1 -- a strided view with a singleton axis is still strided."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200.GPU_ray_tracing_functions import _describe

    class Fake:
        def __init__(self, shape, strides, typestr="<f4"):
            self.__cuda_array_interface__ = {"data": (4096, False), "shape": shape, "typestr": typestr,
                                             "strides": strides, "version": 3}

    _describe(Fake((1, 5, 4), (80, 16, 4)), "ok")              # dense
    _describe(Fake((1, 5, 4), (12345, 16, 4)), "ok")           # singleton axis: its stride is irrelevant
    _describe(Fake((3, 1, 4), (16, 999, 4)), "ok")
    _describe(Fake((0, 4), (1, 1)), "ok")                      # empty
    _describe(Fake((5, 4), None), "ok")
    for shape, strides in (((1, 5, 4), (160, 32, 8)),          # L = 1 table, every other element
                           ((3, 1, 4), (32, 32, 8)),
                           ((5, 4), (4, 20)),                  # transposed
                           ((2, 3, 4), (96, 16, 4))):          # padded rows
        with pytest.raises(ValueError):
            _describe(Fake(shape, strides), "bad")


def test_integration_stub_matches_the_library():
    """INTEGRATION.md section 3 shows the ctypes stub a maintainer would add; its struct must have the size
    the built library reports (the stub had gone stale)."""
    import ctypes as C
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = text[text.index("class wgrt_problem_t(C.Structure)"):]
    block = block[:block.index("lib = C.CDLL")]
    ns = {"C": C}
    exec(block, ns)
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi
    lib = _capi.load_library()
    assert C.sizeof(ns["wgrt_problem_t"]) == lib.wgrt_problem_size() == C.sizeof(_capi.WgrtProblem)
    assert [f[0] for f in ns["wgrt_problem_t"]._fields_] == [f[0] for f in _capi.WgrtProblem._fields_]
    assert lib.wgrt_legacy_problem_size() == C.sizeof(_capi.WgrtLegacyProblem)
    # every entry point named in the table of section 1 exists in the library
    table = text[text.index("## 1."):text.index("## 2.")]
    for name in set(re.findall(r"`(wgrt_[a-z0-9_]+)(?![a-z0-9_.])", table)) - {"wgrt_problem_t"}:
        assert any(e == name or e.startswith(name + "_") for e in _capi.EXPORTED_SYMBOLS), name
