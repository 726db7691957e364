"""GPU parity against the REFERENCE's own kernels, compiled by Numba to PTX (oracle/build_ref_ptx.py)
and launched through the CUDA driver on the B200 (oracle/ref_numba_cuda.py).

This pins the engine -- and the CPU oracle -- to the reference at sizes the CPU simulator can never
reach: millions of rays through GPU_ray_tracing_functions.py:833-1246 itself, with Numba's own FMA
contraction and libdevice transcendentals.  Bar: bit-equal ``matrix_EB`` and ``rng_states``.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import GPU_ray_tracing_functions as GRTF
from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import synthetic_inputs as si
from gpu_ray_tracing_for_waveguide_based_ar_display_b200.couplers_coor import WaveguideDesign
from oracle import ref_numba_cuda as ref

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref PTX not built (needs /root/reference)")

KERNEL = GRTF.process_rays_kernel_pro_fullColor


def to_device(args):
    import torch
    out = []
    for a in args:
        if isinstance(a, np.ndarray):
            v = a.view(np.float64) if a.dtype == np.complex128 else a
            t = torch.from_numpy(np.ascontiguousarray(v.view(np.int32) if v.dtype == np.uint32 else v)).cuda()
            out.append(GRTF._TorchAlias(t, a.shape, a.dtype))
        else:
            out.append(a)
    return out


def run_reference(scene, num_iter=1):
    import torch
    EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
    d = to_device(scene.kernel_args(EB, rng))
    for _ in range(num_iter):
        ref.launch("process_rays_kernel_pro_fullColor", d, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d[32]._t.cpu().numpy().reshape(EB.shape), d[12]._t.cpu().numpy().view(np.uint32)


def run_engine(kernel, scene, num_iter=1):
    EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
    for _ in range(num_iter):
        kernel[(scene.rays.num_rays + 255) // 256, 256](*scene.kernel_args(EB, rng))
    return EB, rng


def assert_same(a, b, what):
    EBa, ra = a; EBb, rb = b
    bad = np.flatnonzero(ra != rb)
    assert bad.size == 0, f"{what}: rng_states differ for {bad.size} rays, first {bad[:8]}"
    assert np.array_equal(EBa, EBb), f"{what}: matrix_EB differs in {np.count_nonzero(EBa != EBb)} bins"


@needs_ref
def test_reference_kernel_reproduces_simulator_golden():
    """The PTX really is the reference: it reproduces the fixtures its CPU simulator produced."""
    from conftest import golden_bins, load_golden_walk
    for name in ("walk_small", "walk_c1", "walk_mid", "walk_deep", "walk_fine", "walk_thin", "walk_mix", "walk_pol"):
        scene, g = load_golden_walk(name)
        got = run_reference(scene, int(g["num_iter"]))
        assert_same(got, (golden_bins(g), g["rng_states"]), name)


@needs_ref
@pytest.mark.parametrize("cfg", [
    dict(nx=5, ny=5, rays=64, seed=101, lmd=[1], it=4),             # BASELINE config 1
    dict(nx=20, ny=15, rays=5000, seed=201, lmd=None, it=2),        # 4.5 M rays, runner-sized cells
    dict(nx=41, ny=41, rays=1000, seed=202, lmd=None, it=1),        # BASELINE config 3 grid, 5 M rays
    # strong polarisation mixing (cross-pol 0.3-0.7 of the diagonal Jones entries, random phases, order
    # efficiencies summing near 1): 1.44 M rays x 2 launches
    dict(nx=12, ny=10, rays=4000, seed=205, lmd=None, it=2,
         eff=dict(incouple=0.6, incouple_m1=0.25, ic_zero=0.55, ic_cross=0.3, fc_zero=0.5, fc_turn=0.33,
                  oc_zero=0.55, oc_cross=0.25, outcouple=0.08, cross_pol=(0.3, 0.7))),
    # general (elliptical) input polarisation, delta_phase != 0
    dict(nx=12, ny=10, rays=4000, seed=206, lmd=None, it=1, ray_pol="mixed",
         eff=dict(incouple=0.8, incouple_m1=0.1, ic_zero=0.85, ic_cross=0.08, fc_zero=0.7, fc_turn=0.25,
                  oc_zero=0.8, oc_cross=0.08, outcouple=0.08, cross_pol=(0.1, 0.3))),
])
def test_engine_equals_reference_kernel(cfg):
    scene = si.make_scene(cfg["nx"], cfg["ny"], cfg["rays"], seed=cfg["seed"], lmd_subset=cfg["lmd"],
                          eff=cfg.get("eff"), ray_pol=cfg.get("ray_pol"))
    want = run_reference(scene, cfg["it"])
    assert want[0].sum() > 0
    assert_same(run_engine(KERNEL, scene, cfg["it"]), want, "fast vs reference kernel")
    assert_same(run_engine(KERNEL.configured(strict=True), scene, cfg["it"]), want, "strict vs reference kernel")


@needs_ref
def test_engine_equals_reference_kernel_deep_walks():
    """BASELINE config 5 (divergence stress) at 1.4 M rays."""
    d = WaveguideDesign(t=0.3, num_FC=15, fov_x_deg=24.0)
    eff = dict(incouple=0.9, incouple_m1=0.08, ic_zero=0.9, ic_cross=0.05, fc_zero=0.6, fc_turn=0.3,
               oc_zero=0.85, oc_cross=0.06, outcouple=0.05)
    scene = si.make_scene(12, 10, 4000, seed=77, design=d, eff=eff)
    want = run_reference(scene)
    assert_same(run_engine(KERNEL, scene), want, "fast/deep vs reference kernel")


@needs_ref
def test_engine_equals_reference_kernel_fine_eyebox():
    """BASELINE config 4 (320 x 480 bins per FoV cell)."""
    scene = si.make_scene(10, 8, 4000, eb=(320, 480), seed=78,
                          eff=dict(incouple=0.9, ic_zero=0.95, fc_zero=0.8, fc_turn=0.15, outcouple=0.1))
    want = run_reference(scene)
    assert_same(run_engine(KERNEL, scene), want, "fast/fine eyebox vs reference kernel")


@needs_ref
def test_cpu_oracle_equals_reference_kernel(oracle):
    """Closes the triangle: CPU oracle == reference GPU kernel on a case too big for the simulator."""
    scene = si.make_scene(8, 6, 2000, seed=203)
    want = run_reference(scene, 2)
    EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
    for _ in range(2):
        oracle.trace(*scene.kernel_args(EB, rng))
    assert_same((EB, rng), want, "CPU oracle vs reference kernel")


@needs_ref
def test_single_lambda_twin_equals_reference_kernel():
    """process_rays_kernel_pro (GRTF:419-831): the 32-argument single-wavelength kernel."""
    from oracle import make_golden
    import torch
    scene = si.make_scene(9, 7, 3000, seed=204)
    lam = 2
    sel = scene.rays.lmd_num == lam
    shape = scene.eb_shape[1:]
    EB_r = np.zeros(shape, np.float32); rng_r = scene.rays.rng_states[sel].copy()
    args_r, _ = make_golden.single_lambda_args(scene, lam, EB_r, rng_r)
    d = to_device(args_r)
    ref.launch("process_rays_kernel_pro", d, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = (d[31]._t.cpu().numpy().reshape(shape), d[11]._t.cpu().numpy().view(np.uint32))
    assert want[0].sum() > 0
    EB = np.zeros(shape, np.float32); rng = scene.rays.rng_states[sel].copy()
    args, _ = make_golden.single_lambda_args(scene, lam, EB, rng)
    GRTF.process_rays_kernel_pro[(len(rng) + 255) // 256, 256](*args)
    assert_same((EB, rng), want, "single-lambda twin vs reference kernel")
