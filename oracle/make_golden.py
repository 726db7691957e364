"""Generate golden vectors by executing the REFERENCE kernel itself (test infrastructure only).

Runs /root/reference/GPU_ray_tracing_functions.py unmodified under Numba's CUDA simulator
(``NUMBA_ENABLE_CUDASIM=1``) -- the reference's only CPU path -- on the deterministic synthetic
scenes of ``synthetic_inputs.make_scene`` and stores the results under tests/golden/.  Two
non-invasive shims are needed: ``matplotlib`` is stubbed in ``sys.modules``
(imported at GRTF:5-7, not installed here) and an int-coercing ``range`` is injected into the
module globals (GRTF:905 calls ``range(1e5)``, legal for the JIT, a TypeError in the simulator).

The reference cannot travel to the GPU box, hence the committed fixtures.  Usage (in the build
container, where /root/reference exists)::

    python oracle/make_golden.py            # writes tests/golden/*.npz  (a few minutes, 8 procs)

Each fixture stores the scene recipe, a SHA-256 of every input array (so a drifting generator is
caught), the final ``rng_states`` and the non-zero bins of ``matrix_EB``.  ``walk_small`` also
stores its complete inputs, so the oracle stays pinned even if the generators change.
"""
from __future__ import annotations

import builtins
import hashlib
import os
import sys
import time
import types

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
os.environ.setdefault("NUMBA_DISABLE_JIT", "0")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE_DIR = os.environ.get("WGRT_REFERENCE_DIR", "/root/reference")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def import_reference():
    """Import the reference module with the two shims."""
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.path", "matplotlib.colors"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.Path = object
            mod.LogNorm = object
            sys.modules[name] = mod
    sys.path.insert(0, REFERENCE_DIR)
    import GPU_ray_tracing_functions as GRTF  # the reference file, unmodified
    GRTF.range = lambda *a: builtins.range(*map(int, a))
    return GRTF


SCENES = {
    # BASELINE.json config 1: single wavelength 532 nm, 5x5 FoV grid, 64 rays per FoV
    "walk_c1": dict(num_FOV_x=5, num_FOV_y=5, num_rays_per_FoV=64, seed=11, lmd_subset=[1], num_iter=2),
    # all three wavelengths, self-contained fixture (inputs stored)
    "walk_small": dict(num_FOV_x=4, num_FOV_y=3, num_rays_per_FoV=48, seed=5, lmd_subset=None, num_iter=1,
                       store_inputs=True),
    # default efficiencies, more rays: statistics for the out-coupler states
    "walk_mid": dict(num_FOV_x=6, num_FOV_y=5, num_rays_per_FoV=200, seed=3, lmd_subset=None, num_iter=1),
    # lossless gratings: long walks, many fold/out-coupler events per ray
    "walk_deep": dict(num_FOV_x=3, num_FOV_y=3, num_rays_per_FoV=400, seed=23, lmd_subset=None, num_iter=1,
                      eff=dict(incouple=0.9, incouple_m1=0.08, ic_zero=0.9, ic_cross=0.05, fc_zero=0.7,
                               fc_turn=0.28, oc_zero=0.85, oc_cross=0.06, outcouple=0.06)),
    # BASELINE config 4: high-resolution eyebox bins (320 x 480 per FoV cell)
    "walk_fine": dict(num_FOV_x=3, num_FOV_y=3, num_rays_per_FoV=240, seed=31, lmd_subset=None, num_iter=1,
                      eb=(320, 480),
                      eff=dict(incouple=0.9, ic_zero=0.95, fc_zero=0.8, fc_turn=0.15, outcouple=0.1)),
    # BASELINE config 5: thin plate, wide FoV, 15 fold slices, strong turn orders (long, branching walks)
    "walk_thin": dict(num_FOV_x=3, num_FOV_y=2, num_rays_per_FoV=300, seed=37, lmd_subset=None, num_iter=1,
                      design=dict(t=0.3, num_FC=15, fov_x_deg=24.0),
                      eff=dict(incouple=0.9, incouple_m1=0.08, ic_zero=0.9, ic_cross=0.05, fc_zero=0.6, fc_turn=0.3,
                               oc_zero=0.85, oc_cross=0.06, outcouple=0.05)),
    # strong polarisation mixing: cross-polarisation amplitudes 0.3-0.7 of the diagonal ones with random
    # phases in every Jones quartet, order efficiencies summing near 1 -- the cancellation in the
    # off-diagonal terms of J^H J that real RCWA tables produce
    "walk_mix": dict(num_FOV_x=3, num_FOV_y=3, num_rays_per_FoV=300, seed=43, lmd_subset=None, num_iter=1,
                     eff=dict(incouple=0.6, incouple_m1=0.25, ic_zero=0.55, ic_cross=0.3, fc_zero=0.5, fc_turn=0.33,
                              oc_zero=0.55, oc_cross=0.25, outcouple=0.08, cross_pol=(0.3, 0.7))),
    # general input polarisation: elliptical states (te, tm in [0.2, 1], delta_phase in [-3, 3]) -- the
    # delta_phase != 0 entry of E_field_cal (GRTF:136-138) that the runner's TE / TM rays never take
    "walk_pol": dict(num_FOV_x=3, num_FOV_y=3, num_rays_per_FoV=300, seed=47, lmd_subset=None, num_iter=1,
                     ray_pol="mixed",
                     eff=dict(incouple=0.8, incouple_m1=0.1, ic_zero=0.85, ic_cross=0.08, fc_zero=0.7, fc_turn=0.25,
                              oc_zero=0.8, oc_cross=0.08, outcouple=0.08, cross_pol=(0.1, 0.3))),
}


def scene_from_recipe(r):
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200.synthetic_inputs import make_scene
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200.couplers_coor import WaveguideDesign
    design = WaveguideDesign(**r["design"]) if r.get("design") else None
    return make_scene(r["num_FOV_x"], r["num_FOV_y"], r["num_rays_per_FoV"], seed=r["seed"],
                      lmd_subset=r.get("lmd_subset"), eff=r.get("eff"), eb=tuple(r.get("eb", (80, 120))),
                      design=design, ray_pol=r.get("ray_pol"))


def input_digest(scene) -> str:
    h = hashlib.sha256()
    for a in scene.kernel_args(scene.new_matrix_EB()):
        if isinstance(a, np.ndarray):
            h.update(str(a.dtype).encode()); h.update(str(a.shape).encode())
            h.update(np.ascontiguousarray(a).tobytes())
        else:
            h.update(repr(float(a)).encode())
    return h.hexdigest()


def _run_chunk(job):
    """Worker: trace rays [lo, hi) for num_iter launches in the simulator."""
    recipe, lo, hi, num_iter = job
    GRTF = import_reference()
    scene = scene_from_recipe(recipe)
    sub = scene.rays.take(slice(lo, hi))
    scene.rays = sub
    # the zero-state reseed (GRTF:28-29) depends on the absolute index; it never triggers for the
    # runner's seeds, which we assert so that chunking stays exact
    assert np.all(sub.rng_states != 0)
    EB = scene.new_matrix_EB()
    rng = sub.rng_states.copy()
    threads = 32
    blocks = (sub.num_rays + threads - 1) // threads
    args = list(scene.kernel_args(EB, rng))
    if recipe.get("ray_pol"):
        # Simulator promotion trap: in the simulator `python_complex * np.float32` is
        # complex64, so E_field_cal's `phase * Etm_abs` (GRTF:138) would lose precision for amplitudes
        # other than 0 / 1 -- the JIT-compiled kernel converts float32 -> float64 first and loses nothing.
        # Feeding the SAME values as float64 arrays makes the simulator compute what the compiled kernel
        # computes (the reference file itself is untouched; the reference's own PTX then confirms this
        # fixture on the GPU with real float32 arrays, tests/test_gpu_reference_kernel.py).
        for k in range(12):
            args[k] = args[k].astype(np.float64)
    for _ in range(num_iter):
        GRTF.process_rays_kernel_pro_fullColor[blocks, threads](*args)
        assert np.all(rng != 0)
    nz = np.flatnonzero(EB)
    return lo, hi, rng, nz, EB.ravel()[nz]


def make_walk(name: str, recipe: dict, procs: int):
    import multiprocessing as mp
    scene = scene_from_recipe(recipe)
    N = scene.rays.num_rays
    num_iter = recipe.get("num_iter", 1)
    step = max(32, (N + 4 * procs - 1) // (4 * procs))
    jobs = [(recipe, lo, min(N, lo + step), num_iter) for lo in range(0, N, step)]
    t0 = time.time()
    with mp.get_context("spawn").Pool(procs) as pool:
        parts = pool.map(_run_chunk, jobs)
    dt = time.time() - t0
    rng = np.zeros(N, dtype=np.uint32)
    EB = scene.new_matrix_EB().ravel()
    for lo, hi, r, nz, val in parts:
        rng[lo:hi] = r
        EB[nz] += val
    nz = np.flatnonzero(EB)
    out = dict(recipe=np.array(repr({k: v for k, v in recipe.items() if k != "store_inputs"})),
               digest=np.array(input_digest(scene)), num_iter=np.array(num_iter),
               rng_states=rng, eb_index=nz.astype(np.int64), eb_value=EB[nz].astype(np.float32),
               eb_shape=np.array(scene.eb_shape, dtype=np.int64),
               sim_seconds=np.array(dt), sim_procs=np.array(procs))
    if recipe.get("store_inputs"):
        args = scene.kernel_args(scene.new_matrix_EB())
        from gpu_ray_tracing_for_waveguide_based_ar_display_b200.GPU_ray_tracing_functions import _ARG_NAMES
        for nm, a in zip(_ARG_NAMES[:-1], args[:-1]):
            out["in_" + nm] = np.asarray(a)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {N} rays x {num_iter} launches in {dt:.1f}s ({N * num_iter / dt:.0f} rays/s, "
          f"{procs} procs), deposits={EB.sum():.0f} -> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")


def make_units(procs: int):
    """Unit-level golden vectors from the reference device functions (plain callables in the simulator)."""
    GRTF = import_reference()
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200.couplers_coor import couplers_coor_full_color
    rs = np.random.default_rng(77)
    # xorshift32, GRTF:25-34
    states = np.concatenate(([0, 1, 0xFFFFFFFF, 0x9E3779B9], rs.integers(1, 2 ** 32, 60))).astype(np.uint32)
    st = states.copy()
    last = np.zeros(len(st))
    for i in range(len(st)):
        for _ in range(7):
            last[i] = GRTF.get_uniform_random_number(st, i)
    # E_field_cal, GRTF:132-152
    n = 400
    ete = rs.uniform(0, 1, n); etm = rs.uniform(0, 1, n)
    ete[:8] = [1, 0, 1, 0, 0.5, 0, 0, 1]; etm[:8] = [0, 1, 0, 1, 0.5, 0, 0, 1]
    delta = rs.uniform(-np.pi, np.pi, n); delta[:4] = 0.0
    jones = (rs.normal(size=(n, 4)) + 1j * rs.normal(size=(n, 4))) * 0.5
    jones[4] = 0.0                       # both outputs below eps -> phases forced to 0
    jones[5, [0, 2]] = 0.0               # te output exactly 0
    jones[6, :] = [1, 0, 0, -1]
    ef = np.zeros((n, 3))
    for i in range(n):
        ef[i] = GRTF.E_field_cal(float(ete[i]), float(etm[i]), float(delta[i]),
                                 complex(jones[i, 0]), complex(jones[i, 1]), complex(jones[i, 2]),
                                 complex(jones[i, 3]))
    # is_inside_or_on_edge over the design's rings, GRTF:36-71
    out = couplers_coor_full_color(5, 5)
    IC, FC, FC_offset, OC, OC_offset, eff_reg1, eff_reg2 = out[:7]
    rings = dict(IC=(IC, np.array([0, len(IC)])), FC=(FC, FC_offset), OC=(OC, OC_offset),
                 eff_reg1=(eff_reg1, np.array([0, len(eff_reg1)])),
                 eff_reg2=(eff_reg2, np.array([0, len(eff_reg2)])))
    loc = {}
    for nm, (verts, off) in rings.items():
        lo = verts.min(0) - 1.0; hi = verts.max(0) + 1.0
        pts = rs.uniform(lo, hi, size=(500, 2))
        # adversarial points: vertices, edge midpoints, points 5e-13 / 5e-12 off an edge
        mids = 0.5 * (verts + np.roll(verts, 1, axis=0))
        nrm = np.roll(verts, 1, axis=0) - verts
        nrm = np.stack((-nrm[:, 1], nrm[:, 0]), 1)
        nrm /= np.maximum(np.hypot(nrm[:, 0], nrm[:, 1]), 1e-300)[:, None]
        extra = np.concatenate((verts, mids, mids + 5e-13 * nrm, mids - 5e-13 * nrm, mids + 5e-12 * nrm,
                                mids - 5e-12 * nrm))
        pts = np.concatenate((pts, extra[rs.permutation(len(extra))[:300]]))
        res = np.full(len(pts), -1, dtype=np.int32)
        for k in range(len(pts)):
            for i in range(len(off) - 1):
                if GRTF.is_inside_or_on_edge(pts[k, 0], pts[k, 1], verts, int(off[i]), int(off[i + 1])):
                    res[k] = i
                    break
        loc[nm + "_verts"] = verts; loc[nm + "_off"] = np.asarray(off, dtype=np.int64)
        loc[nm + "_pts"] = pts; loc[nm + "_hit"] = res
    # is_inside_or_on_edge_4d on per-FoV eyebox rectangles, GRTF:73-108: the design's rectangles, a rotated
    # one and one in another vertex order; points at and around the 1e-9 band of every side and corner
    eff_reg_FOV = out[7]
    r0 = np.asarray(eff_reg_FOV[2, 2], dtype=np.float64)
    c, s_ = np.cos(0.3), np.sin(0.3)
    ctr = r0.mean(0)
    rects = [np.asarray(eff_reg_FOV[0, 0]), np.asarray(eff_reg_FOV[4, 3]),
             (r0 - ctr) @ np.array([[c, -s_], [s_, c]]).T + ctr, r0[[1, 2, 3, 0]],
             np.array([[0.0, 1.0], [0.0, 0.0], [1.0, 0.0], [1.0, 1.0]])]
    offs = np.array([0.0, 5e-13, -5e-13, 5e-12, -5e-12, 5e-10, -5e-10, 2e-9, -2e-9, 1e-6, -1e-6])
    rect_pts, rect_hit = [], []
    for rect in rects:
        pts = []
        for i in range(4):
            a, b = rect[i - 1], rect[i]
            d = b - a
            nrm = np.array([-d[1], d[0]]) / np.hypot(*d)
            for t in (0.0, 1.0, -1e-3, 1.001, 0.5, 0.137, 0.9):
                for o in offs:
                    pts.append(a + t * d + o * nrm)
            for o in offs:
                pts.append(b + o * np.array([1.0, 1.0])); pts.append(b + o * np.array([1.0, -1.0]))
        lo, hi = rect.min(0), rect.max(0)
        pts = np.concatenate((np.array(pts), rs.uniform(lo - 0.3 * (hi - lo), hi + 0.3 * (hi - lo), size=(200, 2))))
        fov = np.ascontiguousarray(rect.reshape(1, 1, 4, 2))
        rect_pts.append(pts)
        rect_hit.append(np.array([bool(GRTF.is_inside_or_on_edge_4d(float(q[0]), float(q[1]), fov, 0, 0)) for q in pts]))
    loc["rect_verts"] = np.array(rects); loc["rect_pts"] = np.array(rect_pts); loc["rect_hit"] = np.array(rect_hit)
    path = os.path.join(GOLDEN_DIR, "units.npz")
    np.savez_compressed(path, xs_in=states, xs_out=st, xs_last=last, ef_ete=ete, ef_etm=etm,
                        ef_delta=delta, ef_jones=jones, ef_out=ef, **loc)
    print(f"units -> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")


def single_lambda_args(scene, lam, EB, rng):
    """The 32 arguments of process_rays_kernel_pro for wavelength index `lam` of a full-colour scene."""
    a = list(scene.kernel_args(None, rng))
    sel = scene.rays.lmd_num == lam
    rays = [np.ascontiguousarray(x[sel]) for x in a[:12]]
    out = rays[:8] + rays[9:12] + [rng]
    out += a[13:23]
    out += [np.ascontiguousarray(a[k][lam]) for k in (23, 24, 25)]              # lut_ic*: [X,Y,C]
    out += [np.ascontiguousarray(a[k][:, lam]) for k in (26, 27, 28, 29)]       # lut_fc*, lut_oc*: [n,X,Y,C]
    out += [np.ascontiguousarray(a[30][lam]), np.ascontiguousarray(a[31][lam]), EB]
    return out, sel


def _run_single_lambda_chunk(job):
    recipe, lam, lo, hi = job
    GRTF = import_reference()
    scene = scene_from_recipe(recipe)
    sel = scene.rays.lmd_num == lam
    rng_all = scene.rays.rng_states[sel].copy()
    EB = np.zeros(scene.eb_shape[1:], dtype=np.float32)
    args, _ = single_lambda_args(scene, lam, EB, rng_all)
    for k in list(range(8)) + [8, 9, 10, 11]:
        args[k] = np.ascontiguousarray(args[k][lo:hi])
    GRTF.process_rays_kernel_pro[(hi - lo + 31) // 32, 32](*args)
    nz = np.flatnonzero(EB)
    return lo, hi, args[11], nz, EB.ravel()[nz]


def make_single_lambda(procs: int):
    """Golden for the single-wavelength twin process_rays_kernel_pro (GRTF:419-831)."""
    import multiprocessing as mp
    recipe = dict(SCENES["walk_deep"]); recipe["num_rays_per_FoV"] = 300
    lam = 1
    scene = scene_from_recipe(recipe)
    n = int(np.count_nonzero(scene.rays.lmd_num == lam))
    step = max(32, (n + 4 * procs - 1) // (4 * procs))
    jobs = [(recipe, lam, lo, min(n, lo + step)) for lo in range(0, n, step)]
    with mp.get_context("spawn").Pool(procs) as pool:
        parts = pool.map(_run_single_lambda_chunk, jobs)
    rng = np.zeros(n, dtype=np.uint32); EB = np.zeros(int(np.prod(scene.eb_shape[1:])), dtype=np.float32)
    for lo, hi, r, nz, val in parts:
        rng[lo:hi] = r; EB[nz] += val
    nz = np.flatnonzero(EB)
    path = os.path.join(GOLDEN_DIR, "walk_single_lambda.npz")
    np.savez_compressed(path, recipe=np.array(repr(recipe)), lam=np.array(lam), digest=np.array(input_digest(scene)),
                        rng_states=rng, eb_index=nz.astype(np.int64), eb_value=EB[nz], eb_shape=np.array(scene.eb_shape[1:]))
    print(f"walk_single_lambda: {n} rays, deposits={EB.sum():.0f} -> {path}")


def make_eval(procs: int):
    """U_fov / U_EB / output_image / pupil sums from the REFERENCE evaluation() itself.  Only `colour`
    (not installed) is stubbed -- with this repository's restatement of Lab / CIEDE2000 -- so delta_e is
    NOT a reference value (parity unpinned) while everything else is."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import AR_system_evaluation_functions as mine
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import synthetic_inputs as si
    from oracle import oracle
    col = types.ModuleType("colour")
    col.SDS_ILLUMINANTS = {"D65": "D65"}
    col.sd_to_XYZ = lambda sd: mine._XYZ_D65_SPD.copy()
    x_w, y_w = mine._WHITE_XY
    white = np.array([x_w / y_w, 1.0, (1 - x_w - y_w) / y_w]) * 100.0
    col.XYZ_to_Lab = lambda xyz: mine._xyz_to_lab(np.asarray(xyz, dtype=np.float64), white)
    col.delta_E = lambda a, b, method="CIE 2000": mine._delta_e_2000(np.asarray(a), np.asarray(b))
    sys.modules["colour"] = col
    sys.path.insert(0, REFERENCE_DIR)
    import AR_system_evaluation_functions as REF
    eff = dict(incouple=0.9, incouple_m1=0.05, ic_zero=0.95, ic_cross=0.02, fc_zero=0.8, fc_turn=0.18,
               oc_zero=0.85, oc_cross=0.02, outcouple=0.12)
    scene = si.make_scene(6, 5, 20000, seed=41, eff=eff)
    EB = scene.new_matrix_EB(); rng = scene.rays.rng_states.copy()
    num_iter = 2
    for _ in range(num_iter):
        oracle.trace(*scene.kernel_args(EB, rng), num_threads=procs)
    EB2 = EB / 20000 / num_iter                                  # RUN:197
    captured = {}
    real_sum = np.sum
    delta_e, U_fov, U_EB, output_image = REF.evaluation(EB2)
    # pupil sums restated exactly as the reference lines 91-109 (they are not returned by evaluation())
    size = 30; radius = size / 2
    yy, xx = np.ogrid[:size, :size]
    mask = (np.sqrt((xx - (radius - 0.5)) ** 2 + (yy - (radius - 0.5)) ** 2) <= radius).astype(np.float32)
    y0l = np.arange(0, 80 - size + 1, 8); x0l = np.arange(0, 120 - size + 1, 12)
    perceive = np.zeros(EB2.shape[:3] + (len(y0l), len(x0l)), dtype=EB2.dtype)
    for iy, y0 in enumerate(y0l):
        for ix, x0 in enumerate(x0l):
            perceive[:, :, :, iy, ix] = np.sum(EB2[:, :, :, y0:y0 + size, x0:x0 + size] * mask[None, None, None], axis=(-1, -2))
    nz = np.flatnonzero(EB)
    path = os.path.join(GOLDEN_DIR, "eval.npz")
    np.savez_compressed(path, eb_index=nz.astype(np.int64), eb_value=EB.ravel()[nz].astype(np.float32),
                        eb_shape=np.array(EB.shape), rays_per_fov=np.array(20000), num_iter=np.array(num_iter),
                        U_fov=np.array(U_fov), U_EB=np.array(U_EB), delta_e_stubbed=np.array(delta_e),
                        output_image=output_image.astype(np.float32), perceive=perceive)
    print(f"eval: deposits={EB.sum():.0f} U_fov={U_fov:.4f} U_EB={U_EB:.4f} delta_e(stub)={delta_e:.3f} -> {path} "
          f"({os.path.getsize(path) / 1e3:.0f} kB)")


LEGACY_RECIPE = dict(num_FOV_x=3, num_FOV_y=2, scene_seed=5, lam=1, lut_seed=3, points=4, point_seed=7, max_steps=400,
                     generations=5, capacity=8192, eb=(80, 120))


def legacy_inputs(r=LEGACY_RECIPE):
    """(geometry, tables, initial rows) of the legacy-tracer fixture."""
    from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import legacy, synthetic_inputs as si
    scene = si.make_scene(r["num_FOV_x"], r["num_FOV_y"], 2, seed=r["scene_seed"], build_rays=False)
    g, l = legacy.make_legacy_luts(scene, r["lam"], seed=r["lut_seed"])
    pts = si.points_in_disc(scene.geom["IC"], r["points"], r["point_seed"])
    return g, l, legacy.initial_rows(pts, r["num_FOV_x"], r["num_FOV_y"])


def sort_rows(rows):
    """Canonical order of ray rows (the append order of children depends on thread scheduling)."""
    rows = np.asarray(rows)
    return rows[np.lexsort(rows.T[::-1])]


def make_legacy(procs: int):
    """Golden for the legacy energy-splitting kernel process_rays_kernel (GRTF:192-417) and pack_active_to_front
    (GRTF:178-190): `generations` rounds of launch + compaction of the unmodified reference kernels under the
    simulator; after every launch all rows (sorted), the child counter, and after every compaction the packed
    rows (sorted); finally the bins."""
    GRTF = import_reference()
    r = LEGACY_RECIPE
    g, l, rows0 = legacy_inputs(r)
    cap = r["capacity"]
    a = np.zeros((cap, 13)); b = np.zeros((cap, 13))
    a[:len(rows0)] = rows0
    EB = np.zeros((r["num_FOV_y"], r["num_FOV_x"]) + tuple(r["eb"]), dtype=np.float32)
    count = len(rows0)
    out = dict(recipe=np.array(repr(r)))
    h = hashlib.sha256()
    for arr in list(g.values()) + list(l.values()) + [rows0]:
        h.update(np.ascontiguousarray(arr).tobytes())
    out["digest"] = np.array(h.hexdigest())
    t0 = time.time()
    for gen in range(r["generations"]):
        counter = np.array([count], dtype=np.int32)
        GRTF.process_rays_kernel[(count + 31) // 32, 32](
            a, count, counter, r["max_steps"], g["IC"], g["FC"], g["FC_offset"], g["OC"], g["OC_offset"], g["eff_reg1"],
            g["eff_reg2"], g["eff_reg_FOV"], g["eff_reg_FOV_range"], l["lut_ic1"], l["lut_ic2"], l["lut_fc1"], l["lut_fc2"],
            l["lut_oc"], g["lut_TIR"], g["lut_gap"], EB)
        total = int(counter[0])
        assert total <= cap
        out[f"launch{gen}_rows"] = sort_rows(a[:total])
        out[f"launch{gen}_counter"] = np.array(total)
        oc = np.zeros(1, dtype=np.int32)
        b[:] = 0
        GRTF.pack_active_to_front[(total + 31) // 32, 32](a, b, total, oc)
        count = int(oc[0])
        out[f"pack{gen}_rows"] = sort_rows(b[:count])
        a, b = b, a
    nz = np.flatnonzero(EB)
    out.update(eb_index=nz.astype(np.int64), eb_value=EB.ravel()[nz], eb_shape=np.array(EB.shape),
               sim_seconds=np.array(time.time() - t0))
    path = os.path.join(GOLDEN_DIR, "legacy.npz")
    np.savez_compressed(path, **out)
    print(f"legacy: {len(rows0)} initial rows, {r['generations']} generations, live {count}, energy {EB.sum():.6f} in "
          f"{len(nz)} bins, {time.time() - t0:.1f}s -> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")


if __name__ == "__main__":
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    procs = int(os.environ.get("WGRT_GOLDEN_PROCS", os.cpu_count() or 1))
    which = sys.argv[1:] or ["units", "eval", "single_lambda", "legacy"] + list(SCENES)
    for w in which:
        if w == "units":
            make_units(procs)
        elif w == "eval":
            make_eval(procs)
        elif w == "single_lambda":
            make_single_lambda(procs)
        elif w == "legacy":
            make_legacy(procs)
        else:
            make_walk(w, SCENES[w], procs)
