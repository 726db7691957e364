"""Deterministic synthetic inputs for the ray-propagation hot path.

The reference loads seven RCWA look-up tables with ``np.load`` and fetches them from Google
Drive (/root/reference/download_lut.py:13-19, gpu_ray_tracing_pro_fullColor.py:28-34); those
files are not reachable here.  ``make_luts`` builds complex128 arrays of the shapes the kernel
indexes:

    lut_ic1, lut_ic2, lut_ic3 : [L, X, Y, 42]
    lut_fc1, lut_fc2          : [nFC, L, X, Y, 26]
    lut_oc1, lut_oc2          : [nOC, L, X, Y, 42]

channel 0 = polar angle of the outgoing direction, channel 1 = azimuth, then 8 groups of P
orders (P = 5 for IC/OC tables, 3 for FC tables).  A Jones quartet for order p uses groups
(g, g+1, g+4, g+5), g = 0 for glass->glass events and g = 2 for air<->glass events.

``build_ray_set`` lays the rays out exactly as the reference runner does
(gpu_ray_tracing_pro_fullColor.py:59-158): for every FoV cell (x outer, y inner) and wavelength,
one block of ``num_rays_per_FoV`` rays whose first half is TE and second half TM, all sharing the
same ``num_rays_per_FoV/2`` start points inside the in-coupler.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np

from .couplers_coor import WaveguideDesign, couplers_coor_full_color

__all__ = ["make_luts", "build_ray_set", "RaySet", "Scene", "make_scene", "initial_rng_states",
           "LUT_NAMES", "RAY_FIELDS"]

LUT_NAMES = ("lut_ic1", "lut_ic2", "lut_ic3", "lut_fc1", "lut_fc2", "lut_oc1", "lut_oc2")
RAY_FIELDS = ("x", "y", "gap_x", "gap_y", "pol", "azi", "m", "n", "lmd_num", "te", "tm", "delta_phase")


def _fill_lut(shape_prefix, P, theta, phi, diag_targets, rng, cross=0.02, floor=0.01, cross_range=None):
    """One LUT. ``diag_targets[(g0, p)]`` = |diagonal Jones| array broadcastable to shape_prefix.
    ``cross_range=(lo, hi)``: cross-polarisation amplitudes uniform in [lo, hi] x the diagonal one
    (strong polarisation mixing) instead of ``cross`` x U(0.5, 1.5); same random stream either way."""
    def xpol(amp, u):
        if cross_range is None:
            return cross * amp * u
        return (cross_range[0] + (cross_range[1] - cross_range[0]) * (u - 0.5)) * amp
    C = 2 + 8 * P
    mag = np.full(shape_prefix + (C,), floor)
    mag *= rng.uniform(0.5, 1.5, size=mag.shape)
    for (g0, p), amp in diag_targets.items():
        amp = np.broadcast_to(amp, shape_prefix)
        c_tt = 2 + g0 * P + p
        c_x1 = 2 + (g0 + 1) * P + p
        c_x2 = 2 + (g0 + 4) * P + p
        c_mm = 2 + (g0 + 5) * P + p
        mag[..., c_tt] = amp * rng.uniform(0.95, 1.05, size=shape_prefix)
        mag[..., c_mm] = 0.93 * amp * rng.uniform(0.95, 1.05, size=shape_prefix)
        mag[..., c_x1] = xpol(amp, rng.uniform(0.5, 1.5, size=shape_prefix))
        mag[..., c_x2] = xpol(amp, rng.uniform(0.5, 1.5, size=shape_prefix))
    phase = rng.uniform(-np.pi, np.pi, size=mag.shape)
    lut = (mag * np.exp(1j * phase)).astype(np.complex128)
    lut[..., 0] = np.broadcast_to(theta, shape_prefix)
    lut[..., 1] = np.broadcast_to(phi, shape_prefix)
    return lut


def make_luts(angles: Dict[str, np.ndarray], n_g: float, n_fc: int, n_oc: int, seed: int = 0,
              eff: Optional[Dict[str, float]] = None) -> Dict[str, np.ndarray]:
    """Synthetic RCWA tables with physically plausible per-event efficiencies.

    ``angles``: the [L, X, Y] angle tables of the design (th_in_ic, th_out_ic, th_out_ic2,
    th_out_fc, th_out_oc and the matching phi_*).  ``eff`` overrides the target efficiencies.
    """
    e = dict(incouple=0.25, incouple_m1=0.03, ic_zero=0.82, ic_cross=0.02,
             fc_zero=0.86, fc_turn=0.08, oc_zero=0.80, oc_cross=0.03, outcouple=0.12)
    if eff:
        e.update(eff)
    # eff["cross_pol"] = (lo, hi): strong polarisation mixing in every Jones quartet (see _fill_lut)
    xr = tuple(e["cross_pol"]) if e.get("cross_pol") is not None else None
    rng = np.random.default_rng(seed)

    def fill(*a):   # every table of this call shares the mixing option
        return _fill_lut(*a, cross_range=xr)
    th_in, th_ic, th_ic2 = angles["th_in_ic"], angles["th_out_ic"], angles["th_out_ic2"]
    th_fc, th_oc = angles["th_out_fc"], angles["th_out_oc"]
    c_in, c_ic, c_ic2, c_fc, c_oc = (np.cos(a) for a in (th_in, th_ic, th_ic2, th_fc, th_oc))
    L, X, Y = th_in.shape
    base = (L, X, Y)
    # smooth FoV / wavelength modulation so the maps are not flat
    gx = np.linspace(-1, 1, X)[None, :, None]
    gy = np.linspace(-1, 1, Y)[None, None, :] if Y > 1 else np.zeros((1, 1, 1))
    gl = np.linspace(-1, 1, L)[:, None, None] if L > 1 else np.zeros((1, 1, 1))
    mod = 1.0 + 0.10 * gx - 0.07 * gy + 0.05 * gl

    def amp(target, factor):
        return np.sqrt(np.clip(target, 0, None) / factor)

    luts = {}
    # in-coupling from air: eff = |J E|^2 * cos(th_new) / cos(th_in) * n_g     (GRTF:868-869)
    luts["lut_ic1"] = fill(base, 5, th_in, angles["phi_in_ic"], {
        (2, 1): amp(e["incouple"] * mod, c_ic / c_in * n_g),
        (2, 3): amp(e["incouple_m1"] * mod, c_ic2 / c_in * n_g)}, rng)
    # inside the in-coupler, +1 direction: eff = |J E|^2 * cos(th_new)/cos(th_cur)   (GRTF:917-918)
    luts["lut_ic2"] = fill(base, 5, th_ic, angles["phi_out_ic"], {
        (0, 2): amp(e["ic_zero"], 1.0),
        (0, 4): amp(e["ic_cross"], c_ic2 / c_ic)}, rng)
    luts["lut_ic3"] = fill(base, 5, th_ic2, angles["phi_out_ic2"], {
        (0, 0): amp(e["ic_cross"], c_ic / c_ic2),
        (0, 2): amp(e["ic_zero"], 1.0)}, rng)
    # fold coupler: slices further from the in-coupler turn a little more light
    sl_fc = (1.0 + 0.5 * np.arange(n_fc) / max(n_fc - 1, 1))[:, None, None, None]
    fbase = (n_fc,) + base
    luts["lut_fc1"] = fill(fbase, 3, th_ic, angles["phi_out_ic"], {
        (0, 1): amp(e["fc_zero"], 1.0),
        (0, 0): amp(e["fc_turn"] * sl_fc * mod, c_fc / c_ic)}, rng)
    luts["lut_fc2"] = fill(fbase, 3, th_fc, angles["phi_out_fc"], {
        (0, 2): amp(e["fc_turn"] * sl_fc, c_ic / c_fc),
        (0, 1): amp(e["fc_zero"], 1.0)}, rng)
    sl_oc = (1.0 + 0.8 * np.arange(n_oc) / max(n_oc - 1, 1))[:, None, None, None]
    obase = (n_oc,) + base
    luts["lut_oc1"] = fill(obase, 5, th_fc, angles["phi_out_fc"], {
        (0, 2): amp(e["oc_zero"], 1.0),
        (0, 0): amp(e["oc_cross"], c_oc / c_fc),
        (2, 1): amp(e["outcouple"] * sl_oc * mod, c_in / c_fc / n_g)}, rng)
    luts["lut_oc2"] = fill(obase, 5, th_oc, angles["phi_out_oc"], {
        (0, 4): amp(e["oc_cross"], c_fc / c_oc),
        (0, 2): amp(e["oc_zero"], 1.0),
        (2, 3): amp(e["outcouple"] * sl_oc * mod, c_in / c_oc / n_g)}, rng)
    return luts


def initial_rng_states(num_rays: int, offset: int = 0) -> np.ndarray:
    """Per-ray xorshift32 seeds, gpu_ray_tracing_pro_fullColor.py:158."""
    return (np.uint32(0x9E3779B9) * (np.arange(offset, offset + num_rays, dtype=np.uint32) + np.uint32(1)))


def points_in_disc(IC: np.ndarray, num_points: int, seed: int) -> np.ndarray:
    """Seeded uniform points strictly inside the (convex) in-coupler ring."""
    from .GPU_ray_tracing_functions import generate_points_in_polygon
    return generate_points_in_polygon(IC, num_points, rng=np.random.default_rng(seed))


@dataclass
class RaySet:
    """The twelve float32 SoA arrays + uint32 RNG states the kernel takes (RUN:65-76, 158)."""
    x: np.ndarray
    y: np.ndarray
    gap_x: np.ndarray
    gap_y: np.ndarray
    pol: np.ndarray
    azi: np.ndarray
    m: np.ndarray
    n: np.ndarray
    lmd_num: np.ndarray
    te: np.ndarray
    tm: np.ndarray
    delta_phase: np.ndarray
    rng_states: np.ndarray

    @property
    def num_rays(self) -> int:
        return int(self.x.shape[0])

    def arrays(self):
        return tuple(getattr(self, f) for f in RAY_FIELDS)

    def take(self, sl) -> "RaySet":
        return RaySet(*(getattr(self, f)[sl].copy() for f in RAY_FIELDS), self.rng_states[sl].copy())


def build_ray_set(points: np.ndarray, num_FOV_x: int, num_FOV_y: int, n_lmd: int,
                  num_rays_per_FoV: int, lmd_subset: Optional[Sequence[int]] = None,
                  cells: Optional[np.ndarray] = None, ray_pol: Optional[str] = None, pol_seed: int = 0) -> RaySet:
    """Ray arrays in the runner's order (RUN:82-115).

    ``points``: [num_rays_per_FoV//2, 2] start points.  ``lmd_subset`` restricts the wavelength
    indices that get rays (BASELINE config 1 traces 532 nm only).  ``cells`` optionally gives an
    explicit [K,3] list of (m, n, lmd) cells instead of the full grid (used by the multi-GPU
    partitioner, which hands each rank a contiguous range of the runner's cell sequence).
    """
    half = num_rays_per_FoV // 2
    assert points.shape == (half, 2) and 2 * half == num_rays_per_FoV
    if cells is None:
        lm = np.arange(n_lmd) if lmd_subset is None else np.asarray(list(lmd_subset))
        ii, jj, ll = np.meshgrid(np.arange(num_FOV_x), np.arange(num_FOV_y), lm, indexing="ij")
        cells = np.stack((ii.ravel(), jj.ravel(), ll.ravel()), axis=1)
    K = len(cells)
    N = K * num_rays_per_FoV
    f32 = np.float32
    px = points[:, 0].astype(f32)
    py = points[:, 1].astype(f32)
    x = np.tile(np.concatenate((px, px)), K)
    y = np.tile(np.concatenate((py, py)), K)
    zeros = np.zeros(N, dtype=f32)
    m = np.repeat(cells[:, 0].astype(f32), num_rays_per_FoV)
    n = np.repeat(cells[:, 1].astype(f32), num_rays_per_FoV)
    lm_arr = np.repeat(cells[:, 2].astype(f32), num_rays_per_FoV)
    te_block = np.concatenate((np.ones(half, f32), np.zeros(half, f32)))
    te = np.tile(te_block, K)
    tm = np.tile(1.0 - te_block, K).astype(f32)
    dl = zeros.copy()
    if ray_pol == "mixed":
        # general input polarisation (the kernel's signature allows it; the runner only uses TE / TM with
        # delta_phase = 0): elliptical states, amplitudes and phase difference per ray
        prng = np.random.default_rng(pol_seed)
        te = prng.uniform(0.2, 1.0, N).astype(f32)
        tm = prng.uniform(0.2, 1.0, N).astype(f32)
        dl = prng.uniform(-3.0, 3.0, N).astype(f32)
    elif ray_pol is not None:
        raise ValueError("ray_pol must be None or 'mixed'")
    return RaySet(x, y, zeros.copy(), zeros.copy(), zeros.copy(), zeros.copy(), m, n, lm_arr,
                  te, tm, dl, initial_rng_states(N))


@dataclass
class Scene:
    """Everything one launch of the kernel needs, as host arrays."""
    geom: Dict[str, np.ndarray]
    n_g: float
    luts: Dict[str, np.ndarray]
    rays: RaySet
    eb_shape: tuple          # (L, num_FOV_y, num_FOV_x, EBy, EBx)
    meta: Dict[str, object]

    def new_matrix_EB(self) -> np.ndarray:
        return np.zeros(self.eb_shape, dtype=np.float32)

    def kernel_args(self, matrix_EB, rng_states=None):
        """The 33 positional arguments of process_rays_kernel_pro_fullColor (GRTF:834-841)."""
        g, l = self.geom, self.luts
        rs = self.rays.rng_states if rng_states is None else rng_states
        return (*self.rays.arrays(), rs,
                g["IC"], g["FC"], g["FC_offset"], g["OC"], g["OC_offset"], self.n_g,
                g["eff_reg1"], g["eff_reg2"], g["eff_reg_FOV"], g["eff_reg_FOV_range"],
                l["lut_ic1"], l["lut_ic2"], l["lut_ic3"], l["lut_fc1"], l["lut_fc2"],
                l["lut_oc1"], l["lut_oc2"], g["lut_TIR"], g["lut_gap"], matrix_EB)


def make_scene(num_FOV_x: int, num_FOV_y: int, num_rays_per_FoV: int, eb=(80, 120), seed: int = 0,
               lmd_subset: Optional[Sequence[int]] = None, design: Optional[WaveguideDesign] = None,
               eff: Optional[Dict[str, float]] = None, build_rays: bool = True,
               ray_pol: Optional[str] = None) -> Scene:
    out = couplers_coor_full_color(num_FOV_x, num_FOV_y, design=design)
    (IC, FC, FC_offset, OC, OC_offset, eff_reg1, eff_reg2, eff_reg_FOV, eff_reg_FOV_range,
     lut_TIR, lut_gap, _fres, _Lic, _pic, _Lfc, _pfc, _Loc, _poc, n_g, lmd,
     th_in_ic, phi_in_ic, th_out_ic, phi_out_ic, th_out_fc, phi_out_fc,
     th_out_ic2, phi_out_ic2, th_out_oc, phi_out_oc, _glow, *_k) = out
    geom = dict(IC=np.ascontiguousarray(IC), FC=np.ascontiguousarray(FC),
                FC_offset=np.asarray(FC_offset, dtype=np.int64),
                OC=np.ascontiguousarray(OC), OC_offset=np.asarray(OC_offset, dtype=np.int64),
                eff_reg1=np.ascontiguousarray(eff_reg1), eff_reg2=np.ascontiguousarray(eff_reg2),
                eff_reg_FOV=np.ascontiguousarray(eff_reg_FOV),
                eff_reg_FOV_range=np.ascontiguousarray(eff_reg_FOV_range),
                lut_TIR=np.ascontiguousarray(lut_TIR), lut_gap=np.ascontiguousarray(lut_gap))
    angles = dict(th_in_ic=th_in_ic, phi_in_ic=phi_in_ic, th_out_ic=th_out_ic, phi_out_ic=phi_out_ic,
                  th_out_fc=th_out_fc, phi_out_fc=phi_out_fc, th_out_ic2=th_out_ic2,
                  phi_out_ic2=phi_out_ic2, th_out_oc=th_out_oc, phi_out_oc=phi_out_oc)
    luts = make_luts(angles, float(n_g), len(FC_offset) - 1, len(OC_offset) - 1, seed=seed, eff=eff)
    L = len(lmd)
    if build_rays:
        pts = points_in_disc(IC, num_rays_per_FoV // 2, seed + 1)
        rays = build_ray_set(pts, num_FOV_x, num_FOV_y, L, num_rays_per_FoV, lmd_subset, ray_pol=ray_pol,
                             pol_seed=seed + 2)
    else:
        rays = None
    return Scene(geom, float(n_g), luts, rays, (L, num_FOV_y, num_FOV_x, eb[0], eb[1]),
                 dict(num_FOV_x=num_FOV_x, num_FOV_y=num_FOV_y, num_rays_per_FoV=num_rays_per_FoV,
                      seed=seed, lmd_subset=None if lmd_subset is None else list(lmd_subset)))
