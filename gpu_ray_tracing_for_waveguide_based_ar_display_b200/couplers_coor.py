"""Waveguide design -> coupler polygons and analytic per-FoV tables.

Producer of the hot path's geometry inputs.  It
restates what ``couplers_coor.couplers_coor_full_color`` computes
(/root/reference/couplers_coor.py:122-750) without shapely / matplotlib, which
are not installed here, and with the design constants lifted into a
``WaveguideDesign`` record so that the stress configurations of BASELINE.json
(thin plate, wide FoV, many fold-coupler slices) can be generated too.

Differences from the reference are confined to *how* things are computed:

* the k-space sweeps are vectorised NumPy instead of triple Python loops;
* ``shapely`` band clipping (couplers_coor.py:417-452, 558-600) is a
  Sutherland-Hodgman clip of a convex ring against the band rectangle;
* ``LineString.simplify(1e-3)`` (couplers_coor.py:402-404, 552-554) is a plain
  Douglas-Peucker pass on the open hull ring.

Vertex order / ring start of the clipped slices may differ from GEOS; the ray
walk only ever runs even-odd containment tests on these rings, so that does not
change any result.  The arrays returned here are *inputs* to the oracle and to
the CUDA engine alike, hence they cannot affect parity.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np
from scipy.spatial import ConvexHull

deg = np.pi / 180.0

__all__ = ["WaveguideDesign", "couplers_coor_full_color", "design_tables"]


@dataclass
class WaveguideDesign:
    """Design constants (defaults = couplers_coor.py:125-188)."""

    fov_x_deg: float = 18.0
    aspect: float = 4.0 / 3.0
    lmd_nm: Sequence[float] = (465, 532, 630)
    n_g: float = 1.9
    n_air: float = 1.0
    plate_x: float = 60.0          # half-width used for the slicing bands
    plate_y: float = 50.0
    t: float = 0.7                 # substrate thickness (mm)
    num_FC: int = 7
    num_OC: int = 6
    r: float = 2.0                 # in-coupler radius (mm)
    x_ic0: float = -28.0
    y_ic0: float = 15.0
    n_ic_pts: int = 100
    x_eb: float = 12.0
    y_eb: float = 8.0
    er: float = -20.0              # eye relief (signed)
    x_eb0: float = 0.0
    y_eb0: float = 15.0
    Lambda_ic: float = 388.0
    phi_ic_deg: float = -38.0
    Lambda_oc: float = 388.0
    phi_oc_deg: float = -142.0
    n_sweep: int = 50              # FoV samples of the layout sweep (couplers_coor.py:128-129)
    simplify_tol: float = 1e-3

    lmd: np.ndarray = field(init=False, repr=False)

    def __post_init__(self):
        self.lmd = np.asarray(self.lmd_nm)


# --------------------------------------------------------------------------
# small planar-geometry helpers
# --------------------------------------------------------------------------
def _hull_indices(px: np.ndarray, py: np.ndarray) -> np.ndarray:
    return ConvexHull(np.column_stack((px, py))).vertices


def _clip_halfplane(ring: List[Tuple[float, float]], axis: int, bound: float, keep_le: bool):
    """One Sutherland-Hodgman pass: keep the side ``coord <= bound`` (or >=)."""
    out: List[Tuple[float, float]] = []
    n = len(ring)
    if n == 0:
        return out

    def inside(p):
        return p[axis] <= bound if keep_le else p[axis] >= bound

    def cross_pt(p, q):
        tpar = (bound - p[axis]) / (q[axis] - p[axis])
        other = p[1 - axis] + tpar * (q[1 - axis] - p[1 - axis])
        return (bound, other) if axis == 0 else (other, bound)

    prev = ring[-1]
    prev_in = inside(prev)
    for cur in ring:
        cur_in = inside(cur)
        if cur_in:
            if not prev_in:
                out.append(cross_pt(prev, cur))
            out.append(cur)
        elif prev_in:
            out.append(cross_pt(prev, cur))
        prev, prev_in = cur, cur_in
    return out


def _clip_to_band(px, py, xlo, xhi, ylo, yhi) -> np.ndarray:
    """Convex ring ∩ axis-aligned rectangle -> closed ring ``[V+1, 2]`` (first == last)."""
    ring = list(zip(map(float, px), map(float, py)))
    ring = _clip_halfplane(ring, 1, yhi, True)
    ring = _clip_halfplane(ring, 1, ylo, False)
    ring = _clip_halfplane(ring, 0, xhi, True)
    ring = _clip_halfplane(ring, 0, xlo, False)
    # drop consecutive duplicates produced by vertices lying on a cut line
    dedup: List[Tuple[float, float]] = []
    for p in ring:
        if not dedup or (abs(p[0] - dedup[-1][0]) > 0 or abs(p[1] - dedup[-1][1]) > 0):
            dedup.append(p)
    if len(dedup) > 1 and dedup[0] == dedup[-1]:
        dedup.pop()
    if len(dedup) < 3:
        return np.zeros((0, 2))
    dedup.append(dedup[0])
    return np.asarray(dedup, dtype=np.float64)


def _douglas_peucker(pts: np.ndarray, tol: float) -> np.ndarray:
    """Open-polyline simplification; end points are always kept."""
    n = len(pts)
    if n < 3:
        return pts.copy()
    keep = np.zeros(n, dtype=bool)
    keep[0] = keep[-1] = True
    stack = [(0, n - 1)]
    while stack:
        a, b = stack.pop()
        if b <= a + 1:
            continue
        pa, pb = pts[a], pts[b]
        seg = pb - pa
        seg_len2 = float(seg @ seg)
        inner = pts[a + 1:b]
        if seg_len2 == 0.0:
            d = np.hypot(inner[:, 0] - pa[0], inner[:, 1] - pa[1])
        else:
            tt = np.clip(((inner - pa) @ seg) / seg_len2, 0.0, 1.0)
            proj = pa + tt[:, None] * seg
            d = np.hypot(inner[:, 0] - proj[:, 0], inner[:, 1] - proj[:, 1])
        k = int(np.argmax(d))
        if d[k] > tol:
            keep[a + 1 + k] = True
            stack.append((a, a + 1 + k))
            stack.append((a + 1 + k, b))
    return pts[keep]


def _slice_convex_region(px, py, angle, num_target, half_extent):
    """Rotate a convex ring by ``angle``, cut it into horizontal bands, rotate back.

    Follows the slicing rule of couplers_coor.py:306-320 / 408-452 (fold coupler) and
    couplers_coor.py:460-475 / 557-600 (out-coupler).
    """
    rot = np.array([[np.cos(angle), np.sin(angle)],
                    [-np.sin(angle), np.cos(angle)]])
    inv = np.array([[np.cos(angle), -np.sin(angle)],
                    [np.sin(angle), np.cos(angle)]])
    rp = rot @ np.vstack((px, py))
    start_line = np.max(rp[1])
    end_line = np.min(rp[1])
    span = start_line - end_line
    slice_width = span / (num_target + 0.001)
    num_slices = int(np.ceil(span / slice_width))
    if (span % slice_width) < slice_width / 4:
        num_slices -= 1
    rings = []
    for i in range(1, num_slices + 1):
        col_start = start_line - (i - 1) * slice_width
        col_end = end_line if i == num_slices else start_line - i * slice_width
        ring = _clip_to_band(rp[0], rp[1], -half_extent, half_extent, col_end, col_start)
        if len(ring) == 0:
            continue
        rings.append((inv @ ring.T).T)
    return rings


# --------------------------------------------------------------------------
# k-space helpers (vectorised)
# --------------------------------------------------------------------------
def _air_angles(fx, fy):
    th = np.arctan(np.sqrt(np.tan(fx) ** 2 + np.tan(fy) ** 2))
    ph = np.arctan2(np.tan(fy), np.tan(fx))
    return th, ph


def _fc_quads(d: WaveguideDesign, fx, fy, k0, kg):
    """Tangent-line intersections of the in-coupled pupil with the out-coupler footprint.

    fx, fy: FoV angles, any shape S;  k0: scalar.  Returns x[S,4], y[S,4] in the corner order
    (b22,b11) (b21,b11) (b21,b12) (b22,b12) of couplers_coor.py:369-377, plus the per-FoV
    eyebox shift (dx, dy).
    """
    kgx_ic, kgy_ic, kgx_fc, kgy_fc = kg
    th, ph = _air_angles(fx, fy)
    kx0 = d.n_air * k0 * np.sin(th) * np.cos(ph)
    ky0 = d.n_air * k0 * np.sin(th) * np.sin(ph)
    kx_ic = kx0 + kgx_ic
    ky_ic = ky0 + kgy_ic
    k1 = ky_ic / kx_ic
    b11 = d.y_ic0 - k1 * d.x_ic0 + d.r * np.sqrt(1 + k1 ** 2)
    b12 = d.y_ic0 - k1 * d.x_ic0 - d.r * np.sqrt(1 + k1 ** 2)
    kx_fc = kx_ic + kgx_fc
    ky_fc = ky_ic + kgy_fc
    dx = d.er * np.tan(th) * np.cos(ph)
    dy = d.er * np.tan(th) * np.sin(ph)
    xl = d.x_eb0 - d.x_eb / 2 + dx
    xr = d.x_eb0 + d.x_eb / 2 + dx
    yb = d.y_eb0 - d.y_eb / 2 + dy
    yt = d.y_eb0 + d.y_eb / 2 + dy
    k2 = ky_fc / kx_fc
    neg = k2 <= 0
    b21 = np.where(neg, yb - k2 * xl, yt - k2 * xl)
    b22 = np.where(neg, yt - k2 * xr, yb - k2 * xr)
    den = k1 - k2
    xs = np.stack(((b22 - b11) / den, (b21 - b11) / den, (b21 - b12) / den, (b22 - b12) / den), axis=-1)
    bb = np.stack((b11, b11, b12, b12), axis=-1)
    ys = k1[..., None] * xs + bb
    return xs, ys, (kx0, ky0, kx_ic, ky_ic, kx_fc, ky_fc), (dx, dy)


def design_tables(d: WaveguideDesign, num_FOV_x: int, num_FOV_y: int):
    """Per-(wavelength, FoV) angle tables, ``lut_gap`` and ``lut_TIR`` (couplers_coor.py:614-711)."""
    FoV_x = d.fov_x_deg * deg
    FoV_y = FoV_x / d.aspect
    lmd = np.asarray(d.lmd, dtype=np.float64)
    k0 = (2 * np.pi / lmd)[:, None, None]
    phi_ic = d.phi_ic_deg * deg
    phi_oc = d.phi_oc_deg * deg
    kg_ic = 2 * np.pi / d.Lambda_ic
    kgx_ic, kgy_ic = kg_ic * np.cos(phi_ic), kg_ic * np.sin(phi_ic)
    kg_oc = 2 * np.pi / d.Lambda_oc
    kgx_oc, kgy_oc = kg_oc * np.cos(phi_oc + 180 * deg), kg_oc * np.sin(phi_oc + 180 * deg)
    kgx_fc, kgy_fc = kgx_oc - kgx_ic, kgy_oc - kgy_ic

    fxs = np.linspace(-FoV_x / 2, FoV_x / 2, num_FOV_x)
    fys = np.linspace(-FoV_y / 2, FoV_y / 2, num_FOV_y)
    FX, FY = np.meshgrid(fxs, fys, indexing="ij")
    th_in1, ph_in1 = _air_angles(FX, FY)
    L = len(lmd)
    th_in = np.broadcast_to(th_in1, (L,) + th_in1.shape).copy()
    ph_in = np.broadcast_to(ph_in1, (L,) + ph_in1.shape).copy()
    kx = d.n_air * k0 * np.sin(th_in) * np.cos(ph_in)
    ky = d.n_air * k0 * np.sin(th_in) * np.sin(ph_in)
    kn2 = (k0 * d.n_g) ** 2

    def direction(kxg, kyg):
        kz = np.sqrt(kn2 - kxg ** 2 - kyg ** 2)
        return np.arctan(np.sqrt((kxg ** 2 + kyg ** 2) / kz ** 2)), np.arctan2(kyg, kxg)

    th_ic2, ph_ic2 = direction(kx - kgx_ic, ky - kgy_ic)
    kxg_ic, kyg_ic = kx + kgx_ic, ky + kgy_ic
    th_ic, ph_ic = direction(kxg_ic, kyg_ic)
    kxg_fc, kyg_fc = kxg_ic + kgx_fc, kyg_ic + kgy_fc
    th_fc, ph_fc = direction(kxg_fc, kyg_fc)
    th_oc, ph_oc = direction(kxg_fc - 2 * kgx_oc, kyg_fc - 2 * kgy_oc)
    th_glow = np.arcsin(np.sin(th_in) / d.n_g)

    lut_gap = np.zeros((L, num_FOV_x, num_FOV_y, 8))
    for k, (th, ph) in enumerate(((th_ic, ph_ic), (th_fc, ph_fc), (th_ic2, ph_ic2), (th_oc, ph_oc))):
        lut_gap[..., 2 * k] = 2 * d.t * np.tan(th) * np.cos(ph)
        lut_gap[..., 2 * k + 1] = 2 * d.t * np.tan(th) * np.sin(ph)

    lut_TIR = np.zeros((L, num_FOV_x, num_FOV_y, 4))
    with np.errstate(invalid="ignore"):
        for k, th in enumerate((th_ic, th_fc, th_ic2, th_oc)):
            root = np.sqrt(d.n_g ** 2 * np.sin(th) ** 2 - 1)
            delta_s = 2 * np.arctan(root / (d.n_g * np.cos(th)))
            delta_p = 2 * np.arctan(d.n_g * root / np.cos(th))
            lut_TIR[..., k] = delta_s - delta_p

    # Fresnel table of couplers_coor.py:640-647 (returned, never used by the ray walk); the
    # reference overwrites it per wavelength, so the last wavelength's values survive.
    thg = th_glow[-1]
    thi = th_in[-1]
    lut_Fresnel = np.zeros((num_FOV_x, num_FOV_y, 4))
    lut_Fresnel[..., 0] = (d.n_g * np.cos(thg) - np.cos(thi)) / (d.n_g * np.cos(thg) + np.cos(thi))
    lut_Fresnel[..., 1] = (np.cos(thg) - d.n_g * np.cos(thi)) / (np.cos(thg) + d.n_g * np.cos(thi))
    lut_Fresnel[..., 2] = 2 * d.t * np.tan(thg) * np.cos(ph_in[-1])
    lut_Fresnel[..., 3] = 2 * d.t * np.tan(thg) * np.cos(ph_in[-1])

    angles = dict(th_in_ic=th_in, phi_in_ic=ph_in, th_out_ic=th_ic, phi_out_ic=ph_ic,
                  th_out_fc=th_fc, phi_out_fc=ph_fc, th_out_ic2=th_ic2, phi_out_ic2=ph_ic2,
                  th_out_oc=th_oc, phi_out_oc=ph_oc, th_out_oc_glow=th_glow)
    return lut_TIR, lut_gap, lut_Fresnel, angles, (FX, FY)


def couplers_coor_full_color(num_FOV_x: int = 120, num_FOV_y: int = 80, design: WaveguideDesign | None = None):
    """Same 37-entry tuple as the reference function (couplers_coor.py:740-750)."""
    d = design if design is not None else WaveguideDesign()
    FoV_x = d.fov_x_deg * deg
    FoV_y = FoV_x / d.aspect
    lmd = np.asarray(d.lmd)
    k0 = 2 * np.pi / lmd
    phi_ic = d.phi_ic_deg * deg
    phi_oc = d.phi_oc_deg * deg

    t_ic = np.linspace(0, 2 * np.pi, d.n_ic_pts)
    X_ic = d.x_ic0 + d.r * np.sin(t_ic)
    Y_ic = d.y_ic0 + d.r * np.cos(t_ic)

    x_oc = np.tan(FoV_x / 2) * abs(d.er) * 2 + d.x_eb
    y_oc = np.tan(FoV_y / 2) * abs(d.er) * 2 + d.y_eb
    X_oc = np.array([-x_oc / 2, -x_oc / 2, x_oc / 2, x_oc / 2]) + d.x_eb0
    Y_oc = np.array([-y_oc / 2, y_oc / 2, y_oc / 2, -y_oc / 2]) + d.y_eb0

    kg_ic = 2 * np.pi / d.Lambda_ic
    kgx_ic, kgy_ic = kg_ic * np.cos(phi_ic), kg_ic * np.sin(phi_ic)
    kg_oc = 2 * np.pi / d.Lambda_oc
    kgx_oc, kgy_oc = kg_oc * np.cos(phi_oc + 180 * deg), kg_oc * np.sin(phi_oc + 180 * deg)
    kgx_fc, kgy_fc = kgx_oc - kgx_ic, kgy_oc - kgy_ic
    Lambda_fc = 2 * np.pi / np.sqrt(kgx_fc ** 2 + kgy_fc ** 2)
    phi_fc = np.arctan2(kgy_fc, kgx_fc)
    kg = (kgx_ic, kgy_ic, kgx_fc, kgy_fc)

    # ---- layout sweep over the n_sweep x n_sweep FoV grid (couplers_coor.py:222-275) ----
    FoV_X = np.linspace(-FoV_x / 2, FoV_x / 2, d.n_sweep)
    FoV_Y = np.linspace(-FoV_y / 2, FoV_y / 2, d.n_sweep)
    SX, SY = np.meshgrid(FoV_X, FoV_Y, indexing="ij")
    n_cells = d.n_sweep * d.n_sweep
    kx0 = np.zeros((len(lmd), n_cells)); ky0 = np.zeros_like(kx0)
    kx_ic = np.zeros_like(kx0); ky_ic = np.zeros_like(kx0)
    kx_fc = np.zeros_like(kx0); ky_fc = np.zeros_like(kx0)
    xf_parts, yf_parts = [], []
    for li in range(len(lmd)):
        xs, ys, kk, _ = _fc_quads(d, SX.ravel(), SY.ravel(), k0[li], kg)
        kx0[li], ky0[li], kx_ic[li], ky_ic[li], kx_fc[li], ky_fc[li] = kk
        xf_parts.append(xs); yf_parts.append(ys)
    x_f = np.concatenate([p.ravel() for p in xf_parts])
    y_f = np.concatenate([p.ravel() for p in yf_parts])

    bd = _hull_indices(x_f, y_f)
    hull_fc_x, hull_fc_y = x_f[bd], y_f[bd]
    x_all = list(hull_fc_x); y_all = list(hull_fc_y)

    # ---- nine probe FoVs (couplers_coor.py:279-289, 328-377) ----
    eps = np.finfo(float).eps
    F9x = np.array([-FoV_x / 2, eps, FoV_x / 2, -FoV_x / 2, eps, FoV_x / 2, FoV_x / 2, eps, -FoV_x / 2])
    F9y = np.array([FoV_y / 2, FoV_y / 2, FoV_y / 2, eps, eps, eps, -FoV_y / 2, -FoV_y / 2, -FoV_y / 2])
    n9 = len(F9x)
    x_fc_FOV = np.zeros((n9 * len(lmd), 4)); y_fc_FOV = np.zeros_like(x_fc_FOV)
    for li in range(len(lmd)):
        xs, ys, _, _ = _fc_quads(d, F9x, F9y, k0[li], kg)
        x_fc_FOV[li::len(lmd)] = xs
        y_fc_FOV[li::len(lmd)] = ys

    # ---- effective region 2: in-coupler + fold-coupler footprints (couplers_coor.py:384-404) ----
    for i in range(n9 * len(lmd)):
        xc = np.hstack((x_fc_FOV[i], X_ic)); yc = np.hstack((y_fc_FOV[i], Y_ic))
        b = _hull_indices(xc, yc)
        x_all.extend(xc[b]); y_all.extend(yc[b])
    xa = np.array(x_all); ya = np.array(y_all)
    b = _hull_indices(xa, ya)
    eff_reg2 = _douglas_peucker(np.column_stack((xa[b], ya[b])), d.simplify_tol)

    # ---- fold-coupler slices (couplers_coor.py:306-320, 408-452) ----
    fc_rings = _slice_convex_region(hull_fc_x, hull_fc_y, np.pi / 2 + phi_ic, d.num_FC, d.plate_x)

    # ---- out-coupler slices (couplers_coor.py:455-475, 557-600) ----
    b = _hull_indices(X_oc, Y_oc)
    oc_rings = _slice_convex_region(X_oc[b], Y_oc[b], 3 * np.pi / 2 + phi_oc, d.num_OC, d.plate_x)

    # ---- eyebox footprints of the nine probes (couplers_coor.py:478-499) ----
    th9, ph9 = _air_angles(F9x, F9y)
    dx9 = d.er * np.tan(th9) * np.cos(ph9); dy9 = d.er * np.tan(th9) * np.sin(ph9)
    xl, xr = d.x_eb0 - d.x_eb / 2, d.x_eb0 + d.x_eb / 2
    yb, yt = d.y_eb0 - d.y_eb / 2, d.y_eb0 + d.y_eb / 2
    x_oc_FOV = np.stack((xl + dx9, xl + dx9, xr + dx9, xr + dx9), axis=-1)
    y_oc_FOV = np.stack((yt + dy9, yb + dy9, yb + dy9, yt + dy9), axis=-1)

    # ---- effective region 1: everything (couplers_coor.py:538-554) ----
    for i in range(n9):
        for li in range(len(lmd)):
            ex = np.concatenate([x_oc_FOV[i], x_fc_FOV[i * len(lmd) + li]])
            ey = np.concatenate([y_oc_FOV[i], y_fc_FOV[i * len(lmd) + li]])
            b = _hull_indices(ex, ey)
            x_all.extend(ex[b]); y_all.extend(ey[b])
    xa = np.array(x_all); ya = np.array(y_all)
    b = _hull_indices(xa, ya)
    eff_reg1 = _douglas_peucker(np.column_stack((xa[b], ya[b])), d.simplify_tol)

    # ---- per-FoV tables (couplers_coor.py:502-532, 614-711) ----
    lut_TIR, lut_gap, lut_Fresnel, ang, (FX, FY) = design_tables(d, num_FOV_x, num_FOV_y)
    thg, phg = _air_angles(FX, FY)
    dxg = d.er * np.tan(thg) * np.cos(phg); dyg = d.er * np.tan(thg) * np.sin(phg)
    rect_x = np.stack((xl + dxg, xl + dxg, xr + dxg, xr + dxg), axis=-1)
    rect_y = np.stack((yt + dyg, yb + dyg, yb + dyg, yt + dyg), axis=-1)
    eff_reg_FOV = np.stack((rect_x, rect_y), axis=-1)
    eff_reg_FOV_range = np.stack((xl + dxg, xr + dxg, yb + dyg, yt + dyg), axis=-1)

    IC = np.stack((X_ic, Y_ic), axis=1)
    FC = np.concatenate(fc_rings, axis=0)
    FC_offset = np.cumsum([0] + [len(r) for r in fc_rings])
    OC = np.concatenate(oc_rings, axis=0)
    OC_offset = np.cumsum([0] + [len(r) for r in oc_rings])

    return (IC,
            FC, FC_offset,
            OC, OC_offset,
            eff_reg1,
            eff_reg2,
            eff_reg_FOV, eff_reg_FOV_range,
            lut_TIR, lut_gap, lut_Fresnel,
            d.Lambda_ic, phi_ic, Lambda_fc, phi_fc, d.Lambda_oc, phi_oc, d.n_g, lmd,
            ang["th_in_ic"], ang["phi_in_ic"], ang["th_out_ic"], ang["phi_out_ic"],
            ang["th_out_fc"], ang["phi_out_fc"],
            ang["th_out_ic2"], ang["phi_out_ic2"], ang["th_out_oc"], ang["phi_out_oc"],
            ang["th_out_oc_glow"],
            kx0, ky0, kx_ic, ky_ic, kx_fc, ky_fc)
