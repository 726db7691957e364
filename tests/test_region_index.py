"""The cell-grid region index must answer exactly like the reference's literal polygon scan --
for ordinary points and for points engineered to sit on / next to edges and vertices."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from gpu_ray_tracing_for_waveguide_based_ar_display_b200 import _capi
from gpu_ray_tracing_for_waveguide_based_ar_display_b200.couplers_coor import (
    WaveguideDesign, couplers_coor_full_color)


def locate(verts, off, px, py, mode):
    lib = _capi.load_library()
    verts = np.ascontiguousarray(verts, dtype=np.float64); off = np.ascontiguousarray(off, dtype=np.int64)
    px = np.ascontiguousarray(px, dtype=np.float64); py = np.ascontiguousarray(py, dtype=np.float64)
    out = np.zeros(len(px), dtype=np.int32)
    _capi.check(lib.wgrt_debug_locate(verts.ctypes.data, len(verts), off.ctypes.data, len(off) - 1,
                                      px.ctypes.data, py.ctypes.data, len(px), out.ctypes.data, mode), lib)
    return out


def adversarial_points(verts, rs, n_random):
    lo = verts.min(0) - 0.05 * np.ptp(verts, 0) - 0.5
    hi = verts.max(0) + 0.05 * np.ptp(verts, 0) + 0.5
    pts = [rs.uniform(lo, hi, size=(n_random, 2))]
    prev = np.roll(verts, 1, axis=0)
    for t in (0.0, 0.5, 1.0, 0.123456789, 1e-9, 1 - 1e-9):
        on = prev + t * (verts - prev)
        nrm = np.stack((-(verts - prev)[:, 1], (verts - prev)[:, 0]), 1)
        nrm /= np.maximum(np.hypot(nrm[:, 0], nrm[:, 1]), 1e-300)[:, None]
        for eps in (0.0, 3e-13, -3e-13, 9e-13, -9e-13, 2e-12, -2e-12, 1e-9, -1e-9, 1e-6, -1e-6):
            pts.append(on + eps * nrm)
    # points sharing a y (or x) coordinate with a vertex: exercises the half-open straddle rule
    k = rs.integers(0, len(verts), 400)
    pts.append(np.stack((rs.uniform(lo[0], hi[0], 400), verts[k, 1]), 1))
    pts.append(np.stack((verts[k, 0], rs.uniform(lo[1], hi[1], 400)), 1))
    pts.append(np.array([[np.nan, 0.0], [0.0, np.nan], [np.inf, 0.0], [-np.inf, np.inf], [1e300, -1e300]]))
    return np.concatenate(pts)


@pytest.mark.parametrize("design", [None, WaveguideDesign(t=0.3, num_FC=15, fov_x_deg=24.0)])
def test_grid_index_equals_literal_scan(design, oracle):
    out = couplers_coor_full_color(5, 5, design=design)
    IC, FC, FC_off, OC, OC_off, r1, r2 = out[:7]
    rs = np.random.default_rng(5)
    for name, verts, off in (("IC", IC, [0, len(IC)]), ("FC", FC, FC_off), ("OC", OC, OC_off),
                             ("eff_reg1", r1, [0, len(r1)]), ("eff_reg2", r2, [0, len(r2)])):
        pts = adversarial_points(verts, rs, 200000)
        want = oracle.locate(verts, off, pts[:, 0], pts[:, 1])
        lit = locate(verts, off, pts[:, 0], pts[:, 1], 0)
        grid = locate(verts, off, pts[:, 0], pts[:, 1], 1)
        assert np.array_equal(lit, want), f"{name}: literal GPU scan vs oracle"
        bad = np.flatnonzero(grid != want)
        assert bad.size == 0, f"{name}: grid index differs at {pts[bad[:5]]} got {grid[bad[:5]]} want {want[bad[:5]]}"
        atlas = locate(verts, off, pts[:, 0], pts[:, 1], 2)     # the warp walk's path: atlas word, then the grids
        bad = np.flatnonzero(atlas != want)
        assert bad.size == 0, f"{name}: atlas differs at {pts[bad[:5]]} got {atlas[bad[:5]]} want {want[bad[:5]]}"
        zone = locate(verts, off, pts[:, 0], pts[:, 1], 3)      # the production walk's path: zone id -> word, then the grids
        bad = np.flatnonzero(zone != want)
        assert bad.size == 0, f"{name}: zone grids differ at {pts[bad[:5]]} got {zone[bad[:5]]} want {want[bad[:5]]}"
        assert (want >= 0).sum() > 1000 and (want < 0).sum() > 1000


def test_grid_index_odd_ring_sets(oracle):
    """Non-convex ring, overlapping rings (first-hit order matters), empty ring, degenerate ring."""
    star_t = np.linspace(0, 2 * np.pi, 11)[:-1]
    star = np.stack((np.cos(star_t), np.sin(star_t)), 1) * np.where(np.arange(10) % 2, 0.4, 1.0)[:, None]
    sq = np.array([[-0.5, -0.5], [0.5, -0.5], [0.5, 0.5], [-0.5, 0.5], [-0.5, -0.5]])
    line = np.array([[2.0, 2.0], [3.0, 3.0]])
    verts = np.concatenate((star, sq, sq + 0.25, line))
    off = np.array([0, 10, 10, 15, 20, 22])          # ring 1 is empty
    rs = np.random.default_rng(9)
    pts = adversarial_points(verts, rs, 100000)
    want = oracle.locate(verts, off, pts[:, 0], pts[:, 1])
    assert np.array_equal(locate(verts, off, pts[:, 0], pts[:, 1], 1), want)
    assert np.array_equal(locate(verts, off, pts[:, 0], pts[:, 1], 2), want)
    assert np.array_equal(locate(verts, off, pts[:, 0], pts[:, 1], 3), want)
    assert set(np.unique(want)) >= {-1, 0, 2, 3}


def test_many_slices_overflow_the_zone_table(oracle):
    """A ring set with 250 slices in a checkerboard has more distinct atlas words than the zone table holds
    entries for the finer cells; whatever form the index takes, lookups must equal the literal scan."""
    n = 250
    xs = np.arange(n) % 25
    ys = np.arange(n) // 25
    sq = np.array([[0.0, 0.0], [0.9, 0.0], [0.9, 0.9], [0.0, 0.9]])
    verts = np.concatenate([sq + np.array([x, y], dtype=np.float64) for x, y in zip(xs, ys)])
    off = np.arange(0, 4 * n + 1, 4)
    rs = np.random.default_rng(11)
    pts = adversarial_points(verts, rs, 50000)
    want = oracle.locate(verts, off, pts[:, 0], pts[:, 1])
    for mode in (1, 2, 3):
        assert np.array_equal(locate(verts, off, pts[:, 0], pts[:, 1], mode), want), mode
    assert len(np.unique(want)) == n + 1
