"""Property tests of the CPU oracle's building blocks against independent NumPy / pure-Python restatements
of the reference expressions (GPU_ray_tracing_functions.py:25-71, 124-152).  The golden vectors pin the
oracle to outputs of the reference itself; these tests widen the input space (hypothesis)."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def _xorshift_py(s, idx):
    # GRTF:25-34
    if s == 0:
        s = (0x6D2B79F5 ^ (idx + 1)) & 0xFFFFFFFF
    s ^= (s << 13) & 0xFFFFFFFF
    s ^= s >> 17
    s ^= (s << 5) & 0xFFFFFFFF
    return s, s * (1.0 / 4294967296.0)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 2 ** 32 - 1), min_size=1, max_size=50), st.integers(1, 40))
def test_xorshift_matches_python(oracle, states, draws):
    got_s, got_u = oracle.xorshift(np.array(states, dtype=np.uint32), draws)
    for i, s in enumerate(states):
        u = None
        for _ in range(draws):
            s, u = _xorshift_py(s, i)
        assert int(got_s[i]) == s and got_u[i] == u
        assert 0.0 < got_u[i] < 1.0


def _inside_or_on_edge_np(px, py, poly):
    """GRTF:36-71 restated with NumPy scalars, same expressions, same order."""
    n = len(poly)
    tol = 1e-12
    j = n - 1
    for i in range(n):
        x1, y1 = poly[j]; x2, y2 = poly[i]
        if not (px < min(x1, x2) - tol or px > max(x1, x2) + tol or py < min(y1, y2) - tol or py > max(y1, y2) + tol):
            if abs((x2 - x1) * (py - y1) - (y2 - y1) * (px - x1)) <= tol:
                return True
        j = i
    inside = False
    j = n - 1
    for i in range(n):
        xi, yi = poly[i]; xj, yj = poly[j]
        if (yi > py) != (yj > py):
            if px < (xj - xi) * (py - yi) / (yj - yi + 1e-20) + xi:
                inside = not inside
        j = i
    return inside


@settings(max_examples=40, deadline=None)
@given(st.integers(3, 12), st.integers(0, 2 ** 31 - 1))
def test_locate_matches_literal_restatement(oracle, nverts, seed):
    rs = np.random.default_rng(seed)
    # two rings: a random star-shaped polygon and a shifted copy (first-hit order matters where they overlap)
    ang = np.sort(rs.uniform(0, 2 * np.pi, nverts))
    rad = rs.uniform(0.3, 1.0, nverts)
    ring = np.stack((rad * np.cos(ang), rad * np.sin(ang)), 1)
    verts = np.concatenate((ring, ring + rs.uniform(-0.4, 0.4, 2)))
    off = np.array([0, nverts, 2 * nverts])
    pts = rs.uniform(-1.5, 1.5, (200, 2))
    k = rs.integers(0, len(verts), 40)
    pts = np.concatenate((pts, verts[k], 0.5 * (verts[k] + np.roll(verts, 1, 0)[k])))     # vertices, near-edge points
    got = oracle.locate(verts, off, pts[:, 0], pts[:, 1])
    for (x, y), g in zip(pts, got):
        want = -1
        for r in range(2):
            if _inside_or_on_edge_np(np.float64(x), np.float64(y), verts[off[r]:off[r + 1]]):
                want = r
                break
        assert g == want, (x, y, g, want)


@settings(max_examples=60, deadline=None)
@given(st.floats(0, 1), st.floats(0, 1), st.floats(-math.pi, math.pi), st.integers(0, 2 ** 31 - 1))
def test_efield_matches_complex_arithmetic(oracle, ete, etm, delta, seed):
    """E_field_cal (GRTF:132-152) against the same formula in NumPy complex128: amplitudes to 1e-13, phase
    difference modulo 2 pi to 1e-9 where both amplitudes are well above the 1e-20 rule."""
    rs = np.random.default_rng(seed)
    J = (rs.normal(size=4) + 1j * rs.normal(size=4)).astype(np.complex128)      # call order: te_te, te_tm, tm_te, tm_tm
    out = oracle.efield([ete], [etm], [delta], J[None, :])[0]
    te_in = complex(ete, 0.0)
    tm_in = etm * complex(math.cos(delta), math.sin(delta))
    e_te = J[0] * te_in + J[2] * tm_in
    e_tm = J[1] * te_in + J[3] * tm_in
    assert out[0] == pytest.approx(abs(e_te), rel=1e-13, abs=1e-300)
    assert out[1] == pytest.approx(abs(e_tm), rel=1e-13, abs=1e-300)
    assert -math.pi - 1e-12 <= out[2] <= math.pi + 1e-12
    if abs(e_te) > 1e-6 and abs(e_tm) > 1e-6:
        want = np.angle(e_tm) - np.angle(e_te)
        d = (out[2] - want + math.pi) % (2 * math.pi) - math.pi
        assert abs(d) < 1e-9 or abs(abs(d) - 2 * math.pi) < 1e-9
