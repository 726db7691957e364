// wgrt_region.cuh -- exact-equivalent acceleration of the reference's point-in-region tests.
//
// The reference answers "which coupler polygon contains (x, y)?" by running
// is_inside_or_on_edge (GRTF:63-71 -> GRTF:52-61 and GRTF:36-50) over every ring in order and
// every edge of each ring, twice.  That scan is ~92 % of its arithmetic.
// Here each region set (in-coupler, effective regions 1 and 2, fold-coupler slices, out-coupler
// slices) gets a two-level uniform grid over its bounding box.  The coarse level (64 x 64 bytes,
// resident in L1) settles every query that falls in a coarse cell no edge comes near; coarse cells
// marked MIXED defer to the fine level (4096 x 4096 by default, populated only under MIXED coarse
// cells).  Cell codes, on both levels:
//
//   cell code  0..253 : every point that maps to this cell is inside ring <code> and in no
//                       earlier ring -> the reference's first-hit index, no arithmetic
//   CELL_NONE         : every such point is outside all rings
//   CELL_AMBIG        : an edge of a not-yet-excluded ring passes within the safety margin of the
//                       cell -> run the reference's own expressions, but only on the edges whose
//                       y-range meets this cell row (a per-row edge bit mask).  Edges outside that
//                       mask cannot satisfy (yi>y)!=(yj>y) nor the tolerance box of
//                       point_on_segment, so skipping them cannot change either pass.  A detail
//                       word per cell narrows the scan further: rings before `first` are certainly
//                       missed, rings [first, stop) are tested exactly in order, and if none
//                       contains the point the answer is `dflt` (the first ring certainly hit after
//                       them, or none).
//
// Booleans are therefore identical to the literal scan for every input point; only the amount of
// work differs.  tests/test_region_index.py checks this against the oracle on adversarial points
// (vertices, edge midpoints, points 5e-13 and 5e-12 off an edge).
#pragma once

#include "wgrt_device.cuh"

namespace wgrt {

// Region descriptor as the walk kernels see it (shared memory copy: static + device-computed part).
struct alignas(16) Region {
  double x0, y0, inv_dx, inv_dy;   // fine-cell coordinates: fx = (x - x0) * inv_dx in [0, n)
  const uint8_t* coarse;
  const uint8_t* cells;
  const uint32_t* detail;
  const uint32_t* rowmask;
  const double* verts;
  const int64_t* offsets;
  int nverts, npoly, n, nc, shift, words;
  int pad_[2];
};

__device__ __forceinline__ void region_load(Region& r, const RegionStatic& st, const RegionDyn& dy) {
  r.x0 = dy.x0; r.y0 = dy.y0; r.inv_dx = dy.inv_dx; r.inv_dy = dy.inv_dy;
  r.coarse = st.coarse; r.cells = st.cells; r.detail = st.detail; r.rowmask = st.rowmask; r.verts = st.verts;
  r.offsets = st.offsets;
  r.nverts = st.nverts; r.npoly = st.npoly; r.n = st.n; r.nc = st.nc; r.shift = st.shift; r.words = st.words;
}

__device__ __forceinline__ int ring_begin(const int64_t* off, int nverts, int k) {
  return off ? static_cast<int>(off[k]) : (k == 0 ? 0 : nverts);
}

// The reference's two passes over ring [s, e), restricted to edges whose bit is set in `mask`
// (bit i <-> the edge ending at vertex i, i.e. (prev(i) -> i) exactly as GRTF:40-49 / 66-70 pair
// them).  Pass 1's early `return True` and pass 2's parity commute with skipping, so both passes
// are fused into one sweep: result = any(on_segment) or odd(crossings).
template <bool COUNT>
__device__ __forceinline__ bool ring_test_masked(double px, double py, const double* __restrict__ verts, int s,
                                                 int e, const uint32_t* __restrict__ mask, Counts* cn) {
  if (e <= s) return false;
  bool on_edge = false, inside = false;
  const int w0 = s >> 5, w1 = (e - 1) >> 5;
  for (int w = w0; w <= w1; ++w) {
    uint32_t bits = __ldg(mask + w);
    if (w == w0) bits &= 0xffffffffu << (s & 31);
    if (w == w1 && ((e & 31) != 0)) bits &= 0xffffffffu >> (32 - (e & 31));
    while (bits) {
      const int i = (w << 5) + __ffs(bits) - 1;
      bits &= bits - 1;
      const int j = (i == s) ? e - 1 : i - 1;
      const double2 vi = __ldg(reinterpret_cast<const double2*>(verts) + i);
      const double2 vj = __ldg(reinterpret_cast<const double2*>(verts) + j);
      // GRTF:68 calls point_on_segment(px, py, poly, start + j, start + i)
      on_edge |= on_segment_literal<COUNT>(px, py, vj.x, vj.y, vi.x, vi.y, 1e-12, cn);
      if ((vi.y > py) != (vj.y > py)) {
        if (COUNT) cn->c[WGRT_CNT_STRADDLE]++;
        if (px < (vj.x - vi.x) * (py - vi.y) / (vj.y - vi.y + 1e-20) + vi.x) inside = !inside;
      }
    }
  }
  return on_edge || inside;
}

// detail word of an ambiguous cell: first | stop << 8 | dflt << 16  (dflt 255 = none)
__host__ __device__ __forceinline__ uint32_t pack_detail(int first, int stop, int dflt) {
  return static_cast<uint32_t>(first) | (static_cast<uint32_t>(stop) << 8) | (static_cast<uint32_t>(dflt) << 16);
}

template <bool COUNT>
__device__ __noinline__ int region_locate_exact(const Region& r, double x, double y, int cell, int iy, Counts* cn) {
  if (COUNT) cn->c[WGRT_CNT_EXACT_FALLBACK]++;
  const uint32_t d = __ldg(r.detail + cell);
  const int first = d & 0xff, stop = (d >> 8) & 0xff, dflt = (d >> 16) & 0xff;
  const uint32_t* mask = r.rowmask + static_cast<size_t>(iy) * r.words;
  for (int k = first; k < stop; ++k) {
    const int s = ring_begin(r.offsets, r.nverts, k), e = ring_begin(r.offsets, r.nverts, k + 1);
    if (ring_test_masked<COUNT>(x, y, r.verts, s, e, mask, cn)) return k;
  }
  return dflt == 255 ? -1 : dflt;
}

// Index of the first ring of the set containing (x, y) -- what the reference's
// `for i in range(len(offset)-1): if is_inside_or_on_edge(...): ... break` finds -- or -1.
template <bool COUNT>
__device__ __forceinline__ int region_locate(const Region& r, double x, double y, Counts* cn) {
  if (COUNT) cn->c[WGRT_CNT_POLY_TESTS]++;
  const double fx = (x - r.x0) * r.inv_dx;
  const double fy = (y - r.y0) * r.inv_dy;
  // also rejects NaN coordinates, which the literal test classifies as outside
  const double lim = static_cast<double>(r.n);
  if (!(fx >= 0.0 && fy >= 0.0 && fx < lim && fy < lim)) return -1;
  const int ix = static_cast<int>(fx), iy = static_cast<int>(fy);
  uint8_t code = __ldg(r.coarse + (iy >> r.shift) * r.nc + (ix >> r.shift));
  if (code == CELL_AMBIG) {  // MIXED coarse cell: ask the fine level
    const int cell = iy * r.n + ix;
    code = __ldg(r.cells + cell);
    if (code == CELL_AMBIG) return region_locate_exact<COUNT>(r, x, y, cell, iy, cn);
  }
  return code == CELL_NONE ? -1 : code;
}

// Several region queries for the same point, with the memory accesses of all of them in flight
// together: first every coarse byte, then every needed fine byte, then (rarely) the exact scans.
// `want[k]` switches query k off (result -1).  Same answers as N calls of region_locate.
template <bool COUNT, int N>
__device__ __forceinline__ void region_locate_multi(const Region* const (&regs)[N], const bool (&want)[N], double x,
                                                    double y, int (&hit)[N], Counts* cn) {
  int cell[N], row[N];
  uint8_t code[N];
  bool ok[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const Region& r = *regs[k];
    const double fx = (x - r.x0) * r.inv_dx, fy = (y - r.y0) * r.inv_dy, lim = static_cast<double>(r.n);
    ok[k] = want[k] && fx >= 0.0 && fy >= 0.0 && fx < lim && fy < lim;
    const int ix = ok[k] ? static_cast<int>(fx) : 0, iy = ok[k] ? static_cast<int>(fy) : 0;
    row[k] = iy;
    cell[k] = iy * r.n + ix;
    code[k] = CELL_NONE;
    if (ok[k]) code[k] = __ldg(r.coarse + (iy >> r.shift) * r.nc + (ix >> r.shift));
    if (COUNT && want[k]) cn->c[WGRT_CNT_POLY_TESTS]++;
  }
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (code[k] == CELL_AMBIG) code[k] = __ldg(regs[k]->cells + cell[k]);  // MIXED coarse cell: fine level
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (code[k] == CELL_AMBIG) hit[k] = region_locate_exact<COUNT>(*regs[k], x, y, cell[k], row[k], cn);
    else hit[k] = code[k] == CELL_NONE ? -1 : code[k];
  }
}

}  // namespace wgrt
