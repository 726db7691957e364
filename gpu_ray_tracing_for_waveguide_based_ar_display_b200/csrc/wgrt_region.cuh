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
      WGRT_CHECK(i >= s && i < e && j >= s && j < e);
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
  WGRT_CHECK(cell >= 0 && cell < r.n * r.n && iy >= 0 && iy < r.n);
  const uint32_t d = __ldg(r.detail + cell);
  const int first = d & 0xff, stop = (d >> 8) & 0xff, dflt = (d >> 16) & 0xff;
  WGRT_CHECK(first <= stop && stop <= r.npoly && (dflt == 255 || dflt < r.npoly));
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
  WGRT_CHECK(ix >= 0 && ix < r.n && iy >= 0 && iy < r.n && (iy >> r.shift) < r.nc && (ix >> r.shift) < r.nc);
  uint8_t code = __ldg(r.coarse + (iy >> r.shift) * r.nc + (ix >> r.shift));
  if (code == CELL_AMBIG) {  // MIXED coarse cell: ask the fine level
    const int cell = iy * r.n + ix;
    code = __ldg(r.cells + cell);
    if (code == CELL_AMBIG) return region_locate_exact<COUNT>(r, x, y, cell, iy, cn);
  }
  return code == CELL_NONE ? -1 : code;
}

// region_locate for callers that expect the fine level to decide (the walk's rare path: the atlas already said
// "an edge is near"): the coarse byte, the fine byte and the detail word are requested together instead of one
// after the other -- `cells` and `detail` are allocated for every fine cell, only meaningful under MIXED coarse
// cells, and only used there.  Same answers as region_locate.
template <bool COUNT>
__device__ __forceinline__ int region_locate_eager(const Region& r, double x, double y, Counts* cn) {
  if (COUNT) cn->c[WGRT_CNT_POLY_TESTS]++;
  const double fx = (x - r.x0) * r.inv_dx;
  const double fy = (y - r.y0) * r.inv_dy;
  const double lim = static_cast<double>(r.n);
  if (!(fx >= 0.0 && fy >= 0.0 && fx < lim && fy < lim)) return -1;
  const int ix = static_cast<int>(fx), iy = static_cast<int>(fy);
  const int cell = iy * r.n + ix;
  const uint8_t coarse = __ldg(r.coarse + (iy >> r.shift) * r.nc + (ix >> r.shift));
  const uint8_t fine = __ldg(r.cells + cell);
  const uint32_t d = __ldg(r.detail + cell);
  const uint32_t mask0 = __ldg(r.rowmask + static_cast<size_t>(iy) * r.words);   // (warms the line the scan reads)
  if (coarse != CELL_AMBIG) return coarse == CELL_NONE ? -1 : coarse;
  if (fine != CELL_AMBIG) return fine == CELL_NONE ? -1 : fine;
  if (COUNT) cn->c[WGRT_CNT_EXACT_FALLBACK]++;
  (void)mask0;
  const int first = d & 0xff, stop = (d >> 8) & 0xff, dflt = (d >> 16) & 0xff;
  const uint32_t* mask = r.rowmask + static_cast<size_t>(iy) * r.words;
  for (int k = first; k < stop; ++k) {
    const int s = ring_begin(r.offsets, r.nverts, k), e = ring_begin(r.offsets, r.nverts, k + 1);
    if (ring_test_masked<COUNT>(x, y, r.verts, s, e, mask, cn)) return k;
  }
  return dflt == 255 ? -1 : dflt;
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

// ---------------------------------------------------------------------------------------------
// Atlas: all five region answers for a point from one word.
//   bits 0-1 in-coupler, 2-3 effective region 1, 4-5 effective region 2: 0 outside, 1 inside,
//            2 MIXED (an edge of the ring comes near this atlas cell)
//   bits 8-15 fold-coupler slice, 16-23 out-coupler slice: first-hit ring index 0..253,
//            CELL_NONE (no slice), CELL_AMBIG (MIXED)
//   bit 31   some field of the word is MIXED
// Two levels over the union of the sets' bounding boxes: 64 x 64 words (16 KB, L1 resident) and,
// under level-1 cells with a MIXED field, 64 x 64 finer words each (4096 x 4096 overall, the touched
// part L2 resident) in the same format.  A field still MIXED at level 2 defers to that set's own two-level
// grid (region_locate above), which ends in the reference's literal edge expressions; every other
// field is certain for every point that maps to the cell (same safety margin as the per-set grids).
// Points outside the atlas are outside every ring's bounding box.
// ---------------------------------------------------------------------------------------------
struct alignas(16) Atlas {
  double x0, y0, inv_dx, inv_dy;   // level-1 cell coordinates
  const uint32_t* words;           // level 1
  const uint32_t* words2;          // level 2
};
constexpr uint32_t ATLAS_OUTSIDE = (static_cast<uint32_t>(CELL_NONE) << 8) | (static_cast<uint32_t>(CELL_NONE) << 16);
constexpr uint32_t ATLAS_ANY_MIXED = 1u << 31;
enum { ATLAS_SHIFT_IC = 0, ATLAS_SHIFT_R1 = 2, ATLAS_SHIFT_R2 = 4, ATLAS_SHIFT_FC = 8, ATLAS_SHIFT_OC = 16 };

__device__ __forceinline__ uint32_t atlas_lookup(const Atlas& a, double x, double y) {
  const double fx = (x - a.x0) * a.inv_dx, fy = (y - a.y0) * a.inv_dy;
  const double lim = static_cast<double>(ATLAS_N);
  if (!(fx >= 0.0 && fy >= 0.0 && fx < lim && fy < lim)) return ATLAS_OUTSIDE;   // also NaN
  WGRT_CHECK(static_cast<int>(fx) >= 0 && static_cast<int>(fx) < ATLAS_N && static_cast<int>(fy) >= 0 && static_cast<int>(fy) < ATLAS_N);
  uint32_t word = __ldg(a.words + static_cast<int>(fy) * ATLAS_N + static_cast<int>(fx));
  if (word & ATLAS_ANY_MIXED) {
    const double sub = static_cast<double>(1 << ATLAS_SUB_SHIFT);
    const int ix = min(static_cast<int>(fx * sub), ATLAS_N2 - 1), iy = min(static_cast<int>(fy * sub), ATLAS_N2 - 1);
    WGRT_CHECK(ix >= 0 && iy >= 0 && (ix >> ATLAS_SUB_SHIFT) == static_cast<int>(fx) && (iy >> ATLAS_SUB_SHIFT) == static_cast<int>(fy));
    word = __ldg(a.words2 + static_cast<size_t>(iy) * ATLAS_N2 + ix);
  }
  return word;
}

// zone atlas as the walk reads it (shared memory copy)
struct alignas(16) ZoneAtlas {
  double x0, y0, inv_dx, inv_dy;     // level-1 zone cells
  const uint16_t* level1;
  const uint16_t* level2;
  const uint32_t* trans;
  const uint32_t* words;
  int outside_zone, valid;
};
__device__ __forceinline__ int zone_lookup(const ZoneAtlas& a, double x, double y) {
  const double fx = (x - a.x0) * a.inv_dx, fy = (y - a.y0) * a.inv_dy;
  const double lim = static_cast<double>(ZONE_N1);
  if (!(fx >= 0.0 && fy >= 0.0 && fx < lim && fy < lim)) return a.outside_zone;   // also NaN
  int z = __ldg(a.level1 + static_cast<int>(fy) * ZONE_N1 + static_cast<int>(fx));
  if (z == ZONE_MIXED) {
    const double sub = static_cast<double>(1 << ZONE_SUB_SHIFT);
    const int ix = min(static_cast<int>(fx * sub), ATLAS_N2 - 1), iy = min(static_cast<int>(fy * sub), ATLAS_N2 - 1);
    z = __ldg(a.level2 + static_cast<size_t>(iy) * ATLAS_N2 + ix);
  }
  return z;
}

// single-ring sets (in-coupler, effective regions): inside?
template <bool COUNT>
__device__ __forceinline__ bool atlas_inside(uint32_t word, int shift, const Region& r, double x, double y, Counts* cn) {
  const uint32_t c = (word >> shift) & 3u;
  if (c == 2u) return region_locate<COUNT>(r, x, y, cn) >= 0;
  if (COUNT) cn->c[WGRT_CNT_POLY_TESTS]++;
  return c == 1u;
}

// multi-ring sets (coupler slices): first-hit ring index or -1
template <bool COUNT>
__device__ __forceinline__ int atlas_hit(uint32_t word, int shift, const Region& r, double x, double y, Counts* cn) {
  const uint32_t c = (word >> shift) & 0xffu;
  if (c == CELL_AMBIG) return region_locate<COUNT>(r, x, y, cn);
  if (COUNT) cn->c[WGRT_CNT_POLY_TESTS]++;
  return c == CELL_NONE ? -1 : static_cast<int>(c);
}

// Make the fields of `word` named in `need` (bit r = region set r) certain: every needed field that
// is still MIXED after the two atlas levels is answered by that set's own grid / literal scan and
// patched into the word.  One call site for all sets: the rare path is compiled once.
template <bool COUNT>
__device__ __noinline__ uint32_t atlas_resolve(uint32_t word, uint32_t need, const Region* __restrict__ regions, double x,
                                               double y, Counts* cn) {
  // fields that are needed AND still MIXED (usually exactly one)
  uint32_t todo = need & ((((word >> ATLAS_SHIFT_IC) & 3u) == 2u ? 1u << REG_IC : 0u) |
                          (((word >> ATLAS_SHIFT_R1) & 3u) == 2u ? 1u << REG_R1 : 0u) |
                          (((word >> ATLAS_SHIFT_R2) & 3u) == 2u ? 1u << REG_R2 : 0u) |
                          (((word >> ATLAS_SHIFT_FC) & 0xffu) == CELL_AMBIG ? 1u << REG_FC : 0u) |
                          (((word >> ATLAS_SHIFT_OC) & 0xffu) == CELL_AMBIG ? 1u << REG_OC : 0u));
  // field position of set r: REG_IC 0, REG_R1 2, REG_R2 4, REG_FC 8, REG_OC 16, one byte each in a constant
  constexpr unsigned long long kShifts =
      (static_cast<unsigned long long>(ATLAS_SHIFT_IC) << (8 * REG_IC)) | (static_cast<unsigned long long>(ATLAS_SHIFT_R1) << (8 * REG_R1)) |
      (static_cast<unsigned long long>(ATLAS_SHIFT_R2) << (8 * REG_R2)) | (static_cast<unsigned long long>(ATLAS_SHIFT_FC) << (8 * REG_FC)) |
      (static_cast<unsigned long long>(ATLAS_SHIFT_OC) << (8 * REG_OC));
  while (todo) {
    const int r = __ffs(todo) - 1;
    todo &= todo - 1;
    const bool multi = r == REG_FC || r == REG_OC;
    const int shift = static_cast<int>((kShifts >> (8 * r)) & 0xffu);
    const uint32_t mask = multi ? 0xffu : 3u;
    const int hit = region_locate_eager<COUNT>(regions[r], x, y, cn);
    const uint32_t v = multi ? (hit < 0 ? static_cast<uint32_t>(CELL_NONE) : static_cast<uint32_t>(hit)) : (hit >= 0 ? 1u : 0u);
    word = (word & ~(mask << shift)) | (v << shift);
  }
  return word;
}

}  // namespace wgrt
