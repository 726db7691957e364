// wgrt_device.cuh -- device helpers shared by the strict walk, the fast walk and the unit hooks.
//
// "literal" functions evaluate exactly the expressions of the reference device functions
// (GPU_ray_tracing_functions.py, GRTF below) in the same order; translation units that must
// stay comparable with the CPU oracle are compiled with -fmad=false.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/wgrt.h"

// Checked build (libwgrt_checked.so, -DWGRT_CHECKED): every index the production walk forms into shared
// memory, the Jones scratch, the atlas, the region grids and the ray / bin arrays is asserted to be in
// range; violations are counted (wgrt_debug_check_failures) instead of corrupting memory.  This stands in
// for compute-sanitizer's memcheck, which the GPU pool does not allow.  In the normal build the macro
// vanishes.
#if defined(WGRT_CHECKED) && defined(WGRT_CHECK_TU)
static __device__ unsigned long long wgrt_check_fail_count;
#define WGRT_CHECK(cond)                                              \
  do {                                                                \
    if (!(cond)) atomicAdd(&wgrt_check_fail_count, 1ull);             \
  } while (0)
#else
#define WGRT_CHECK(cond) ((void)0)
#endif

namespace wgrt {

constexpr unsigned FULL_MASK = 0xffffffffu;

struct cplx {
  double re, im;
};

__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }

// GRTF:25-34 with the state kept in a register.
__device__ __forceinline__ double xorshift_draw(uint32_t& s, int64_t index) {
  if (s == 0u) s = 0x6D2B79F5u ^ static_cast<uint32_t>(index + 1);
  s ^= s << 13;
  s ^= s >> 17;
  s ^= s << 5;
  return static_cast<double>(s) * (1.0 / 4294967296.0);
}

// Per-thread event counts, flushed once per thread with warp-aggregated atomics.
struct Counts {
  unsigned long long c[WGRT_NUM_COUNTERS];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < WGRT_NUM_COUNTERS; ++k) c[k] = 0ull;
  }
  __device__ __forceinline__ void flush(unsigned long long* global) {
#pragma unroll
    for (int k = 0; k < WGRT_NUM_COUNTERS; ++k) {
      unsigned long long v = c[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(global + k, v);
    }
  }
};

// Rays the fast walk refuses to decide (a draw within TIE_TOL of a threshold, wgrt_walk.cu): their
// launch-relative indices, re-walked literally by walk_redo_kernel (wgrt_strict.cu) after the walk.
constexpr unsigned REDO_CAP = 1u << 16;
struct RedoList {
  unsigned count;        // pushes attempted in this launch (may exceed REDO_CAP)
  unsigned pad_[3];
  long long idx[REDO_CAP];
};
// false when the list is full: the caller then keeps its own decision (parity falls back to
// "a few ulp from the threshold" for that ray; needs > 65536 near ties in ONE launch)
__device__ __forceinline__ bool redo_push(RedoList* r, long long ray) {
  const unsigned slot = atomicAdd(&r->count, 1u);
  if (slot >= REDO_CAP) return false;
  r->idx[slot] = ray;
  return true;
}

// GRTF:52-61
template <bool COUNT>
__device__ __forceinline__ bool on_segment_literal(double px, double py, double x1, double y1, double x2,
                                                   double y2, double tol, Counts* cn) {
  if ((px < fmin(x1, x2) - tol) || (px > fmax(x1, x2) + tol) || (py < fmin(y1, y2) - tol) ||
      (py > fmax(y1, y2) + tol))
    return false;
  if (COUNT) cn->c[WGRT_CNT_CROSS]++;
  return fabs((x2 - x1) * (py - y1) - (y2 - y1) * (px - x1)) <= tol;
}

// GRTF:63-71 then GRTF:36-50 over ring [start, end) of a [V,2] vertex array.
template <bool COUNT>
__device__ __noinline__ bool inside_or_on_edge_literal(double px, double py, const double* __restrict__ poly,
                                                       int64_t start, int64_t end, Counts* cn) {
  const int64_t nv = end - start;
  if (COUNT) cn->c[WGRT_CNT_POLY_TESTS]++;
  int64_t j = nv - 1;
  for (int64_t i = 0; i < nv; ++i) {
    const double2 a = *reinterpret_cast<const double2*>(poly + 2 * (start + j));
    const double2 b = *reinterpret_cast<const double2*>(poly + 2 * (start + i));
    if (COUNT) cn->c[WGRT_CNT_EDGE_VISITS]++;
    if (on_segment_literal<COUNT>(px, py, a.x, a.y, b.x, b.y, 1e-12, cn)) return true;
    j = i;
  }
  bool inside = false;
  j = nv - 1;
  for (int64_t i = 0; i < nv; ++i) {
    const double2 vi = *reinterpret_cast<const double2*>(poly + 2 * (start + i));
    const double2 vj = *reinterpret_cast<const double2*>(poly + 2 * (start + j));
    if (COUNT) cn->c[WGRT_CNT_EDGE_VISITS]++;
    if ((vi.y > py) != (vj.y > py)) {
      if (COUNT) cn->c[WGRT_CNT_STRADDLE]++;
      if (px < (vj.x - vi.x) * (py - vi.y) / (vj.y - vi.y + 1e-20) + vi.x) inside = !inside;
    }
    j = i;
  }
  return inside;
}

template <bool COUNT>
__device__ __forceinline__ int first_hit_literal(double x, double y, const double* __restrict__ poly,
                                                 const int64_t* __restrict__ off, int64_t npoly, Counts* cn) {
  for (int64_t i = 0; i < npoly; ++i)
    if (inside_or_on_edge_literal<COUNT>(x, y, poly, off[i], off[i + 1], cn)) return static_cast<int>(i);
  return -1;
}

// GRTF:124-130
__device__ __forceinline__ double wrap_pi_literal(double x) {
  const double pi = 3.141592653589793;
  const double two_pi = 2.0 * pi;
  x = x + pi;
  x = x - two_pi * floor(x / two_pi);
  x = x - pi;
  return x;
}

// GRTF:132-152; q in CALL order (E_te_te, E_te_tm, E_tm_te, E_tm_tm)
__device__ __forceinline__ void efield_literal(double Ete_abs, double Etm_abs, double delta, const cplx q[4],
                                               double& o_te, double& o_tm, double& o_delta) {
  double sn, cs;
  sincos(delta, &sn, &cs);
  const cplx phase{cs, sn};
  const cplx te_in{Ete_abs, 0.0};
  const cplx tm_in = cmul(phase, cplx{Etm_abs, 0.0});
  const cplx a = q[0], b = q[2], c = q[1], d = q[3];
  const cplx Ete_out = cadd(cmul(a, te_in), cmul(b, tm_in));
  const cplx Etm_out = cadd(cmul(c, te_in), cmul(d, tm_in));
  o_te = hypot(Ete_out.re, Ete_out.im);
  o_tm = hypot(Etm_out.re, Etm_out.im);
  const double eps = 1e-20;
  const double phi_te = o_te >= eps ? atan2(Ete_out.im, Ete_out.re) : 0.0;
  const double phi_tm = o_tm >= eps ? atan2(Etm_out.im, Etm_out.re) : 0.0;
  o_delta = wrap_pi_literal(phi_tm - phi_te);
}

// GRTF:154-165 with the wavelength slice of GRTF:1168.  Deposits are all 1.0f, so float32
// accumulation is exact and order independent (counts stay far below 2^24).
__device__ __forceinline__ void deposit_bin(const wgrt_problem_t& p, int64_t lm, int64_t m, int64_t n, double x,
                                            double y, double xmin, double xmax, double ymin, double ymax) {
  const double dx = (xmax - xmin) / static_cast<double>(p.EBx);
  const double dy = (ymax - ymin) / static_cast<double>(p.EBy);
  const int64_t ix = static_cast<int64_t>(floor((x - xmin) / dx));
  const int64_t iy = static_cast<int64_t>(floor((y - ymin) / dy));
  const int64_t flat = (((lm * p.Y + n) * p.X + m) * p.EBy + iy) * p.EBx + ix;
  const int64_t total = p.L * p.Y * p.X * p.EBy * p.EBx;
  WGRT_CHECK(ix >= 0 && ix <= p.EBx && iy >= 0 && iy <= p.EBy);   // == only for a point exactly on the max edge
  if (flat >= 0 && flat < total) atomicAdd(p.matrix_EB + flat, 1.0f);
}

// ---- launch plumbing shared by the translation units (defined in wgrt_api.cu / per TU) -------
struct RegionStatic {          // host-known part of one region set
  const double* verts;         // [V,2]
  const int64_t* offsets;      // [npoly+1]; nullptr = one ring [0, nverts)
  uint8_t* coarse;             // [nc*nc] coarse cell codes (small: stays in L1)
  uint8_t* cells;              // [n*n] fine cell codes; only cells under MIXED coarse cells are valid
  uint32_t* detail;            // [n*n] ring range to test exactly in ambiguous fine cells
  uint32_t* rowmask;           // [n][words] edges relevant to each fine cell row
  uint32_t* rowmask_coarse;    // [nc][words] same per coarse row (used while building)
  int nverts, npoly, n, nc, shift, words;  // n = nc << shift fine cells per axis
};
struct RegionDyn {             // computed on the device from the vertex data
  double x0, y0, inv_dx, inv_dy, cell_dx, cell_dy;
};
enum { REG_IC = 0, REG_R1, REG_R2, REG_FC, REG_OC, NUM_REGIONS };
constexpr uint8_t CELL_NONE = 255, CELL_AMBIG = 254;

// The atlas: ONE coarse grid over the union of all region sets' bounding boxes whose 32-bit words
// answer all five region queries for a point at once (see wgrt_region.cuh).
struct AtlasDyn {              // computed on the device
  double x0, y0, inv_dx, inv_dy;
};
constexpr int ATLAS_N = 64;    // level 1: ATLAS_N x ATLAS_N words = 16 KB, stays in L1
constexpr int ATLAS_SUB_SHIFT = 6;                      // level 2: every level-1 cell split 64 x 64 (measured on C2:
                                                        // 32 x 32 9.2 ms, 64 x 64 8.8 ms, 128 x 128 8.7 ms)
constexpr int ATLAS_N2 = ATLAS_N << ATLAS_SUB_SHIFT;    // 4096 x 4096 words = 64 MB; only
                                                        // cells under MIXED level-1 cells are populated

// The ZONE form of the atlas (what the production walk reads).  Every distinct atlas word -- a combination
// of (in-coupler, effective region 1 / 2, fold slice, out-coupler slice) answers, certain or MIXED -- is a
// "zone" with a 16-bit id; the two grid levels store zone ids instead of words (half the bytes, and twice the
// first-level resolution in the same L1 footprint), and a per-geometry TRANSITION TABLE trans[state][zone]
// holds what the walk's loop head does with a ray of that region state in that zone: next state, lost,
// event row or free-bounce vector / phase.  The table is produced by running the walk's own decode function
// on every (state, zone word) pair, so the two cannot disagree; entries whose state needs a MIXED field are
// flagged and take the word path (zone word -> atlas_resolve -> decode).
#ifndef WGRT_ZONE_REFINE
#define WGRT_ZONE_REFINE 1                         // level-1 zone cells per atlas level-1 cell and axis: 2^refine
#endif
constexpr int ZONE_REFINE = WGRT_ZONE_REFINE;
constexpr int ZONE_N1 = ATLAS_N << ZONE_REFINE;    // level 1: 128 x 128 uint16 = 32 KB (refine 1), 64 x 64 = 8 KB (refine 0)
constexpr int ZONE_SUB_SHIFT = ATLAS_SUB_SHIFT - ZONE_REFINE;   // level 2 = the 4096 x 4096 level-2 grid of the atlas
constexpr int ZONE_CAP = 4096;                     // distinct zones supported (hash slots 2 x that)
constexpr uint16_t ZONE_MIXED = 0xFFFFu;           // level-1 marker: ask level 2
constexpr int ZONE_STATES = 8;                     // region states 0..5 and the two pending states
struct ZoneDyn {                                   // computed on the device
  double x0, y0, inv_dx, inv_dy;                   // level-1 cell coordinates (same box as the atlas)
  int num_zones;                                   // <= ZONE_CAP, else the zone form is invalid (word path everywhere)
  int outside_zone;                                // zone of points outside the atlas box
  int valid, pad_;
};
struct ZoneSet {
  uint16_t* level1;                 // [ZONE_N1 * ZONE_N1]
  uint16_t* level2;                 // [ATLAS_N2 * ATLAS_N2], populated under MIXED level-1 cells
  uint32_t* words;                  // [ZONE_CAP] zone -> atlas word
  uint32_t* trans;                  // [ZONE_STATES * ZONE_CAP] transition table (wgrt_walk.cu: decode_transition)
  uint32_t* hash_keys;              // [2 * ZONE_CAP] open addressing, word -> slot
  uint16_t* hash_zone;              // [2 * ZONE_CAP] slot -> zone id
  uint32_t* words1;                 // [ZONE_N1 * ZONE_N1] scratch: level-1 words at the zone resolution
  ZoneDyn* dyn;
};

struct RegionSet {
  RegionStatic st[NUM_REGIONS];
  ZoneSet zones;
  uint32_t* atlas;                  // device [ATLAS_N * ATLAS_N] then [ATLAS_N2 * ATLAS_N2]
  AtlasDyn* atlas_dyn;              // device
  void* regions;                    // device Region[NUM_REGIONS] (wgrt_region.cuh), refreshed at every index build
  RegionDyn* dyn;                   // device array [NUM_REGIONS]
  unsigned long long* hash_state;   // device: {hash of the index now built, dirty flag of this launch}
  const unsigned long long* dirty;  // = hash_state + 1
};

cudaError_t launch_walk_strict(const wgrt_problem_t& p, unsigned long long* counters, cudaStream_t s);
cudaError_t launch_region_build(const RegionSet& rs, bool force, cudaStream_t s);
// transition-table entry for (state, zone word): defined next to the walk that applies it (wgrt_walk.cu)
cudaError_t launch_zone_transitions(const RegionSet& rs, int n_FC, int n_OC, cudaStream_t s);
void set_tie_tolerance(double tol);
cudaError_t walk_check_failures(unsigned long long* out, bool reset);   // cudaErrorNotSupported unless WGRT_CHECKED
cudaError_t launch_debug_deposit_inside(const double* rect, const double* px, const double* py, int64_t n, int32_t* out,
                                        int literal, cudaStream_t s);
size_t walk_warp_scratch_bytes(const wgrt_problem_t& p, int num_sms);
cudaError_t launch_walk_warp(const wgrt_problem_t& p, const RegionSet& rs, int* work_counter,
                             unsigned long long* counters, int num_sms, double* jones_scratch, RedoList* redo,
                             cudaStream_t s);
cudaError_t launch_walk_redo(const wgrt_problem_t& p, const RedoList* redo, unsigned long long* counters, cudaStream_t s);
cudaError_t launch_debug_locate_literal(const double* verts, const int64_t* off, int64_t npoly, const double* px,
                                        const double* py, int64_t n, int32_t* out, cudaStream_t s);
cudaError_t launch_debug_locate_grid(const RegionSet& rs, int region, const double* px, const double* py, int64_t n,
                                     int32_t* out, unsigned long long* counters, int via_atlas, cudaStream_t s);
cudaError_t launch_debug_efield(const double* ete, const double* etm, const double* delta, const double* jones,
                                int64_t n, double* out, cudaStream_t s);
cudaError_t launch_debug_xorshift(uint32_t* states, int64_t n, int draws, double* out_last, cudaStream_t s);
cudaError_t launch_seed_rng(uint32_t* states, int64_t n, int64_t first_index, cudaStream_t s);
cudaError_t launch_fma_peak(int num_sms, double* fp64_tflops, double* fp32_tflops);
cudaError_t launch_bins_pack_u8(const float* bins, int64_t n, uint8_t* out, unsigned* stats, float limit, int num_sms,
                                cudaStream_t s);
cudaError_t launch_bins_unpack_u8(const uint8_t* in, int64_t n, float* bins, int num_sms, cudaStream_t s);
cudaError_t launch_eval_metrics(const float* perceive, int64_t Yf, int64_t Xf, int n_epy, int n_epx,
                                const wgrt_eval_params_t& prm, double* metrics, float* image, cudaStream_t s);
cudaError_t launch_legacy_step(const wgrt_legacy_problem_t& p, const RegionSet& rs, bool soa, unsigned long long* dropped,
                               cudaStream_t s);
cudaError_t launch_legacy_pack(const double* src, double* dst, int64_t src_len, int64_t src_cap, int64_t dst_cap, bool soa,
                               int32_t* out_count, cudaStream_t s);
cudaError_t launch_legacy_transpose(const double* src, double* dst, int64_t n, int64_t cap, bool to_soa, cudaStream_t s);
cudaError_t launch_pupil_sums(const float* EB, int64_t L, int64_t Yf, int64_t Xf, int64_t EBy, int64_t EBx,
                              int mask, int step_y, int step_x, float* out, float* cell_sums, cudaStream_t s);

}  // namespace wgrt
