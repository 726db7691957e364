// wgrt_walk.cu -- the production ray walk for sm_100a: one persistent CTA per SM, 24 independent warps in it, one
// warp = one cell at a time, every per-step access in shared memory.
//
// Same Monte-Carlo walk as process_rays_kernel_pro_fullColor (GRTF:833-1246), organised so that a
// warp instruction almost always does useful work in almost every lane:
//
//  * A warp claims tiles of consecutive rays (whole FoV-wavelength cells when the rays arrive cell by cell,
//    gpu_ray_tracing_pro_fullColor.py:82-115) and walks them alone.  The warps of a CTA meet once, at the barrier
//    behind the cooperative load of the geometry's zone tables; after it there are no block barriers and no
//    cross-warp queues.  The end-of-cell drain (lanes idling while the longest paths finish) is paid once per
//    ~5000 rays per warp; the last half tile per warp is handed out in quarters so that the SMs do not run half
//    empty while the last whole tiles finish.
//  * Shared memory of the CTA (227 KB): the zone tables of the geometry once (level-1 grid as 8-bit ids 16 KB,
//    transition table 4 KB) and, per warp, the cell's event table (48-byte rows: quadratic form, 1 / cos, meta),
//    the Jones matrices of its orders (64-byte rows, XOR-swizzled), the per-cell constants and the survivor stack:
//    8.6 KB per warp for BASELINE's designs (70 event rows), i.e. 24 warps.  The descriptors of the region index
//    (grid origins / pitches, zone count) sit in constant memory.  What is left in global memory per step: the
//    level-2 zone cell under MIXED level-1 cells (14 % of the lookups), the ray streams of the in-coupling batches,
//    RNG states and bins.  Designs whose tables do not fit keep the Jones rows in a per-warp global scratch (JSM
//    off) and / or read zone ids and transitions from global memory; same code, template / uniform branches.
//  * In-coupling decisions are taken 32 rays at a time with every lane busy (coalesced loads issued together, the
//    next batch prefetched into L1, cut where the cell key changes: run detection costs nothing extra); the ~72 %
//    of rays absorbed there never reach the walk.  Survivors wait on a per-warp shared-memory stack (raw ray, RNG
//    state, chosen order) until a lane is free.
//  * A step = refill, phase B, phase A.  Phase B ("diffract"): lanes that stand on a grating draw,
//    evaluate the efficiencies of ALL orders at once from per-cell quadratic forms (4 FMAs per order
//    instead of a Jones application per order in a divergent if-chain) and pick the order exactly as
//    the reference's if/elif chain does; after a __syncwarp() ONE copy of the order application runs
//    for all of them and for the lanes that just popped a survivor -- one Jones matrix, the chosen
//    one.  Phase A ("go to the next grating"): every lane whose ray moved looks up its ZONE (wgrt_device.cuh:
//    one id per cell names the combination of in-coupler / effective region 1 / 2 / fold slice / out-coupler
//    slice answers there) and the TRANSITION-TABLE entry of (region state, zone): next state, lost, the event's
//    first table row, or the free bounce (GRTF:1049-1052, 1102-1108, 1175-1178) after which it asks again in the
//    next step.
//  * The polarisation state is the un-normalised complex Jones vector (te, tm) plus s = 1/|v|^2.
//    E_field_cal's cos / sin / hypot / atan2 / wrap (GRTF:136-150) and the per-event normalisation
//    (GRTF:876-877 ff.) disappear: efficiencies are v^H M v * s with M = J^H J precomputed per cell
//    and order, TIR phases are complex multiplies by per-cell phasors.  Algebraically the same map;
//    numerically a few ulp apart, i.e. a decision `u <= efficiency` could flip only when u lands within
//    ~1e-15 of a threshold.  Parity is nevertheless BY CONSTRUCTION: a ray whose draw lands within
//    TIE_TOL = 1e-10 of any threshold it is compared with (five orders of magnitude more than the
//    two evaluations can differ) is not decided here at all -- the lane drops it untouched (RNG state
//    not written, nothing deposited: a deposit ends a ray) onto the launch's redo list, and
//    walk_redo_kernel (wgrt_strict.cu) walks it from its start with the reference's literal
//    expressions right after this kernel.  Expected: ~0.3 such rays per 112.5 M-ray launch; their
//    number is reported (WGRT_CNT_NEAR_TIE).
//  * What bounds it (ncu, profiles/r2_walk_warp_ncu_summary.txt): issue slots 67 % busy at 6 warps per scheduler,
//    each warp issuing every ~9 cycles (fixed-latency dependencies 2.6, shared-memory latency 1.0, branches 1.0,
//    the remaining global loads 1.4); 19 of 32 lanes active per instruction (lanes between gratings skip phase B,
//    single-lane literal polygon tests near edges take 15 % of the instructions).
#include <climits>
#include <cstdio>
#include <cstring>

#define WGRT_CHECK_TU 1   // this translation unit carries the bounds assertions of the checked build
#include "wgrt_region.cuh"

namespace wgrt {

namespace {

// Per-cell event table, one row per (event, order), in the warp's slice of shared memory.  A row is what an event
// evaluation reads, laid out for 128-bit shared loads: [Q0 Q1 | Q2 Q3 | 1/cos, meta] -- the quadratic form of the
// order's efficiency, 1 / cos of the new direction and the row's meta word (in the low half of the sixth double).
// The row stride of 12 words puts 8 consecutive rows on disjoint 4-bank groups.  The order's Jones matrix (8 doubles,
// read for the chosen order only; deposit orders have none) lives behind the table in the JSM layout -- 64-byte rows
// whose four 16-byte chunks are XOR-swizzled by the row pair, so that 8 consecutive rows are conflict free as well --
// and in a per-warp global scratch (L1 / L2 resident) otherwise.
constexpr int ROW = 6;
constexpr int R_Q = 0;             // [0..3] quadratic form of the order's efficiency (cos factor folded in)
constexpr int R_INVCOS = 4;        // 1 / cos(theta_new)
constexpr int R_META = 5;          // meta word (integer in the low 32 bits)
constexpr int JROW = 8;            // doubles per Jones row
constexpr int ST_DEAD = -1, ST_PEND_FWD = 6, ST_PEND_BACK = 7;
// |u - cumulative efficiency| below this: the ray is re-walked literally (see the file header)
constexpr double TIE_TOL_DEFAULT = 1e-10;
double g_tie_tol = TIE_TOL_DEFAULT;   // wgrt_debug_set_tie_tolerance (tests widen it to exercise the redo path)
enum { POST_NONE = 0, POST_IC_FWD = 1, POST_IC_BACK = 2, POST_DEPOSIT = 3 };
enum { EV_INIT = 0, EV_S0, EV_S1, EV_S2, EV_S3, EV_S4, EV_S5, NUM_EV };
enum { DIR_IC1 = 0, DIR_IC2, DIR_IC3, DIR_FC1, DIR_FC2, DIR_OC1, DIR_OC2 };

struct OrderSpec {
  int8_t ch[4];   // LUT channels in E_field_cal CALL order (E_te_te, E_te_tm, E_tm_te, E_tm_tm)
  int8_t dir;     // whose channel 0 gives the outgoing polar angle (numerator cosine)
  int8_t tir;     // lut_TIR index added to the phase
  int8_t gap;     // lut_gap pair index of the new bounce vector
  int8_t nstate;  // region state after the order is taken
  int8_t post;    // what happens after the move
  int8_t fmode;   // 0: cos ratio, 1: * n_g (air -> glass), 2: / n_g (glass -> air)
};

// Transcribed from the kernel's call sites: INIT GRTF:860-904, state 0 GRTF:908-953, state 1
// GRTF:954-999 (note the swapped te_tm/tm_te channels at GRTF:957-958), state 2 GRTF:1000-1052,
// state 3 GRTF:1053-1108, state 4 GRTF:1110-1178, state 5 GRTF:1179-1246.
__constant__ OrderSpec kOrders[NUM_EV][3] = {
    /* INIT */ {{{13, 18, 33, 38}, DIR_IC2, 0, 0, 0, POST_IC_FWD, 1},
                {{15, 20, 35, 40}, DIR_IC3, 2, 2, 1, POST_IC_BACK, 1},
                {{0, 0, 0, 0}, 0, 0, 0, 0, 0, 0}},
    /* S0   */ {{{4, 9, 24, 29}, DIR_IC2, 0, 0, 0, POST_IC_FWD, 0},
                {{6, 11, 26, 31}, DIR_IC3, 2, 2, 1, POST_IC_BACK, 0},
                {{0, 0, 0, 0}, 0, 0, 0, 0, 0, 0}},
    /* S1   */ {{{2, 22, 7, 27}, DIR_IC2, 0, 0, 0, POST_IC_FWD, 0},
                {{4, 9, 24, 29}, DIR_IC3, 2, 2, 1, POST_IC_BACK, 0},
                {{0, 0, 0, 0}, 0, 0, 0, 0, 0, 0}},
    /* S2   */ {{{3, 6, 15, 18}, DIR_FC1, 0, 0, 2, POST_NONE, 0},
                {{2, 5, 14, 17}, DIR_FC2, 1, 1, 3, POST_NONE, 0},
                {{0, 0, 0, 0}, 0, 0, 0, 0, 0, 0}},
    /* S3   */ {{{4, 7, 16, 19}, DIR_FC1, 0, 0, 2, POST_NONE, 0},
                {{3, 6, 15, 18}, DIR_FC2, 1, 1, 3, POST_NONE, 0},
                {{0, 0, 0, 0}, 0, 0, 0, 0, 0, 0}},
    /* S4   */ {{{4, 9, 24, 29}, DIR_OC1, 1, 1, 4, POST_NONE, 0},
                {{2, 7, 22, 27}, DIR_OC2, 3, 3, 5, POST_NONE, 0},
                {{13, 18, 33, 38}, DIR_IC1, 0, 0, 0, POST_DEPOSIT, 2}},
    /* S5   */ {{{6, 11, 26, 31}, DIR_OC1, 1, 1, 4, POST_NONE, 0},
                {{4, 9, 24, 29}, DIR_OC2, 3, 3, 5, POST_NONE, 0},
                {{15, 20, 35, 40}, DIR_IC1, 0, 0, 0, POST_DEPOSIT, 2}},
};

// meta word of a row: bits 0-1 TIR index, 2-3 gap pair, 4-6 next state, 7-8 post action; on the
// FIRST row of an event additionally: bit 9 the event has three orders, bit 10 its branches carry
// `and ener_k > threshold` (GRTF:1020 ff.; absent at GRTF:871-999)
constexpr int META_THREE = 1 << 9, META_GATED = 1 << 10;
constexpr int META_JROW_SHIFT = 12;   // bits 12-27: the row's Jones row in the JSM layout

// sinfo[state]: which coupler decides the event, first row, rows per slice, what a miss means,
// which doubled TIR phase / bounce vector a free bounce uses
enum : int {
  SI_REGION_MASK = 7, SI_REGION_NONE = 7,          // bits 0-2: REG_FC / REG_OC / none (in-coupler states)
  SI_ROWBASE_SHIFT = 3, SI_ROWBASE_MASK = 0xfff,   // bits 3-14
  SI_STRIDE_SHIFT = 15,                            // bits 15-16
  SI_MISS_SHIFT = 17,                              // bits 17-18: 0 bounce on, 1 test eff_reg2 first, 2 lost
  SI_PHASE_SHIFT = 19,                             // bit 19
  SI_GAP_SHIFT = 20                                // bits 20-21
};
__host__ __device__ constexpr int make_sinfo(int region, int rowbase, int stride, int miss, int phase, int gap) {
  return region | (rowbase << SI_ROWBASE_SHIFT) | (stride << SI_STRIDE_SHIFT) | (miss << SI_MISS_SHIFT) |
         (phase << SI_PHASE_SHIFT) | (gap << SI_GAP_SHIFT);
}

// region sets whose atlas field a ray in a given state can need (bit r = set r), 5 bits per state
// packed into one word (a per-lane index into __constant__ memory would serialise the warp):
// states 0..5, 6 = pending after a +1 in-coupler order, 7 = pending after a -1 order
__host__ __device__ constexpr unsigned long long need_bits(int state, unsigned sets) {
  return static_cast<unsigned long long>(sets) << (5 * state);
}
constexpr unsigned long long kNeedPacked =
    need_bits(0, 1u << REG_R1) | need_bits(1, 1u << REG_R1) | need_bits(2, (1u << REG_R1) | (1u << REG_FC)) |
    need_bits(3, (1u << REG_R1) | (1u << REG_FC) | (1u << REG_R2) | (1u << REG_OC)) |   // may fall through to state 4
    need_bits(4, (1u << REG_R1) | (1u << REG_OC)) | need_bits(5, (1u << REG_R1) | (1u << REG_OC)) |
    need_bits(6, (1u << REG_IC) | (1u << REG_R1) | (1u << REG_FC)) |                    // becomes 0 or 2
    need_bits(7, (1u << REG_IC) | (1u << REG_R1));                                      // becomes 1 or ends

// sinfo[state] for a design with nFC fold-coupler and nOC out-coupler slices: states 0 and 2 travel along the +1
// in-coupled direction (gap pair 0), state 1 along the -1 direction (2), states 3 and 4 along the folded direction
// (1), state 5 along the conjugate out-coupler direction (3)
__host__ __device__ constexpr int sinfo_of(int state, int nFC, int nOC) {
  return state == 0 ? make_sinfo(SI_REGION_NONE, 2, 0, 0, 0, 0)
       : state == 1 ? make_sinfo(SI_REGION_NONE, 4, 0, 0, 0, 2)
       : state == 2 ? make_sinfo(REG_FC, 6, 2, 0, 0, 0)                          // miss: bounce, 2 T[0]
       : state == 3 ? make_sinfo(REG_FC, 6 + 2 * nFC, 2, 1, 1, 1)                // miss: eff_reg2 test, 2 T[1]
       : state == 4 ? make_sinfo(REG_OC, 6 + 4 * nFC, 3, 0, 1, 1)                // miss: bounce, 2 T[1]
       : state == 5 ? make_sinfo(REG_OC, 6 + 4 * nFC + 3 * nOC, 3, 2, 0, 3)      // miss: lost
       : 0;
}

// What the loop head (GRTF:905-907 and the region tests of the state that follows) does with a ray of region
// state `state` standing in a zone whose atlas word is `word` -- the walk's "go to the next grating" phase as a
// pure function, tabulated per geometry as trans[state][zone] (wgrt_device.cuh: ZoneSet):
//   bits 0-2  region state afterwards        bit 3   the ray is lost (left the effective region / the in-coupler on
//   bit 4     it stands on a grating: bits 12-27 = first event-table row      the -1 order / state-5 miss)
//   bits 5-6  else: free bounce along gap pair, bit 7 its doubled TIR phase
//   bits 8-9  loop iterations consumed (2 when state 3 falls through to state 4, GRTF:1102-1104)
//   bit 31    a field this state needs is MIXED in the word: resolve it first (word path)
constexpr uint32_t ACT_LOST = 1u << 3, ACT_EVENT = 1u << 4, ACT_RESOLVE = 1u << 31;
__host__ __device__ inline uint32_t mixed_fields(uint32_t word) {
  return (((word >> ATLAS_SHIFT_IC) & 3u) == 2u ? 1u << REG_IC : 0u) | (((word >> ATLAS_SHIFT_R1) & 3u) == 2u ? 1u << REG_R1 : 0u) |
         (((word >> ATLAS_SHIFT_R2) & 3u) == 2u ? 1u << REG_R2 : 0u) |
         (((word >> ATLAS_SHIFT_FC) & 0xffu) == CELL_AMBIG ? 1u << REG_FC : 0u) |
         (((word >> ATLAS_SHIFT_OC) & 0xffu) == CELL_AMBIG ? 1u << REG_OC : 0u);
}
__host__ __device__ inline uint32_t decode_transition(uint32_t word, int state, int nFC, int nOC) {
  if (static_cast<uint32_t>(kNeedPacked >> (5 * state)) & 31u & mixed_fields(word)) return ACT_RESOLVE;
  const bool in_ic = ((word >> ATLAS_SHIFT_IC) & 3u) == 1u;
  const bool in_r1 = ((word >> ATLAS_SHIFT_R1) & 3u) == 1u;
  const bool in_r2 = ((word >> ATLAS_SHIFT_R2) & 3u) == 1u;
  const int fc = (word >> ATLAS_SHIFT_FC) & 0xff, oc = (word >> ATLAS_SHIFT_OC) & 0xff;
  int st = state, inc = 0;
  bool lost = false;
  if (st == ST_PEND_FWD) st = in_ic ? 0 : 2;                  // GRTF:883-886
  else if (st == ST_PEND_BACK) { st = 1; lost = !in_ic; }     // GRTF:899-902
  if (!lost) { inc = 1; lost = !in_r1; }                      // GRTF:905-907
  int sinfo = sinfo_of(st, nFC, nOC);
  const int region = sinfo & SI_REGION_MASK;
  int code = region == REG_FC ? fc : region == REG_OC ? oc : 0;
  if (code == CELL_NONE && ((sinfo >> SI_MISS_SHIFT) & 3) == 1 && !in_r2 && !lost) {   // GRTF:1102-1104
    st = 4; inc = 2;
    sinfo = sinfo_of(4, nFC, nOC);
    code = oc;
  }
  uint32_t act = static_cast<uint32_t>(st) | (static_cast<uint32_t>(inc) << 8);
  if (lost) return act | ACT_LOST;
  if (code != CELL_NONE)
    return act | ACT_EVENT | (static_cast<uint32_t>(((sinfo >> SI_ROWBASE_SHIFT) & SI_ROWBASE_MASK) + ((sinfo >> SI_STRIDE_SHIFT) & 3) * code) << 12);
  if (((sinfo >> SI_MISS_SHIFT) & 3) == 2) return act | ACT_LOST;                      // GRTF:1244-1246
  return act | (static_cast<uint32_t>((sinfo >> SI_GAP_SHIFT) & 3) << 5) | (static_cast<uint32_t>((sinfo >> SI_PHASE_SHIFT) & 1) << 7);
}

// Per-launch descriptors of the region index, copied device-to-device from where the index builders left them
// (RegionSet::atlas_dyn, ZoneSet::dyn) into constant memory right before the launch: every lane reads them at
// every step, and constant-bank reads cost neither shared-memory wavefronts nor registers.
struct WalkConst {
  AtlasDyn atlas;
  ZoneDyn zone;
};
__constant__ WalkConst c_walk;

// Zone tables of the geometry in the CTA's shared memory (all warps of the SM read them at every step): the level-1
// grid as 8-bit zone ids (designs with at most 254 zones; BASELINE's have ~40) and the transition table (at most
// TRANS_SM zones).  Larger designs read the 16-bit grid / the table from global memory.
constexpr int TRANS_SM = 128;
constexpr int ZONE_SM_MAX = 254;
constexpr uint8_t ZONE_MIXED8 = 0xFFu;
struct alignas(16) CtaShared {
  uint8_t level1[ZONE_N1 * ZONE_N1];
  uint32_t trans[ZONE_STATES * TRANS_SM];
};

// the word path of the loop head: used where the table says "resolve first" and when there is no table
template <bool COUNT>
__device__ __noinline__ uint32_t resolve_transition(const RegionSet& rs, int z, int state, double x, double y, int nFC,
                                                    int nOC, Counts* cn) {
  uint32_t word;
  if (z >= 0) {
    word = __ldg(rs.zones.words + z);
  } else {
    Atlas atlas;
    atlas.x0 = c_walk.atlas.x0; atlas.y0 = c_walk.atlas.y0; atlas.inv_dx = c_walk.atlas.inv_dx; atlas.inv_dy = c_walk.atlas.inv_dy;
    atlas.words = rs.atlas;
    atlas.words2 = rs.atlas + ATLAS_N * ATLAS_N;
    word = atlas_lookup(atlas, x, y);
  }
  if (word & ATLAS_ANY_MIXED)
    word = atlas_resolve<COUNT>(word, static_cast<uint32_t>(kNeedPacked >> (5 * state)) & 31u,
                                static_cast<const Region*>(rs.regions), x, y, cn);
  return decode_transition(word, state, nFC, nOC);
}

// zone of a point: level 1 from shared memory (SM) or global memory, level 2 (under MIXED level-1 cells) from
// global memory
template <bool SM>
__device__ __forceinline__ int zone_lookup_walk(const uint8_t* __restrict__ level1_sm, const uint16_t* __restrict__ level1,
                                                const uint16_t* __restrict__ level2, double x, double y) {
  const double fx = (x - c_walk.zone.x0) * c_walk.zone.inv_dx, fy = (y - c_walk.zone.y0) * c_walk.zone.inv_dy;
  const double lim = static_cast<double>(ZONE_N1);
  if (!(fx >= 0.0 && fy >= 0.0 && fx < lim && fy < lim)) return c_walk.zone.outside_zone;   // also NaN
  const int cell = static_cast<int>(fy) * ZONE_N1 + static_cast<int>(fx);
  int z;
  bool mixed;
  if (SM) {
    z = level1_sm[cell];
    mixed = z == ZONE_MIXED8;
  } else {
    z = __ldg(level1 + cell);
    mixed = z == ZONE_MIXED;
  }
  if (mixed) {
    const double sub = static_cast<double>(1 << ZONE_SUB_SHIFT);
    const int ix = min(static_cast<int>(fx * sub), ATLAS_N2 - 1), iy = min(static_cast<int>(fy * sub), ATLAS_N2 - 1);
    z = __ldg(level2 + static_cast<size_t>(iy) * ATLAS_N2 + ix);
  }
  return z;
}

struct alignas(16) CellConst {
  cplx ph1[4];      // e^{i T[k]}
  cplx ph2[4];      // e^{i 2 T[k]}
  double gap[8];    // lut_gap[lm, m, n, :]
  double rect[8];   // eff_reg_FOV[m, n, :, :]
  double range[4];  // eff_reg_FOV_range[m, n, :]
  double box[4];    // xmin, xmax, ymin, ymax of the eyebox rectangle when it is axis aligned
  double inv_cos_in;
  double bin_dx, bin_dy;   // (xmax - xmin) / EBx, (ymax - ymin) / EBy of the cell's eyebox range (GRTF:154-157)
  long long bin_base;      // flat index of bin (0, 0) of cell (lm, n, m) in matrix_EB
  int box_ok;       // rect is exactly the axis-aligned rectangle `box` in the runner's vertex order
  int pad_;
};

// deposit_bin (wgrt_device.cuh, GRTF:154-165) with the per-cell quotients and the cell's base index taken from the
// table: the same operations on the same operands, evaluated once per cell instead of once per deposit.
__device__ __forceinline__ void deposit_bin_cell(const wgrt_problem_t& p, const CellConst& cc, double x, double y) {
  const int64_t ix = static_cast<int64_t>(floor((x - cc.range[0]) / cc.bin_dx));
  const int64_t iy = static_cast<int64_t>(floor((y - cc.range[2]) / cc.bin_dy));
  const int64_t flat = cc.bin_base + iy * p.EBx + ix;
  const int64_t total = p.L * p.Y * p.X * p.EBy * p.EBx;
  WGRT_CHECK(ix >= 0 && ix <= p.EBx && iy >= 0 && iy <= p.EBy);
  if (flat >= 0 && flat < total) atomicAdd(p.matrix_EB + flat, 1.0f);
}

// Rays whose in-coupling draw picked an order wait here (a per-warp stack) for a free lane: the raw
// ray as loaded, its index, its RNG state after the draw, the order (bit 31 of idx) and the order's
// efficiency; the lane that pops an entry applies the order in phase B.
#ifndef WGRT_QUEUE_CAP
#define WGRT_QUEUE_CAP 36
#endif
constexpr int QUEUE_CAP = WGRT_QUEUE_CAP;
struct alignas(16) QEntry {   // 32 bytes: two 128-bit shared accesses per push / pop
  uint32_t idx, rng;
  float x, y;
  float te, tm, dl, pad_;
};

struct alignas(16) WarpShared {   // one per warp
  CellConst cc;
  QEntry q[QUEUE_CAP];
};

__host__ __device__ constexpr size_t table_offset() { return (sizeof(WarpShared) + 15) & ~size_t(15); }
// rows = 6 + 4 nFC + 6 nOC table rows, of which 2 nOC (the deposit orders) have no Jones matrix
__host__ __device__ constexpr size_t warp_bytes(int rows, int jrows, bool jsm) {
  return table_offset() + static_cast<size_t>(rows) * ROW * sizeof(double) + (jsm ? static_cast<size_t>(jrows) * JROW * sizeof(double) : 0);
}

// Efficiency of an in-coupling order from its quadratic form.  Evaluated twice per surviving ray -- when the order
// is drawn and when a lane pops the ray -- with explicit operations, so that both give the same bits.
__device__ __forceinline__ double incouple_eff(double2 qa, double2 qb, double t2, double m2, double zre, double zim, double g) {
  return __dmul_rn(__dadd_rn(__fma_rn(qa.x, t2, __dmul_rn(qa.y, m2)), __fma_rn(qb.x, zre, __dmul_rn(qb.y, zim))), g);
}

__device__ __forceinline__ const double* lut_slice(const wgrt_problem_t& p, int which, int i, int64_t cell,
                                                   int64_t cells_per_poly, int32_t& C) {
  switch (which) {
    case DIR_IC1: C = p.C_ic; return p.lut_ic1 + 2 * cell * C;
    case DIR_IC2: C = p.C_ic; return p.lut_ic2 + 2 * cell * C;
    case DIR_IC3: C = p.C_ic; return p.lut_ic3 + 2 * cell * C;
    case DIR_FC1: C = p.C_fc; return p.lut_fc1 + 2 * (i * cells_per_poly + cell) * C;
    case DIR_FC2: C = p.C_fc; return p.lut_fc2 + 2 * (i * cells_per_poly + cell) * C;
    case DIR_OC1: C = p.C_oc; return p.lut_oc1 + 2 * (i * cells_per_poly + cell) * C;
    default: C = p.C_oc; return p.lut_oc2 + 2 * (i * cells_per_poly + cell) * C;
  }
}

// The eyebox rectangle of a FoV cell is (xmin,ymax),(xmin,ymin),(xmax,ymin),(xmax,ymax)
// (couplers_coor.py:514-526).  When the four vertices are exactly that, points clearly inside /
// outside it need no edge arithmetic (see deposit_inside).
__device__ __forceinline__ void eyebox_box(const double* __restrict__ r, CellConst& cc) {
  const double x0 = r[0], y0 = r[1], x1 = r[2], y1 = r[3], x2 = r[4], y2 = r[5], x3 = r[6], y3 = r[7];
  const bool ok = x0 == x1 && x2 == x3 && y1 == y2 && y0 == y3 && x0 < x2 && y1 < y0 && isfinite(x0) &&
                  isfinite(x2) && isfinite(y0) && isfinite(y1);
  cc.box[0] = x0; cc.box[1] = x2; cc.box[2] = y1; cc.box[3] = y0;
  cc.box_ok = ok ? 1 : 0;
}

// Event table and per-cell constants of cell (lm, m, n); the 32 lanes of the warp share the rows.
template <bool JSM>
__device__ void build_cell_tables(const wgrt_problem_t& p, int64_t lm, int64_t m, int64_t n, double* tab,
                                  double* __restrict__ jones, CellConst& cc, int rows, int lane) {
  double* jones_sm = tab + rows * ROW;
  const int64_t cell = (lm * p.X + m) * p.Y + n;
  const int64_t cpp = p.L * p.X * p.Y;
  const int nFC = static_cast<int>(p.n_FC), nOC = static_cast<int>(p.n_OC);
  for (int t = lane; t < rows; t += 32) {
    int ev, i, k;
    if (t < 6) {
      ev = t >> 1; i = 0; k = t & 1;
    } else if (t < 6 + 4 * nFC) {
      const int u = t - 6;
      ev = u < 2 * nFC ? EV_S2 : EV_S3;
      const int v = u < 2 * nFC ? u : u - 2 * nFC;
      i = v >> 1; k = v & 1;
    } else {
      const int u = t - 6 - 4 * nFC;
      ev = u < 3 * nOC ? EV_S4 : EV_S5;
      const int v = u < 3 * nOC ? u : u - 3 * nOC;
      i = v / 3; k = v - 3 * i;
    }
    const OrderSpec sp = kOrders[ev][k];
    // Jones row: table rows in order, the deposit orders (k = 2 of the out-coupler events) left out
    const int oc_first = 6 + 4 * nFC;
    const int jrow = t < oc_first ? t : (k == 2 ? -1 : oc_first + 2 * ((t - oc_first) / 3) + k);
    const int src = ev == EV_INIT ? DIR_IC1 : ev == EV_S0 ? DIR_IC2 : ev == EV_S1 ? DIR_IC3
                  : ev == EV_S2 ? DIR_FC1 : ev == EV_S3 ? DIR_FC2 : ev == EV_S4 ? DIR_OC1 : DIR_OC2;
    int32_t C;
    const double* L = lut_slice(p, src, i, cell, cpp, C);
    double* row = tab + t * ROW;
    cplx J[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(L) + sp.ch[q]);
      J[q] = cplx{v.x, v.y};
      if (!JSM) reinterpret_cast<double2*>(jones + t * JROW)[q] = v;
      else if (jrow >= 0) reinterpret_cast<double2*>(jones_sm + jrow * JROW)[q ^ ((jrow >> 1) & 3)] = v;
    }
    int32_t Cd;
    const double* D = lut_slice(p, sp.dir, i, cell, cpp, Cd);
    const double c_new = cos(__ldg(D));  // cos(theta_new.real)
    double f = c_new;
    if (sp.fmode == 1) f = c_new * p.n_g;
    if (sp.fmode == 2) f = c_new / p.n_g;
    // Ete_out = J0 te + J2 tm, Etm_out = J1 te + J3 tm (GRTF:139-144)  =>
    // |Ete_out|^2 + |Etm_out|^2 = M00 |te|^2 + M11 |tm|^2 + 2 Re(M01 conj(te) tm),  M = J^H J
    const double m00 = J[0].re * J[0].re + J[0].im * J[0].im + (J[1].re * J[1].re + J[1].im * J[1].im);
    const double m11 = J[2].re * J[2].re + J[2].im * J[2].im + (J[3].re * J[3].re + J[3].im * J[3].im);
    const double m01re = (J[0].re * J[2].re + J[0].im * J[2].im) + (J[1].re * J[3].re + J[1].im * J[3].im);
    const double m01im = (J[0].re * J[2].im - J[0].im * J[2].re) + (J[1].re * J[3].im - J[1].im * J[3].re);
    row[R_Q + 0] = f * m00;
    row[R_Q + 1] = f * m11;
    row[R_Q + 2] = 2.0 * f * m01re;
    row[R_Q + 3] = -2.0 * f * m01im;
    row[R_INVCOS] = 1.0 / c_new;
    const int nstate = sp.post == POST_IC_FWD ? ST_PEND_FWD : sp.post == POST_IC_BACK ? ST_PEND_BACK : sp.nstate;
    int meta = (sp.tir & 3) | ((sp.gap & 3) << 2) | ((nstate & 7) << 4) | ((sp.post & 3) << 7);
    if (ev >= EV_S4) meta |= META_THREE;
    if (ev >= EV_S2) meta |= META_GATED;
    if (jrow >= 0) meta |= jrow << META_JROW_SHIFT;
    row[R_META] = __hiloint2double(0, meta);
  }
  const int t = lane;
  if (t < 4) {
    const double T = __ldg(p.lut_TIR + 4 * cell + t);
    double s, c;
    sincos(T, &s, &c);
    cc.ph1[t] = cplx{c, s};
    sincos(2.0 * T, &s, &c);
    cc.ph2[t] = cplx{c, s};
  } else if (t < 12) {
    cc.gap[t - 4] = __ldg(p.lut_gap + 8 * cell + (t - 4));
  } else if (t < 20) {
    cc.rect[t - 12] = __ldg(p.eff_reg_FOV + 8 * (m * p.Y + n) + (t - 12));
  } else if (t < 24) {
    cc.range[t - 20] = __ldg(p.eff_reg_FOV_range + 4 * (m * p.Y + n) + (t - 20));
  } else if (t == 24) {
    cc.inv_cos_in = 1.0 / cos(__ldg(p.lut_ic1 + 2 * cell * p.C_ic));
  } else if (t == 26) {
    eyebox_box(p.eff_reg_FOV + 8 * (m * p.Y + n), cc);
  } else if (t == 27) {
    const double* rg = p.eff_reg_FOV_range + 4 * (m * p.Y + n);
    cc.bin_dx = (__ldg(rg + 1) - __ldg(rg + 0)) / static_cast<double>(p.EBx);
    cc.bin_dy = (__ldg(rg + 3) - __ldg(rg + 2)) / static_cast<double>(p.EBy);
    cc.bin_base = ((lm * p.Y + n) * p.X + m) * p.EBy * p.EBx;
  }
}

__device__ __forceinline__ float ld_stream(const float* ptr) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(ptr));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* ptr) {
  uint32_t v;
  asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(ptr));
  return v;
}
__device__ __forceinline__ void st_stream(uint32_t* ptr, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* ptr) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
}

// is_inside_or_on_edge_4d (GRTF:73-108) on the eyebox rectangle.  For an exactly axis-aligned
// rectangle in the runner's vertex order, a point more than 1e-9 inside every side passes the
// literal test (no edge within the 1e-12 tolerance; the crossing test toggles once, on the x = xmax
// edge, because (xj - xi) * ... / ... + xi == xi exactly for the vertical edges) and a point more
// than 1e-9 outside any side fails it (no edge within tolerance; zero or two crossings).  Only
// points within 1e-9 of the boundary run the literal expressions.
template <bool COUNT>
__device__ __forceinline__ bool deposit_inside(const CellConst& cc, double x, double y, Counts* cn) {
  if (cc.box_ok) {
    const double m = 1e-9;
    if (x > cc.box[0] + m && x < cc.box[1] - m && y > cc.box[2] + m && y < cc.box[3] - m) return true;
    if (x < cc.box[0] - m || x > cc.box[1] + m || y < cc.box[2] - m || y > cc.box[3] + m) return false;
  }
  return inside_or_on_edge_literal<COUNT>(x, y, cc.rect, 0, 4, cn);
}

// unit hook (wgrt_debug_deposit_inside): the walk's eyebox test on arbitrary points / rectangles
__global__ void deposit_inside_kernel(const double* __restrict__ rect, const double* __restrict__ px,
                                      const double* __restrict__ py, int64_t n, int32_t* __restrict__ out, int literal) {
  __shared__ CellConst cc;
  if (threadIdx.x < 8) cc.rect[threadIdx.x] = rect[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) eyebox_box(rect, cc);
  __syncthreads();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = literal ? inside_or_on_edge_literal<false>(px[i], py[i], cc.rect, 0, 4, nullptr)
                   : deposit_inside<false>(cc, px[i], py[i], nullptr);
  if (!literal && cc.box_ok) out[i] |= 2;   // bit 1: the shortcut was armed for this rectangle
}

// floor(log2(v)) of a positive normal double (a zero or subnormal gives -1023: "deep in the underflow range")
__device__ __forceinline__ int binary_exponent(double v) { return ((__double2hiint(v) >> 20) & 0x7ff) - 1023; }

struct Ray {
  double x, y;      // position
  cplx te, tm;      // Jones vector, not normalised
  double s;         // 1 / (|te|^2 + |tm|^2)
  double inv_cos;   // 1 / cos(theta_current.real)
  double ener;      // energy left (threshold > 0 launches)
  int esum;         // THR0 launches: sum of the binary exponents of the efficiencies taken so far (ener >= 2^esum)
  uint32_t rng;
  int state;        // region state 0..5, ST_PEND_* after an in-coupler order, ST_DEAD
  int iter;
  int idx;          // ray index relative to the start of the tile
  int row0;         // >= 0: the ray stands on a grating, first table row of the event; < 0: ask the zone tables
};

// One persistent CTA per SM, WALK_WARPS(JSM) independent warps in it.  The warps share only the geometry's zone
// tables (CtaShared); each walks its own tiles with its own cell table, survivor stack and Jones rows.
#ifndef WGRT_JSM_WARPS
#define WGRT_JSM_WARPS 24
#endif
__host__ __device__ constexpr int walk_max_warps(bool jsm) { return jsm ? WGRT_JSM_WARPS : 24; }

// THR0: the launch's energy threshold is 0 (process_rays_kernel_pro_fullColor, GRTF:859).  The gates `ener_k =
// ener * efficiency_k > 0` (GRTF:1017 ff.) then only ask whether the product is positive, i.e. whether the efficiency
// is and nothing underflowed: the walk tracks the binary exponents of the efficiencies taken instead of their product
// (one register and four multiplies per event less, no efficiency to re-evaluate when a survivor is popped), and a ray
// whose exponent sum nears the underflow range -- hundreds of events deep -- goes to the literal walk like a near tie.
template <bool COUNT, bool IMPLICIT, bool JSM, bool THR0>
__global__ void __launch_bounds__(32 * walk_max_warps(JSM), 1)
walk_warp_kernel(const __grid_constant__ wgrt_problem_t p, const __grid_constant__ RegionSet rs,
                 int* __restrict__ work_counter, const int* __restrict__ tile_size_ptr,
                 unsigned long long* counters, double* __restrict__ jones_scratch, RedoList* __restrict__ redo,
                 const double TIE_TOL, const int warp_stride, const int jones_off, const int esum_limit) {
  // (warp_stride = warp_bytes(rows, jrows, JSM), jones_off = rows * ROW: launch constants, so that the step loop
  // re-derives neither from the slice counts)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int rows = 6 + 4 * static_cast<int>(p.n_FC) + 6 * static_cast<int>(p.n_OC);
  const int jrows = rows - 2 * static_cast<int>(p.n_OC);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;

  // ---- the geometry's zone tables: once per CTA ---------------------------------------------------------------
  const bool zone_ok = c_walk.zone.valid != 0;
  const bool level1_sm = zone_ok && c_walk.zone.num_zones <= ZONE_SM_MAX;
  const bool trans_sm = zone_ok && c_walk.zone.num_zones <= TRANS_SM;
  CtaShared& cta = *reinterpret_cast<CtaShared*>(smem_raw);
  if (level1_sm) {
    const uint4* src = reinterpret_cast<const uint4*>(rs.zones.level1);
    for (int i = threadIdx.x; i < ZONE_N1 * ZONE_N1 / 8; i += blockDim.x) {   // 8 ids per 128-bit load
      const uint4 v = __ldg(src + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t out[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t a = w[2 * h], b = w[2 * h + 1];   // ids a.lo, a.hi, b.lo, b.hi; MIXED (0xFFFF) -> 0xFF
        out[h] = (a & 0xffu) | (((a >> 16) & 0xffu) << 8) | ((b & 0xffu) << 16) | (((b >> 16) & 0xffu) << 24);
      }
      reinterpret_cast<uint2*>(cta.level1)[i] = make_uint2(out[0], out[1]);
    }
  }
  if (trans_sm) {
    const int nz = c_walk.zone.num_zones;
    for (int i = threadIdx.x; i < ZONE_STATES * nz; i += blockDim.x) {
      const int st = i / nz, z = i - st * nz;
      cta.trans[st * TRANS_SM + z] = __ldg(rs.zones.trans + st * ZONE_CAP + z);
    }
  }
  __syncthreads();   // the only block barrier: from here on the warps never meet again

  unsigned char* wbase = smem_raw + sizeof(CtaShared) + static_cast<size_t>(warp) * warp_stride;
  WarpShared& sh = *reinterpret_cast<WarpShared*>(wbase);
  double* tab = reinterpret_cast<double*>(wbase + table_offset());
  const double* jones_sm = tab + jones_off;
  Counts cn;
  if (COUNT) cn.clear();

  const int warps = blockDim.x >> 5;
  double* jones = JSM ? nullptr : jones_scratch + (static_cast<size_t>(blockIdx.x) * warps + warp) * rows * JROW;
  const CellConst& cc = sh.cc;
  const int nFC_i = static_cast<int>(p.n_FC), nOC_i = static_cast<int>(p.n_OC);
  const int64_t tile_size = *tile_size_ptr;
  const int64_t num_tiles = (p.num_rays + tile_size - 1) / tile_size;
  const double threshold = p.threshold;
  const bool has_l = p.lmd_num != nullptr;

  // A tile keeps a warp busy for ~1 ms, and a launch is only ~6 tiles per warp: handing out whole tiles to the end
  // would leave the SMs half empty for half a tile (measured: 8 % of the warp slots).  The last half tile per warp is
  // therefore handed out in TAIL_SPLIT pieces (work units past `big_tiles` address piece q % TAIL_SPLIT of tile
  // big_tiles + q / TAIL_SPLIT).  Pieces pay the table build and the end-of-tile drain again, so more or smaller
  // pieces lose (C2: 4 pieces of the last 0.5 / 1 / 1.5 tiles 7.79 / 7.85 / 7.87 ms, 8 of 1.5: 7.96, 16 of 2: 8.96,
  // none: 8.34 ms).
  constexpr int TAIL_SPLIT = 4;
  const int64_t resident = static_cast<int64_t>(gridDim.x) * warps;
  const int64_t split_tiles = tile_size >= 1024 ? min(num_tiles, resident / 2) : 0;
  const int64_t big_tiles = num_tiles - split_tiles;
  const int64_t piece = (((tile_size + TAIL_SPLIT - 1) / TAIL_SPLIT) + 31) & ~int64_t(31);
  const int64_t work_units = big_tiles + TAIL_SPLIT * split_tiles;

  for (;;) {
    int unit_i = 0;
    if (lane == 0) unit_i = atomicAdd(work_counter, 1);
    const int64_t unit = __shfl_sync(FULL_MASK, unit_i, 0);
    if (unit >= work_units) break;
    int64_t t_begin, t_end;
    if (unit < big_tiles) {
      t_begin = unit * tile_size;
      t_end = min(p.num_rays, t_begin + tile_size);
    } else {
      const int64_t q = unit - big_tiles;
      const int64_t base = (big_tiles + q / TAIL_SPLIT) * tile_size;
      t_begin = base + (q % TAIL_SPLIT) * piece;
      t_end = min(min(p.num_rays, base + tile_size), t_begin + piece);
      if (t_begin >= t_end) continue;
    }
    int64_t cursor = t_begin;   // first ray of the tile nobody has staged yet

    while (cursor < t_end) {
      // ---- the run of rays that share the cell of ray `cursor` ---------------------------------
      float km, kn, kl;
      int64_t run_limit;
      int64_t cell_first = 0;   // runner layout: index (in the launch) of ray 0 of the run's cell
      const int64_t run_first = cursor;
      if (IMPLICIT) {
        const int64_t rpc = 2 * p.runner_points;
        const int64_t cell = p.runner_first_cell + cursor / rpc;  // runner order: x outer, y, lambda inner
        cell_first = (cursor / rpc) * rpc;
        kl = static_cast<float>(cell % p.L);
        kn = static_cast<float>((cell / p.L) % p.Y);
        km = static_cast<float>(cell / (p.L * p.Y));
        run_limit = min(t_end, (cursor / rpc + 1) * rpc);
      } else {
        km = __ldg(p.m + cursor); kn = __ldg(p.n + cursor); kl = has_l ? __ldg(p.lmd_num + cursor) : 0.0f;
        run_limit = t_end;   // cut on the fly where the key changes
      }
      const int64_t m = static_cast<int64_t>(km), n = static_cast<int64_t>(kn), lm = static_cast<int64_t>(kl);
      // (a non-finite key converts to an arbitrary integer; such rays are left untouched)
      // (+-inf converts out of range; NaN converts to 0, hence the explicit test)
      const bool valid = km == km && kn == kn && kl == kl && m >= 0 && m < p.X && n >= 0 && n < p.Y && lm >= 0 && lm < p.L;
      __syncwarp();
      if (valid) build_cell_tables<JSM>(p, lm, m, n, tab, jones, sh.cc, rows, lane);
      __syncwarp();

      Ray r;
      r.state = ST_DEAD;
      r.row0 = -1;
      int qn = 0;          // survivors of in-coupling waiting for a lane (warp uniform)
      bool open = true;    // the run may have more rays nobody has in-coupled yet

      for (;;) {
        // ---- in-coupling decisions, 32 rays at a time, every lane busy (GRTF:860-904 up to the
        //      draw): whenever fewer survivors are queued than lanes are free.  About three quarters
        //      of the rays end right here; their position is never even loaded. -----------------
        const unsigned dead = __ballot_sync(FULL_MASK, r.state == ST_DEAD);
        const int nd = __popc(dead);
        while (open && qn < nd && qn <= QUEUE_CAP - 32) {
          const int64_t i = cursor + lane;
          bool same = i < run_limit;
          float fx = 0.f, fy = 0.f, fte = 0.f, ftm = 0.f, fdl = 0.f;
          uint32_t frng = 0u;
          if (same) {
            if (IMPLICIT) {
              // runner layout (RUN:82-115): P TE rays then P TM rays per cell, ray k starts at point k
              const int64_t kk = i - cell_first;   // (= i % (2 P): the run lies inside one cell)
              const bool te_half = kk < p.runner_points;
              const int64_t pt = te_half ? kk : kk - p.runner_points;
              fx = __ldg(p.x + pt); fy = __ldg(p.y + pt);
              fte = te_half ? 1.0f : 0.0f; ftm = te_half ? 0.0f : 1.0f;
              frng = ld_stream(p.rng_states + i);
            } else {
              // (all loads first: a short-circuit chain of volatile loads would pay the memory latency three times)
              const float vm = ld_stream(p.m + i), vn = ld_stream(p.n + i), vl = has_l ? ld_stream(p.lmd_num + i) : kl;
              fx = ld_stream(p.x + i); fy = ld_stream(p.y + i);
              fte = ld_stream(p.te + i); ftm = ld_stream(p.tm + i); fdl = ld_stream(p.delta_phase + i);
              frng = ld_stream(p.rng_states + i);
              same = vm == km && vn == kn && vl == kl;
            }
            if (i + 32 < run_limit) {   // the next batch into L1 (translation included; measured against L2 prefetches
              prefetch_l1(p.rng_states + i + 32);   // two / three batches ahead: -1 %)
              if (!IMPLICIT) {
                prefetch_l1(p.x + i + 32); prefetch_l1(p.y + i + 32);
                prefetch_l1(p.te + i + 32); prefetch_l1(p.tm + i + 32); prefetch_l1(p.delta_phase + i + 32);
                prefetch_l1(p.m + i + 32); prefetch_l1(p.n + i + 32);
                if (has_l) prefetch_l1(p.lmd_num + i + 32);
              }
            }
          }
          const unsigned okmask = __ballot_sync(FULL_MASK, same);
          const int cnt = okmask == FULL_MASK ? 32 : __ffs(~okmask) - 1;   // leading rays of this run
          if (cnt < 32) open = false;
          int k = -1;
          int ex0 = 0;   // THR0: binary exponent of the chosen in-coupling efficiency
          if (lane < cnt && valid) {   // (rays of a run whose cell indices are out of range are left untouched)
            if (COUNT) {
              cn.c[WGRT_CNT_RAYS]++; cn.c[WGRT_CNT_DRAWS]++; cn.c[WGRT_CNT_DRAW2]++; cn.c[WGRT_CNT_EFIELD] += 2;
            }
            const double u = xorshift_draw(frng, p.ray_index_base + i);
            const double te = static_cast<double>(fte), tm = static_cast<double>(ftm);
            cplx w{tm, 0.0};
            if (fdl != 0.0f) {
              double sn, cs;
              sincos(static_cast<double>(fdl), &sn, &cs);
              w = cplx{tm * cs, tm * sn};
            }
            // GRTF:860-869: the raw amplitudes enter E_field_cal as they are (s = 1)
            const double t2 = te * te, m2 = w.re * w.re + w.im * w.im, zre = te * w.re, zim = te * w.im;
            const double g = cc.inv_cos_in;
            const double2* q = reinterpret_cast<const double2*>(tab);
            const double2 a0 = q[0], a1 = q[1], b0 = q[ROW / 2], b1 = q[ROW / 2 + 1];
            const double e1 = incouple_eff(a0, a1, t2, m2, zre, zim, g);
            const double e2 = incouple_eff(b0, b1, t2, m2, zre, zim, g);
            if ((fabs(u - e1) < TIE_TOL || fabs(u - (e1 + e2)) < TIE_TOL) && redo_push(redo, i)) {
              // near tie: left untouched for the literal re-walk
            } else if (u <= e1) { k = 0; ex0 = binary_exponent(e1); }             // GRTF:871: no energy gate here
            else if (u <= e1 + e2) { k = 1; ex0 = binary_exponent(e2); }          // GRTF:887
            else st_stream(p.rng_states + i, frng);            // GRTF:903-904: absorbed
          }
          const unsigned surv = __ballot_sync(FULL_MASK, k >= 0);
          if (k >= 0) {
            const int slot = qn + __popc(surv & lt_mask);
            WGRT_CHECK(slot >= 0 && slot < QUEUE_CAP && i >= t_begin && i < t_end);
            uint4* qe = reinterpret_cast<uint4*>(&sh.q[slot]);
            const uint32_t idxk = static_cast<uint32_t>(i - t_begin) | (static_cast<uint32_t>(k) << 31);
            qe[0] = make_uint4(idxk, frng, __float_as_uint(fx), __float_as_uint(fy));
            qe[1] = make_uint4(__float_as_uint(fte), __float_as_uint(ftm), __float_as_uint(fdl), static_cast<uint32_t>(ex0));
          }
          qn += __popc(surv);
          // (A NaN cell key never compares equal, not even to itself: the run of such a ray would be empty and the
          // cursor would never advance.  The ray is outside every table: skip it, untouched.)
          cursor += (cnt == 0 && cursor == run_first) ? 1 : cnt;
          __syncwarp();
        }

        // ---- free lanes pop survivors: re-read the ray; phase B applies the order already chosen --
        if (nd && qn) {
          const int take = min(nd, qn);
          const int rank = __popc(dead & lt_mask);
          if (r.state == ST_DEAD && rank < take) {
            const int slot = qn - 1 - rank;
            WGRT_CHECK(slot >= 0 && slot < QUEUE_CAP);
            const uint4* qe = reinterpret_cast<const uint4*>(&sh.q[slot]);
            const uint4 q0 = qe[0], q1 = qe[1];
            const float fdl = __uint_as_float(q1.z);
            r.idx = static_cast<int>(q0.x & 0x7fffffffu);
            r.rng = q0.y;
            r.x = static_cast<double>(__uint_as_float(q0.z));
            r.y = static_cast<double>(__uint_as_float(q0.w));
            const double te = static_cast<double>(__uint_as_float(q1.x)), tm = static_cast<double>(__uint_as_float(q1.y));
            r.te = cplx{te, 0.0};
            r.tm = cplx{tm, 0.0};
            if (fdl != 0.0f) {
              double sn, cs;
              sincos(static_cast<double>(fdl), &sn, &cs);
              r.tm = cplx{tm * cs, tm * sn};
            }
            r.row0 = static_cast<int>(q0.x >> 31);   // the chosen in-coupling row (0 or 1)
            if (THR0) {
              r.esum = static_cast<int>(q1.w);
            } else {   // ener = 1 * efficiency of the in-coupled order (GRTF:882), the very value the draw was compared with
              const double2* q = reinterpret_cast<const double2*>(tab + r.row0 * ROW);
              r.ener = incouple_eff(q[0], q[1], te * te, r.tm.re * r.tm.re + r.tm.im * r.tm.im, te * r.tm.re, te * r.tm.im,
                                    cc.inv_cos_in);
            }
            r.s = 1.0;
            r.inv_cos = cc.inv_cos_in;
            r.iter = -1;                     // marks "order already chosen"
            r.state = 0;
          }
          qn -= take;
          __syncwarp();
        }
        if (__ballot_sync(FULL_MASK, r.state != ST_DEAD) == 0u) {
          if (!open && qn == 0) break;
          continue;
        }
        if (COUNT && lane == 0) cn.c[WGRT_CNT_WARP_STEPS]++;
        bool lost = false;

        // ---- phase B1: diffract -- draw and pick the order ------------------------------------------
        const bool at_event = r.state != ST_DEAD && r.row0 >= 0;
        int k = -1;
        double esel = 1.0;
        if (at_event) {
          if (r.iter >= 0) {
            WGRT_CHECK(r.row0 >= 0 && r.row0 + 1 < rows && t_begin + r.idx < t_end);
            const double2* e = reinterpret_cast<const double2*>(tab + r.row0 * ROW);
            const double2 a0 = e[0], a1 = e[1], a2 = e[2];
            const double2 b0 = e[ROW / 2], b1 = e[ROW / 2 + 1];
            const int meta0 = __double2loint(a2.y);
            const bool three = (meta0 & META_THREE) != 0;
            const bool gated = (meta0 & META_GATED) != 0;
            WGRT_CHECK(!three || r.row0 + 2 < rows);
            double2 c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);
            if (three) { c0 = e[ROW]; c1 = e[ROW + 1]; }
            const double u = xorshift_draw(r.rng, p.ray_index_base + t_begin + r.idx);
            if (COUNT) {
              cn.c[WGRT_CNT_DRAWS]++;
              cn.c[three ? WGRT_CNT_DRAW3 : WGRT_CNT_DRAW2]++;
              cn.c[WGRT_CNT_EFIELD] += three ? 3 : 2;
            }
            const double t2 = r.te.re * r.te.re + r.te.im * r.te.im;
            const double m2 = r.tm.re * r.tm.re + r.tm.im * r.tm.im;
            const double zre = r.te.re * r.tm.re + r.te.im * r.tm.im;   // conj(te) * tm
            const double zim = r.te.re * r.tm.im - r.te.im * r.tm.re;
            const double g = r.s * r.inv_cos;
            const double e1 = (a0.x * t2 + a0.y * m2 + (a1.x * zre + a1.y * zim)) * g;
            const double e2 = (b0.x * t2 + b0.y * m2 + (b1.x * zre + b1.y * zim)) * g;
            const double e3v = (c0.x * t2 + c0.y * m2 + (c1.x * zre + c1.y * zim)) * g;   // two-order events: never selected
            // the reference's if / elif chain (GRTF:919-953, 1020-1048, 1135-1174)
            // (evaluated without branches: the three tests are cheap, a divergent chain is not)
            const double e12 = e1 + e2;
            const bool g1 = THR0 ? e1 > 0.0 : r.ener * e1 > threshold;
            const bool g2 = THR0 ? e2 > 0.0 : r.ener * e2 > threshold;
            const bool g3 = THR0 ? e3v > 0.0 : r.ener * e3v > threshold;
            const bool ok1 = u <= e1 && (!gated || g1);
            const bool ok2 = u <= e12 && (!gated || g2);
            const bool ok3 = three && u <= e12 + e3v && g3;
            k = ok1 ? 0 : ok2 ? 1 : ok3 ? 2 : -1;
            esel = ok1 ? e1 : ok2 ? e2 : e3v;
            bool tie = fabs(u - e1) < TIE_TOL || fabs(u - e12) < TIE_TOL || (three && fabs(u - (e12 + e3v)) < TIE_TOL);
            if (THR0) {   // positive efficiency <=> positive product, unless the product may have underflowed
              if (gated && (r.esum < esum_limit || (k >= 0 && esel < 1e-30))) tie = true;
            } else if (threshold > 0.0 && gated) {   // the energy gates of the single-wavelength twin (GRTF:444)
              const double rt = 1e-9 * threshold;
              tie = tie || fabs(r.ener * e1 - threshold) < rt || fabs(r.ener * e2 - threshold) < rt ||
                    (three && fabs(r.ener * e3v - threshold) < rt);
            }
            if (tie && redo_push(redo, t_begin + r.idx)) k = -2;   // dropped untouched, re-walked literally
          } else {
            r.iter = 0;   // a popped survivor: row0 IS the chosen in-coupling row, ener already holds its efficiency
            k = 0;
          }
        }
        __syncwarp();   // one copy of the code below, run by every lane that has an order to apply

        // ---- phase B2: apply the chosen order -----------------------------------------------------------
        if (at_event) {
          if (k < 0) {
            lost = true;   // absorbed
            if (k == -2) {   // near tie: not even the RNG state is written
              r.state = ST_DEAD;
              lost = false;
            }
          } else {
            WGRT_CHECK(r.row0 >= 0 && r.row0 + k < rows && k <= 2);
            const double2* row = reinterpret_cast<const double2*>(tab + (r.row0 + k) * ROW);
            const double2 im = row[R_INVCOS / 2];   // {1 / cos, meta}
            const int meta = __double2loint(im.y);
            const int post = (meta >> 7) & 3;
            if (post == POST_DEPOSIT) {
              // GRTF:1162-1171: count the ray if it leaves inside this FoV's eyebox rectangle
              if (deposit_inside<COUNT>(cc, r.x, r.y, &cn)) {
                deposit_bin_cell(p, cc, r.x, r.y);
                if (COUNT) cn.c[WGRT_CNT_DEPOSITS]++;
              }
              lost = true;
            } else {
              // apply the chosen order's Jones matrix (GRTF:139-144)
              const int jrow = JSM ? (meta >> META_JROW_SHIFT) & 0xffff : r.row0 + k;
              const int sw = JSM ? (jrow >> 1) & 3 : 0;
              const double2* jr = reinterpret_cast<const double2*>((JSM ? jones_sm : jones) + jrow * JROW);
              const double2 j0 = jr[sw], j1 = jr[1 ^ sw], j2 = jr[2 ^ sw], j3 = jr[3 ^ sw];
              const cplx L0{j0.x, j0.y}, L1{j1.x, j1.y}, L2{j2.x, j2.y}, L3{j3.x, j3.y};
              cplx nte{L0.re * r.te.re - L0.im * r.te.im + (L2.re * r.tm.re - L2.im * r.tm.im),
                       L0.re * r.te.im + L0.im * r.te.re + (L2.re * r.tm.im + L2.im * r.tm.re)};
              cplx ntm{L1.re * r.te.re - L1.im * r.te.im + (L3.re * r.tm.re - L3.im * r.tm.im),
                       L1.re * r.te.im + L1.im * r.te.re + (L3.re * r.tm.im + L3.im * r.tm.re)};
              double nt2 = nte.re * nte.re + nte.im * nte.im;
              double nm2 = ntm.re * ntm.re + ntm.im * ntm.im;
              const double eps2 = 1e-40;
              double n2 = nt2 + nm2;
              if (fmin(nt2, nm2) * r.s < eps2 || n2 < 1e-200) {   // rare
                if (nt2 * r.s < eps2) nte = cplx{sqrt(nt2), 0.0};
                if (nm2 * r.s < eps2) ntm = cplx{sqrt(nm2), 0.0};
                if (n2 < 1e-200) {   // exact rescaling by a power of two: no rounding anywhere
                  const double k2 = 0x1p332, k4 = 0x1p664;
                  nte.re *= k2; nte.im *= k2; ntm.re *= k2; ntm.im *= k2;
                  n2 *= k4;
                }
              }
              r.te = nte;
              r.tm = cmul(ntm, cc.ph1[meta & 3]);         // delta += T[tir] (GRTF:878 ff.)
              r.s = __drcp_rn(n2);
              const int gp = (meta >> 2) & 3;
              r.x += cc.gap[2 * gp];
              r.y += cc.gap[2 * gp + 1];
              r.inv_cos = im.x;
              if (THR0) r.esum += binary_exponent(esel);
              else r.ener *= esel;
              r.state = (meta >> 4) & 7;   // region state, or "pending" after an in-coupler order
              if (COUNT) cn.c[WGRT_CNT_BOUNCES]++;
            }
          }
          r.row0 = -1;
          if (lost) {
            WGRT_CHECK(r.idx >= 0 && t_begin + r.idx < t_end);
            st_stream(p.rng_states + t_begin + r.idx, r.rng);
            r.state = ST_DEAD;
            lost = false;
          }
        }
        __syncwarp();

        // ---- phase A: go to the next grating.  Every lane whose ray moved looks up its zone and the
        //      transition table entry of (region state, zone): one 16-bit id from the CTA's shared copy of the
        //      level-1 grid (a second, finer level in global memory under MIXED cells) and one 32-bit word say
        //      what the loop head does with the ray. ----
        if (r.state != ST_DEAD && r.row0 < 0) {
          uint32_t act = ACT_RESOLVE;
          int z = -1;
          if (trans_sm) {   // (implies level1_sm: the design's zone tables are in shared memory -- one test on the usual path)
            z = zone_lookup_walk<true>(cta.level1, nullptr, rs.zones.level2, r.x, r.y);
            WGRT_CHECK(z >= 0 && z < TRANS_SM && r.state >= 0 && r.state < ZONE_STATES);
            act = cta.trans[r.state * TRANS_SM + z];
          } else if (zone_ok) {
            z = level1_sm ? zone_lookup_walk<true>(cta.level1, nullptr, rs.zones.level2, r.x, r.y)
                          : zone_lookup_walk<false>(nullptr, rs.zones.level1, rs.zones.level2, r.x, r.y);
            WGRT_CHECK(z >= 0 && z < ZONE_CAP && r.state >= 0 && r.state < ZONE_STATES);
            act = __ldg(rs.zones.trans + r.state * ZONE_CAP + z);
          }
          // rare: a field this state needs is MIXED in the zone (per-set grids / literal edges decide), or the
          // design has too many zones for the table (word atlas + decode on the fly)
          if (act & ACT_RESOLVE) act = resolve_transition<COUNT>(rs, z, r.state, r.x, r.y, nFC_i, nOC_i, &cn);
          const int inc = (act >> 8) & 3;
          if (COUNT) cn.c[WGRT_CNT_ITERS] += (inc == 2 && r.iter + 1 > 100000) ? 1 : inc;
          r.iter += inc;
          lost = (act & ACT_LOST) != 0 || r.iter > 100000;   // GRTF:905: at most 100000 iterations
          r.state = static_cast<int>(act & 7u);
          if (!lost) {
            if (act & ACT_EVENT) {
              r.row0 = static_cast<int>((act >> 12) & 0xffffu);
              WGRT_CHECK(r.row0 >= 2 && r.row0 + 1 < rows);
            } else {
              const int g = (act >> 5) & 3;
              r.x += cc.gap[2 * g];
              r.y += cc.gap[2 * g + 1];
              r.tm = cmul(r.tm, cc.ph2[(act >> 7) & 1]);
              if (COUNT) cn.c[WGRT_CNT_BOUNCES]++;
            }
          } else {
            WGRT_CHECK(r.idx >= 0 && t_begin + r.idx < t_end);
            st_stream(p.rng_states + t_begin + r.idx, r.rng);
            r.state = ST_DEAD;
          }
        }
        __syncwarp();
      }
    }
  }
  if (COUNT) cn.flush(counters);
}

// ---------------------------------------------------------------------------------------------
// tile size: a tile should hold whole runs.  One block measures the first run.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pick_tile_warp_kernel(const __grid_constant__ wgrt_problem_t p, int* tile_size,
                                                              int* work_counter, int tile_cap, RedoList* redo) {
  __shared__ int s_run;
  if (threadIdx.x == 0) {
    *work_counter = 0;
    redo->count = 0u;
    s_run = INT_MAX;
  }
  if (p.tile_hint) {
    if (threadIdx.x == 0) *tile_size = static_cast<int>(p.tile_hint);
    return;
  }
  if (p.runner_points > 0) {
    if (threadIdx.x == 0) s_run = static_cast<int>(2 * p.runner_points < (1 << 16) ? 2 * p.runner_points : (1 << 16));
  }
  __syncthreads();
  const int limit = static_cast<int>(p.num_rays < (1 << 16) ? p.num_rays : (1 << 16));
  if (p.runner_points == 0) {
    const bool has_l = p.lmd_num != nullptr;
    const float km = p.m[0], kn = p.n[0], kl = has_l ? p.lmd_num[0] : 0.0f;
    for (int i = 1 + threadIdx.x; i < limit; i += blockDim.x) {
      if (p.m[i] != km || p.n[i] != kn || (has_l && p.lmd_num[i] != kl)) {
        atomicMin(&s_run, i);
        break;  // later indices of this thread are larger
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int64_t run = s_run == INT_MAX ? limit : s_run;
    // a warp walks a tile alone.  Whole runs when they fit; long runs in equal pieces; short runs
    // grouped.  `tile_cap` = rays of the launch / (4 x resident warps): small launches (pipeline
    // chunks, multi-GPU shards) cut their cells into pieces so that every resident warp gets several
    // tiles, but not below ~1250 rays (the drain at the end of a tile is paid per tile) -- unless the launch is so
    // small that 1250-ray tiles would leave resident warps without any (then: one tile per warp, at least 128 rays)
    const int64_t per_warp = 4 * static_cast<int64_t>(tile_cap);
    const int64_t floor_t = per_warp >= 1250 ? 1250 : (per_warp < 128 ? 128 : per_warp);
    const int64_t t_max = tile_cap < floor_t ? floor_t : (tile_cap > 8192 ? 8192 : tile_cap);
    const int64_t t_min = t_max / 2;
    int64_t t;
    if (run > t_max) {
      const int64_t pieces = (run + t_max - 1) / t_max;
      t = (run + pieces - 1) / pieces;
    } else if (run >= t_min) {
      t = run;
    } else {
      t = run * ((t_min + run - 1) / run);
    }
    *tile_size = static_cast<int>(t > 32 ? t : 32);
  }
}

}  // namespace

cudaError_t launch_debug_deposit_inside(const double* rect, const double* px, const double* py, int64_t n, int32_t* out,
                                        int literal, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  deposit_inside_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, s>>>(rect, px, py, n, out, literal);
  return cudaGetLastError();
}

namespace {
__global__ void zone_trans_kernel(const __grid_constant__ RegionSet rs, int nFC, int nOC) {
  if (!*rs.dirty) return;
  const ZoneDyn zd = *rs.zones.dyn;
  if (!zd.valid) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ZONE_STATES * zd.num_zones) return;
  const int st = t / zd.num_zones, z = t - st * zd.num_zones;
  rs.zones.trans[st * ZONE_CAP + z] = decode_transition(rs.zones.words[z], st, nFC, nOC);
}
}  // namespace

cudaError_t launch_zone_transitions(const RegionSet& rs, int n_FC, int n_OC, cudaStream_t s) {
  zone_trans_kernel<<<(ZONE_STATES * ZONE_CAP + 255) / 256, 256, 0, s>>>(rs, n_FC, n_OC);
  return cudaGetLastError();
}

cudaError_t walk_check_failures(unsigned long long* out, bool reset) {
#if defined(WGRT_CHECKED)
  cudaError_t e = cudaMemcpyFromSymbol(out, wgrt_check_fail_count, sizeof(unsigned long long));
  if (e == cudaSuccess && reset) {
    const unsigned long long zero = 0;
    e = cudaMemcpyToSymbol(wgrt_check_fail_count, &zero, sizeof zero);
  }
  return e;
#else
  (void)out; (void)reset;
  return cudaErrorNotSupported;
#endif
}

void set_tie_tolerance(double tol) { g_tie_tol = tol >= 0.0 ? tol : TIE_TOL_DEFAULT; }

size_t walk_warp_scratch_bytes(const wgrt_problem_t& p, int num_sms) {
  const int rows = 6 + 4 * static_cast<int>(p.n_FC) + 6 * static_cast<int>(p.n_OC);
  return static_cast<size_t>(num_sms) * 32 * rows * JROW * sizeof(double);   // at most 32 warps per SM
}

namespace {
int env_int(const char* name, int fallback) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : fallback;
}
}  // namespace

cudaError_t launch_walk_warp(const wgrt_problem_t& p, const RegionSet& rs, int* work_counter,
                             unsigned long long* counters, int num_sms, double* jones_scratch, RedoList* redo,
                             cudaStream_t s) {
  if (p.num_rays == 0) return cudaSuccess;
  int* tile_size = work_counter + 1;  // workspace layout: {tile counter, tile size}
  const int rows = 6 + 4 * static_cast<int>(p.n_FC) + 6 * static_cast<int>(p.n_OC);
  const bool count = (p.flags & WGRT_FLAG_COUNTERS) != 0;
  const bool implicit = p.runner_points > 0;
  // One CTA per SM; its shared memory holds the geometry's zone tables once and a cell table per warp.  With the
  // Jones rows in the table (JSM) every per-step access of the walk is a shared-memory access; that form is used
  // when at least 16 warps of it fit beside the zone tables (up to ~110 event rows: BASELINE's designs have 70), the
  // form with the Jones rows in a global scratch otherwise.  WGRT_WALK_JSM / WGRT_WALK_WARPS override (experiments).
  const size_t max_smem = 227 * 1024;
  const size_t avail = max_smem - sizeof(CtaShared);
  static const int env_jsm = env_int("WGRT_WALK_JSM", -1), env_warps = env_int("WGRT_WALK_WARPS", 0);
  const int jrows = rows - 2 * static_cast<int>(p.n_OC);
  const int fit_jsm = static_cast<int>(avail / warp_bytes(rows, jrows, true));
  const bool jsm = env_jsm >= 0 ? env_jsm != 0 : fit_jsm >= 16;
  int warps = static_cast<int>(avail / warp_bytes(rows, jrows, jsm));
  if (warps > walk_max_warps(jsm)) warps = walk_max_warps(jsm);
  if (env_warps > 0 && env_warps < warps) warps = env_warps;
  if (warps < 1) return cudaErrorInvalidValue;   // more event rows than one warp's table can hold
  const size_t smem = sizeof(CtaShared) + static_cast<size_t>(warps) * warp_bytes(rows, jrows, jsm);
  typedef void (*Kern)(const wgrt_problem_t, const RegionSet, int*, const int*, unsigned long long*, double*, RedoList*,
                       const double, const int, const int, const int);
#define WGRT_WALK_ROW(C, I) walk_warp_kernel<C, I, false, false>, walk_warp_kernel<C, I, false, true>, \
                           walk_warp_kernel<C, I, true, false>, walk_warp_kernel<C, I, true, true>
  const Kern table[16] = {WGRT_WALK_ROW(false, false), WGRT_WALK_ROW(false, true), WGRT_WALK_ROW(true, false),
                          WGRT_WALK_ROW(true, true)};
#undef WGRT_WALK_ROW
  const bool thr0 = p.threshold == 0.0;
  const Kern kern = table[(count ? 8 : 0) + (implicit ? 4 : 0) + (jsm ? 2 : 0) + (thr0 ? 1 : 0)];
  // Function attributes cost tens of microseconds per call: set once per (device, kernel variant, shared-memory size)
  // -- a launch of a small problem (a pipeline chunk, BASELINE config 1) is otherwise dominated by them.
  struct Setup { const void* kern; size_t smem; int device; };
  static Setup cache[16];
  static int cached = 0;
  int device = 0;
  cudaError_t err = cudaGetDevice(&device);
  if (err != cudaSuccess) return err;
  bool known = false;
  for (int i = 0; i < cached; ++i)
    if (cache[i].kern == reinterpret_cast<const void*>(kern) && cache[i].smem == smem && cache[i].device == device) known = true;
  if (!known) {
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    // Shared memory and L1 share the SM's 256 KB and the split comes in steps (100 / 132 / 164 / 196 / 228 KB of shared
    // memory): the smallest step that holds the CTA leaves the most L1 (level-2 zone cells, Jones scratch, ray streams).
    int carve = 100;
    const int steps_kb[5] = {100, 132, 164, 196, 228};
    for (int k = 0; k < 5; ++k)
      if (static_cast<size_t>(steps_kb[k]) * 1024 >= smem + 1024) { carve = (steps_kb[k] * 100 + 227) / 228; break; }
    carve = env_int("WGRT_SMEM_CARVEOUT", carve);
    err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    if (err != cudaSuccess) return err;
    cache[cached % 16] = Setup{reinterpret_cast<const void*>(kern), smem, device};
    if (cached < 16) ++cached;
  }
  const int64_t min_tiles = (p.num_rays + 31) / 32;
  const int64_t want_ctas = (min_tiles + warps - 1) / warps;
  const int grid = static_cast<int>(want_ctas < num_sms ? (want_ctas > 1 ? want_ctas : 1) : num_sms);
  const int64_t resident = static_cast<int64_t>(grid) * warps;
  const int64_t tcap = p.num_rays / (4 * resident);
  pick_tile_warp_kernel<<<1, 1024, 0, s>>>(p, tile_size, work_counter, static_cast<int>(tcap > (1 << 20) ? (1 << 20) : tcap), redo);
  err = cudaMemcpyToSymbolAsync(c_walk, rs.atlas_dyn, sizeof(AtlasDyn), offsetof(WalkConst, atlas), cudaMemcpyDeviceToDevice, s);
  if (err != cudaSuccess) return err;
  err = cudaMemcpyToSymbolAsync(c_walk, rs.zones.dyn, sizeof(ZoneDyn), offsetof(WalkConst, zone), cudaMemcpyDeviceToDevice, s);
  if (err != cudaSuccess) return err;
  if (env_int("WGRT_DEBUG_ZONES", 0)) {
    ZoneDyn zd;
    cudaStreamSynchronize(s);
    cudaMemcpy(&zd, rs.zones.dyn, sizeof zd, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[wgrt] zones: num=%d valid=%d outside=%d rows=%d warps=%d jsm=%d smem=%zu\n", zd.num_zones, zd.valid,
            zd.outside_zone, rows, warps, jsm ? 1 : 0, smem);
  }
  kern<<<grid, 32 * warps, smem, s>>>(p, rs, work_counter, tile_size, counters, jones_scratch, redo, g_tie_tol,
                                      static_cast<int>(warp_bytes(rows, jrows, jsm)), rows * ROW,
                                      g_tie_tol > 1e-6 ? -4 : -900);   // (a widened tie tolerance -- tests -- also
                                                                       // exercises the underflow guard of THR0)
  err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  return launch_walk_redo(p, redo, counters, s);   // the near-tie rays, literally (usually none)
}

}  // namespace wgrt
