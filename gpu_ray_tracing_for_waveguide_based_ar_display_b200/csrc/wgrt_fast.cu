// wgrt_fast.cu -- the region-index / atlas builder, the evaluation kernels, and the round's FIRST
// fast walk (CTA per cell).  The production walk is wgrt_walk.cu (warp per cell); the kernel below
// stays selectable with WGRT_WALK=cta for A/B measurements and runs the same parity suite.
//
// Same Monte-Carlo walk as process_rays_kernel_pro_fullColor (GRTF:833-1246), re-designed:
//
//  * work = tiles of consecutive rays; persistent CTAs pull tiles from a global counter.  Inside a
//    tile, maximal runs of rays that share one (FoV-x, FoV-y, wavelength) cell are walked together
//    (the runner lays rays out cell by cell, gpu_ray_tracing_pro_fullColor.py:82-115).
//  * per run, the CTA gathers everything the cell needs into shared memory ONCE: for every event
//    (in-coupling, the two in-coupler states, fold-coupler slice i x 2 states, out-coupler slice i
//    x 2 states) and every diffraction order the Jones quartet, the cos-ratio factor, the new
//    direction's 1/cos, and which TIR phase / bounce vector / next state the order leads to.  The
//    walk itself is then table driven and identical for all states: no per-state code paths, no
//    global LUT gathers, no transcendental in the loop.
//  * the polarisation state is carried as the Jones vector (a, w) = (|E_te|, |E_tm| e^{i delta})
//    instead of (|E_te|, |E_tm|, delta): E_field_cal's cos/sin/hypot/atan2 (GRTF:136-150) and the
//    TIR phase additions become complex multiplies by per-cell phasors e^{iT}, e^{2iT}.  This is
//    algebraically the same map; results differ from the literal evaluation by a few ulp, i.e. a
//    decision `u <= efficiency` can flip only if u lands within ~1e-15 of the threshold.
//  * point-in-region tests go through the cell grids of wgrt_region.cuh (exactly equivalent).
//  * lanes whose ray terminated are refilled from the run's queue with a warp-aggregated
//    (ballot + popc prefix) fetch, so warps stay full although path lengths are heavy tailed.
//  * per-ray RNG state lives in a register and is stored once; bins get one red.global.add.f32.
#include "wgrt_region.cuh"

namespace wgrt {

namespace {

#ifndef WGRT_WALK_THREADS
#define WGRT_WALK_THREADS 128
#endif
#ifndef WGRT_WALK_MIN_BLOCKS
#define WGRT_WALK_MIN_BLOCKS 5
#endif
constexpr int WALK_THREADS = WGRT_WALK_THREADS;
constexpr int ENTRY_DOUBLES = 12;  // 8 Jones + factor + inv_cos_new + meta + spare
constexpr int ST_DEAD = -1;
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

struct alignas(16) CellConst {
  cplx ph1[4];      // e^{i T[k]}
  cplx ph2[4];      // e^{i 2 T[k]}
  double gap[8];    // lut_gap[lm, m, n, :]
  double rect[8];   // eff_reg_FOV[m, n, :, :]
  double range[4];  // eff_reg_FOV_range[m, n, :]
  double inv_cos_in;
  int sinfo[8];     // per region state 0..5: see SI_* below
};

// sinfo bit layout: which region set decides the event, where the state's rows start, how many rows
// per coupler slice, whether the walk loop's effective-region test applies, what a miss means.
enum : int {
  SI_REGION_MASK = 7, SI_REGION_NONE = 7,       // bits 0-2
  SI_ROWBASE_SHIFT = 3, SI_ROWBASE_MASK = 0xfff,   // bits 3-14
  SI_STRIDE_SHIFT = 15,                            // bits 15-16
  SI_MISS_SHIFT = 17,                              // bits 17-18: 0 bounce on, 1 test eff_reg2 first, 2 lost
  SI_PHASE_SHIFT = 19,                             // bit 19: which doubled TIR phase a free bounce adds
  SI_GAP_SHIFT = 20                                // bits 20-21: lut_gap pair of the state's direction of travel
};
__host__ __device__ constexpr int make_sinfo(int region, int rowbase, int stride, int miss, int phase, int gap) {
  return region | (rowbase << SI_ROWBASE_SHIFT) | (stride << SI_STRIDE_SHIFT) | (miss << SI_MISS_SHIFT) |
         (phase << SI_PHASE_SHIFT) | (gap << SI_GAP_SHIFT);
}
constexpr long long META_THREE = 1ll << 9;   // on the first row of an event: the event has three orders
constexpr long long META_GATED = 1ll << 10;  // ... and its branches carry `and ener_k > threshold`

struct WalkShared {
  Region reg[NUM_REGIONS];
  CellConst cc;
  int tile;        // current tile index
  int run_end;     // end of the current run (ray index, exclusive)
  int run_cursor;  // next ray of the run nobody has in-coupled yet (warps claim 32 at a time)
};

struct WarpQueue;
__host__ __device__ constexpr size_t walk_smem_table_offset() { return (sizeof(WalkShared) + 15) & ~size_t(15); }
__host__ __device__ inline size_t walk_smem_queue_offset(int rows) {
  return walk_smem_table_offset() + ((static_cast<size_t>(rows) * ENTRY_DOUBLES * sizeof(double) + 15) & ~size_t(15));
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

// Fill the event table and the per-cell constants for cell (lm, m, n).  All threads participate.
__device__ void build_cell_tables(const wgrt_problem_t& p, int64_t lm, int64_t m, int64_t n, double* tab,
                                  CellConst& cc, int rows) {
  const int64_t cell = (lm * p.X + m) * p.Y + n;
  const int64_t cpp = p.L * p.X * p.Y;
  const int nFC = static_cast<int>(p.n_FC), nOC = static_cast<int>(p.n_OC);
  for (int t = threadIdx.x; t < rows; t += blockDim.x) {
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
    // Jones source table of this event
    const int src = ev == EV_INIT ? DIR_IC1 : ev == EV_S0 ? DIR_IC2 : ev == EV_S1 ? DIR_IC3
                  : ev == EV_S2 ? DIR_FC1 : ev == EV_S3 ? DIR_FC2 : ev == EV_S4 ? DIR_OC1 : DIR_OC2;
    int32_t C;
    const double* L = lut_slice(p, src, i, cell, cpp, C);
    double* row = tab + t * ENTRY_DOUBLES;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(L) + sp.ch[q]);
      row[2 * q] = v.x;
      row[2 * q + 1] = v.y;
    }
    int32_t Cd;
    const double* D = lut_slice(p, sp.dir, i, cell, cpp, Cd);
    const double c_new = cos(__ldg(D));  // cos(theta_new.real)
    double f = c_new;
    if (sp.fmode == 1) f = c_new * p.n_g;
    if (sp.fmode == 2) f = c_new / p.n_g;
    row[8] = f;
    row[9] = 1.0 / c_new;
    long long meta = (sp.tir & 3) | ((sp.gap & 3) << 2) | ((sp.nstate & 7) << 4) | ((sp.post & 3) << 7);
    if (ev >= EV_S4) meta |= META_THREE;
    if (ev >= EV_S2) meta |= META_GATED;
    row[10] = __longlong_as_double(meta);
    row[11] = 0.0;
  }
  const int t = threadIdx.x;
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
  } else if (t == 25) {
    // the bounce vector a ray travels with is a function of its state: states 0 and 2 follow the
    // +1 in-coupled direction (gap pair 0), state 1 the -1 direction (2), states 3 and 4 the folded
    // direction (1), state 5 the conjugate out-coupler direction (3)
    cc.sinfo[0] = make_sinfo(SI_REGION_NONE, 2, 0, 0, 0, 0);
    cc.sinfo[1] = make_sinfo(SI_REGION_NONE, 4, 0, 0, 0, 2);
    cc.sinfo[2] = make_sinfo(REG_FC, 6, 2, 0, 0, 0);                         // miss: bounce, 2 T[0]
    cc.sinfo[3] = make_sinfo(REG_FC, 6 + 2 * nFC, 2, 1, 1, 1);               // miss: eff_reg2 test, 2 T[1]
    cc.sinfo[4] = make_sinfo(REG_OC, 6 + 4 * nFC, 3, 0, 1, 1);               // miss: bounce, 2 T[1]
    cc.sinfo[5] = make_sinfo(REG_OC, 6 + 4 * nFC + 3 * nOC, 3, 2, 0, 3);     // miss: lost
    cc.sinfo[6] = 0;
    cc.sinfo[7] = 0;
  }
}

// Ray SoA data is touched once per launch: keep it out of L1 (which the region grids and LUT
// gathers want) and mark it evict-first in L2.
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
__device__ __forceinline__ void prefetch_l2(const void* ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

struct Ray {
  double x, y;     // position
  double a;        // |E_te|
  cplx w;          // |E_tm| e^{i delta}
  double inv_cos;  // 1 / cos(theta_current.real)
  double ener;
  uint32_t rng;
  int state;
  int iter;
  int idx;         // ray index relative to the start of the run
};

// Rays that survived in-coupling wait here until a lane of the owning warp is free.  One queue per
// warp (a stack: the warp-uniform fill count lives in a register), filled and drained with
// ballot / popc prefix ranks, so no atomics and no block-wide barrier are involved.
constexpr int QUEUE_CAP = 64;
struct WarpQueue {
  double x[QUEUE_CAP], y[QUEUE_CAP], a[QUEUE_CAP], wre[QUEUE_CAP], wim[QUEUE_CAP], inv_cos[QUEUE_CAP],
      ener[QUEUE_CAP];
  uint32_t rng[QUEUE_CAP];
  int idx[QUEUE_CAP];
  int state[QUEUE_CAP];
};

__device__ __forceinline__ void jones_apply(const double* __restrict__ e, double a, cplx w, cplx& ote, cplx& otm) {
  // Ete_out = L0 * te + L2 * tm ; Etm_out = L1 * te + L3 * tm        (GRTF:139-144)
  const cplx L0{e[0], e[1]}, L1{e[2], e[3]}, L2{e[4], e[5]}, L3{e[6], e[7]};
  ote = cplx{L0.re * a + (L2.re * w.re - L2.im * w.im), L0.im * a + (L2.re * w.im + L2.im * w.re)};
  otm = cplx{L1.re * a + (L3.re * w.re - L3.im * w.im), L1.im * a + (L3.re * w.im + L3.im * w.re)};
}

__device__ __forceinline__ double power_of(cplx ote, cplx otm) {
  return ote.re * ote.re + ote.im * ote.im + (otm.re * otm.re + otm.im * otm.im);
}

// The ray takes the order whose table row is `row` and whose output amplitudes are (kte, ktm):
// normalise, add the TIR phase, advance by the new bounce vector, switch region state.
// (GRTF:872-882 and its eleven siblings; E_field_cal's |.| / atan2 / wrap in Jones-vector form.)
__device__ __forceinline__ void take_order(const CellConst& cc, const double* __restrict__ row, long long meta,
                                           cplx kte, cplx ktm, double esel, Ray& r) {
  const double te2 = kte.re * kte.re + kte.im * kte.im;
  const double tm2 = ktm.re * ktm.re + ktm.im * ktm.im;
  const double inv_norm = rsqrt(te2 + tm2);
  const cplx ph = cc.ph1[meta & 3];
  cplx num;
  const double eps2 = 1e-40;  // (1e-20)^2: E_field_cal zeroes a phase when its amplitude < 1e-20
  if (te2 >= eps2 && tm2 >= eps2) {
    const double inv_te = rsqrt(te2);
    r.a = te2 * inv_te * inv_norm;
    // E_tm * conj(E_te) / |E_te|: amplitude |E_tm|, phase phi_tm - phi_te
    num = cplx{(ktm.re * kte.re + ktm.im * kte.im) * inv_te, (ktm.im * kte.re - ktm.re * kte.im) * inv_te};
  } else {
    const double te_abs = sqrt(te2), tm_abs = sqrt(tm2);
    r.a = te_abs * inv_norm;
    if (te2 < eps2 && tm2 >= eps2) num = ktm;                                                // phi_te := 0
    else if (te2 >= eps2) num = cplx{tm_abs * kte.re / te_abs, -tm_abs * kte.im / te_abs};  // phi_tm := 0
    else num = cplx{tm_abs, 0.0};
  }
  num.re *= inv_norm;
  num.im *= inv_norm;
  r.w = cmul(num, ph);
  const int g = static_cast<int>((meta >> 2) & 3);
  r.x += cc.gap[2 * g];
  r.y += cc.gap[2 * g + 1];
  r.inv_cos = row[9];
  r.ener *= esel;
  r.state = static_cast<int>((meta >> 4) & 7);
}

// In-coupling (GRTF:842-904) for 32 consecutive rays of the run: one ray per lane, fully coalesced
// loads, every lane busy.  Rays that enter the waveguide are pushed on the warp's queue; the others
// (about three quarters with realistic gratings) are finished here.  Returns the new queue fill.
template <bool COUNT, bool IMPLICIT>
__device__ __forceinline__ int incouple_batch(const wgrt_problem_t& p, WalkShared& sh, const double* __restrict__ tab,
                                              int64_t run_begin, int first, int run_len, WarpQueue& q, int qn,
                                              int lane, unsigned lt_mask, Counts* cn) {
  const CellConst& cc = sh.cc;
  const int i = first + lane;
  bool survived = false;
  Ray r;
  if (i < run_len) {
    const int64_t idx = run_begin + i;
    r.idx = i;
    double tm;
    float dlf;
    if (IMPLICIT) {
      // runner layout (RUN:82-115): P TE rays then P TM rays per cell, ray k starts at point k
      const int64_t k = idx % (2 * p.runner_points);
      const bool te_half = k < p.runner_points;
      const int64_t pt = te_half ? k : k - p.runner_points;
      r.x = static_cast<double>(__ldg(p.x + pt));
      r.y = static_cast<double>(__ldg(p.y + pt));
      r.a = te_half ? 1.0 : 0.0;
      tm = te_half ? 0.0 : 1.0;
      dlf = 0.0f;
    } else {
      r.x = static_cast<double>(ld_stream(p.x + idx));
      r.y = static_cast<double>(ld_stream(p.y + idx));
      r.a = static_cast<double>(ld_stream(p.te + idx));
      tm = static_cast<double>(ld_stream(p.tm + idx));
      dlf = ld_stream(p.delta_phase + idx);
    }
    r.rng = ld_stream(p.rng_states + idx);
    // the batch this warp is likely to claim next (the CTA's other warps take the ones in between)
    if (i + WALK_THREADS < run_len) {
      const int64_t nxt = idx + WALK_THREADS;
      prefetch_l2(p.rng_states + nxt);
      if (!IMPLICIT) {
        prefetch_l2(p.x + nxt); prefetch_l2(p.y + nxt); prefetch_l2(p.te + nxt); prefetch_l2(p.tm + nxt);
        prefetch_l2(p.delta_phase + nxt);
      }
    }
    if (dlf == 0.0f) {
      r.w = cplx{tm, 0.0};
    } else {
      double sn, cs;
      sincos(static_cast<double>(dlf), &sn, &cs);
      r.w = cplx{tm * cs, tm * sn};
    }
    r.ener = 1.0;
    r.iter = 0;
    if (COUNT) {
      cn->c[WGRT_CNT_RAYS]++;
      cn->c[WGRT_CNT_DRAWS]++;
      cn->c[WGRT_CNT_DRAW2]++;
      cn->c[WGRT_CNT_EFIELD] += 2;
    }
    const double u = xorshift_draw(r.rng, p.ray_index_base + idx);
    cplx ote, otm;
    const double* row = tab;  // rows 0 and 1: the two in-coupled orders
    jones_apply(row, r.a, r.w, ote, otm);
    const double e1 = power_of(ote, otm) * row[8] * cc.inv_cos_in;
    double esel = e1;
    bool taken = u <= e1;  // GRTF:871: no energy gate on the in-coupling branches
    if (!taken) {
      row += ENTRY_DOUBLES;
      jones_apply(row, r.a, r.w, ote, otm);
      esel = power_of(ote, otm) * row[8] * cc.inv_cos_in;
      taken = u <= e1 + esel;  // GRTF:887
    }
    if (taken) {
      const long long meta = __double_as_longlong(row[10]);
      take_order(cc, row, meta, ote, otm, esel, r);
      if (COUNT) cn->c[WGRT_CNT_BOUNCES]++;
      const bool in_ic = region_locate<COUNT>(sh.reg[REG_IC], r.x, r.y, cn) >= 0;
      if (((meta >> 7) & 3) == POST_IC_FWD) {
        r.state = in_ic ? 0 : 2;  // GRTF:883-886
        survived = true;
      } else {
        survived = in_ic;         // GRTF:899-902
      }
    }
    if (!survived) st_stream(p.rng_states + idx, r.rng);
  }
  __syncwarp();
  const unsigned surv = __ballot_sync(FULL_MASK, survived);
  if (survived) {
    const int slot = qn + __popc(surv & lt_mask);
    q.x[slot] = r.x; q.y[slot] = r.y; q.a[slot] = r.a; q.wre[slot] = r.w.re; q.wim[slot] = r.w.im;
    q.inv_cos[slot] = r.inv_cos; q.ener[slot] = r.ener; q.rng[slot] = r.rng; q.idx[slot] = r.idx;
    q.state[slot] = r.state;
  }
  __syncwarp();
  return qn + __popc(surv);
}

// One loop iteration (GRTF:905-1246) for every ray a warp holds.  ALL 32 lanes call this (lanes
// without a ray idle through it): the step is cut into phases separated by __syncwarp(), so that
// after each divergent piece -- a region query that fell back to the exact edge scan, the
// free-bounce branch, three-order events -- the warp is whole again before the next piece.
// Everything state dependent comes out of shared-memory tables (sinfo, the event rows and their
// meta words): lanes run the same instructions whatever region state their rays are in.
template <bool COUNT>
__device__ __forceinline__ void walk_step(const wgrt_problem_t& p, WalkShared& sh, const double* __restrict__ tab,
                                          int64_t run_begin, int64_t lm, int64_t m, int64_t n, Ray& r, Counts* cn) {
  const CellConst& cc = sh.cc;
  const bool live = r.state != ST_DEAD;
  const int sinfo = live ? cc.sinfo[r.state] : 0;
  bool lost = false;  // the ray ends in this step

  // ---- phase 1: every region query whose point is already known, memory accesses overlapped:
  //      the loop's effective-region test (GRTF:905-907), the coupler slice under the ray, and for
  //      state 3 the eff_reg2 test its miss branch needs (GRTF:1103).  Queries are pure functions
  //      of (x, y), so asking one whose answer ends up unused changes nothing.
  const int region = sinfo & SI_REGION_MASK;
  const int miss = (sinfo >> SI_MISS_SHIFT) & 3;
  int hit = 0;
  bool in_r2 = true;
  if (live) {
    if (COUNT) cn->c[WGRT_CNT_ITERS]++;
    const Region* const regs[3] = {&sh.reg[REG_R1], &sh.reg[region == SI_REGION_NONE ? REG_R1 : region],
                                   &sh.reg[REG_R2]};
    const bool want[3] = {true, region != SI_REGION_NONE, miss == 1};
    int res[3];
    region_locate_multi<COUNT, 3>(regs, want, r.x, r.y, res, cn);
    if (++r.iter > 100000 || res[0] < 0) lost = true;
    if (region != SI_REGION_NONE) hit = res[1];
    in_r2 = res[2] >= 0;
  }
  __syncwarp();

  // ---- phase 3: no grating: free TIR bounce (GRTF:1049-1052, 1102-1108, 1175-1178, 1244-1246) ---
  int query = -1;  // region set to consult in phase 5
  const bool event = live && !lost && hit >= 0;
  if (live && !lost && hit < 0) {
    if (miss == 2) {
      lost = true;
    } else if (miss == 1 && !in_r2) {
      r.state = 4;  // GRTF:1103-1104: state 3 left the fold zone; no move in this iteration
    } else {
      const int g = (sinfo >> SI_GAP_SHIFT) & 3;
      r.x += cc.gap[2 * g];
      r.y += cc.gap[2 * g + 1];
      r.w = cmul(r.w, cc.ph2[(sinfo >> SI_PHASE_SHIFT) & 1]);
      if (COUNT) cn->c[WGRT_CNT_BOUNCES]++;
    }
  }
  __syncwarp();

  // ---- phase 4: grating event: draw once, then try the orders in the reference's order ----------
  int post = POST_NONE;
  if (event) {
    const double* e = tab + (((sinfo >> SI_ROWBASE_SHIFT) & SI_ROWBASE_MASK) + ((sinfo >> SI_STRIDE_SHIFT) & 3) * hit) * ENTRY_DOUBLES;
    const long long meta0 = __double_as_longlong(e[10]);
    const bool three = (meta0 & META_THREE) != 0;
    const bool gated = (meta0 & META_GATED) != 0;  // `and ener_k > threshold` (GRTF:1020 ff.), absent in GRTF:871-999
    const double u = xorshift_draw(r.rng, p.ray_index_base + run_begin + r.idx);
    if (COUNT) {
      cn->c[WGRT_CNT_DRAWS]++;
      cn->c[three ? WGRT_CNT_DRAW3 : WGRT_CNT_DRAW2]++;
      cn->c[WGRT_CNT_EFIELD] += three ? 3 : 2;
    }
    cplx ote, otm;
    const double* row = e;
    jones_apply(row, r.a, r.w, ote, otm);
    double esel = power_of(ote, otm) * row[8] * r.inv_cos;
    double esum = esel;
    const double threshold = p.threshold;
    bool taken = u <= esum && (!gated || r.ener * esel > threshold);
    if (!taken) {
      row = e + ENTRY_DOUBLES;
      jones_apply(row, r.a, r.w, ote, otm);
      esel = power_of(ote, otm) * row[8] * r.inv_cos;
      esum += esel;
      taken = u <= esum && (!gated || r.ener * esel > threshold);
      if (!taken && three) {
        row = e + 2 * ENTRY_DOUBLES;
        jones_apply(row, r.a, r.w, ote, otm);
        esel = power_of(ote, otm) * row[8] * r.inv_cos;
        taken = u <= esum + esel && r.ener * esel > threshold;
      }
    }
    if (!taken) {
      lost = true;  // absorbed: u above every cumulative efficiency
    } else {
      const long long meta = __double_as_longlong(row[10]);
      post = static_cast<int>((meta >> 7) & 3);
      if (post == POST_DEPOSIT) {
        // GRTF:1162-1171: count the ray if it leaves inside this FoV's eyebox rectangle
        if (inside_or_on_edge_literal<COUNT>(r.x, r.y, cc.rect, 0, 4, cn)) {
          deposit_bin(p, lm, m, n, r.x, r.y, cc.range[0], cc.range[1], cc.range[2], cc.range[3]);
          if (COUNT) cn->c[WGRT_CNT_DEPOSITS]++;
        }
        lost = true;
      } else {
        take_order(cc, row, meta, ote, otm, esel, r);
        if (COUNT) cn->c[WGRT_CNT_BOUNCES]++;
        if (post != POST_NONE) query = REG_IC;
      }
    }
  }
  __syncwarp();

  // ---- phase 5: did the in-coupler event leave the ray inside the in-coupler? ---------------------
  if (query >= 0) {
    const bool in = region_locate<COUNT>(sh.reg[REG_IC], r.x, r.y, cn) >= 0;
    if (post == POST_IC_FWD) r.state = in ? 0 : 2;  // GRTF:932-935
    else if (!in) lost = true;                       // GRTF:948-951
  }
  __syncwarp();
  if (lost) {
    st_stream(p.rng_states + run_begin + r.idx, r.rng);
    r.state = ST_DEAD;
  }
}

template <bool COUNT, bool IMPLICIT>
__global__ void __launch_bounds__(WALK_THREADS, WGRT_WALK_MIN_BLOCKS)
walk_fast_kernel(const __grid_constant__ wgrt_problem_t p, const __grid_constant__ RegionSet rs,
                 int* __restrict__ work_counter, const int* __restrict__ tile_size_ptr,
                 unsigned long long* counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WalkShared& sh = *reinterpret_cast<WalkShared*>(smem_raw);
  const int rows = 6 + 4 * static_cast<int>(p.n_FC) + 6 * static_cast<int>(p.n_OC);
  double* tab = reinterpret_cast<double*>(smem_raw + walk_smem_table_offset());
  WarpQueue& queue = reinterpret_cast<WarpQueue*>(smem_raw + walk_smem_queue_offset(rows))[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  Counts cn;
  if (COUNT) cn.clear();

  if (threadIdx.x < NUM_REGIONS) region_load(sh.reg[threadIdx.x], rs.st[threadIdx.x], rs.dyn[threadIdx.x]);
  const int64_t tile_size = *tile_size_ptr;
  const int64_t num_tiles = (p.num_rays + tile_size - 1) / tile_size;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) sh.tile = atomicAdd(work_counter, 1);
    __syncthreads();
    const int64_t tile = sh.tile;
    if (tile >= num_tiles) break;
    const int64_t t_begin = tile * tile_size;
    const int64_t t_end = min(p.num_rays, t_begin + tile_size);

    int64_t run_begin = t_begin;
    while (run_begin < t_end) {
      // ---- find the run of rays sharing the cell of ray run_begin --------------------------
      float km, kn, kl;
      int64_t run_end;
      if (IMPLICIT) {
        const int64_t rpc = 2 * p.runner_points;
        const int64_t cell = p.runner_first_cell + run_begin / rpc;  // runner order: x outer, y, lambda inner
        kl = static_cast<float>(cell % p.L);
        kn = static_cast<float>((cell / p.L) % p.Y);
        km = static_cast<float>(cell / (p.L * p.Y));
        run_end = min(t_end, (run_begin / rpc + 1) * rpc);
        __syncthreads();  // previous run fully walked; shared tables may be overwritten
      } else {
        const bool has_l = p.lmd_num != nullptr;
        km = __ldg(p.m + run_begin); kn = __ldg(p.n + run_begin); kl = has_l ? __ldg(p.lmd_num + run_begin) : 0.0f;
        __syncthreads();  // previous run fully walked; shared tables may be overwritten
        if (threadIdx.x == 0) sh.run_end = static_cast<int>(t_end - t_begin);
        __syncthreads();
        // one pass over the rest of the tile with every load in flight at once (no early exit, no
        // barrier per chunk: that serialised DRAM latency and was 10 % of the kernel)
        int first_bad = INT_MAX;
#pragma unroll 4
        for (int64_t i = run_begin + 1 + threadIdx.x; i < t_end; i += WALK_THREADS) {
          const bool bad = ld_stream(p.m + i) != km || ld_stream(p.n + i) != kn || (has_l && ld_stream(p.lmd_num + i) != kl);
          first_bad = min(first_bad, bad ? static_cast<int>(i - t_begin) : INT_MAX);
        }
        if (first_bad != INT_MAX) atomicMin(&sh.run_end, first_bad);
        __syncthreads();
        run_end = t_begin + sh.run_end;
      }
      const int64_t m = static_cast<int64_t>(km), n = static_cast<int64_t>(kn), lm = static_cast<int64_t>(kl);
      const bool valid = m >= 0 && m < p.X && n >= 0 && n < p.Y && lm >= 0 && lm < p.L;
      if (valid) {
        build_cell_tables(p, lm, m, n, tab, sh.cc, rows);
        if (threadIdx.x == 0) sh.run_cursor = 0;
        __syncthreads();
        const int run_len = static_cast<int>(run_end - run_begin);

        // ---- each warp on its own: in-couple batches of 32 rays whenever more lanes are free
        //      than rays are queued, refill free lanes from the queue, step the rays it holds ----
        Ray r;
        r.state = ST_DEAD;
        int qn = 0;            // rays waiting in this warp's queue (warp uniform)
        bool run_open = true;  // the run still has rays nobody in-coupled
        for (;;) {
          const unsigned dead = __ballot_sync(FULL_MASK, r.state == ST_DEAD);
          const int nd = __popc(dead);
          while (run_open && qn < nd) {
            int first = 0;
            if (lane == 0) first = atomicAdd(&sh.run_cursor, 32);
            first = __shfl_sync(FULL_MASK, first, 0);
            if (first >= run_len) {
              run_open = false;
              break;
            }
            qn = incouple_batch<COUNT, IMPLICIT>(p, sh, tab, run_begin, first, run_len, queue, qn, lane, lt_mask, &cn);
            if (COUNT && lane == 0) cn.c[WGRT_CNT_WARP_BATCHES]++;
          }
          if (nd && qn) {
            const int take = min(nd, qn);
            const int rank = __popc(dead & lt_mask);
            if (r.state == ST_DEAD && rank < take) {
              const int slot = qn - 1 - rank;
              r.x = queue.x[slot]; r.y = queue.y[slot]; r.a = queue.a[slot];
              r.w = cplx{queue.wre[slot], queue.wim[slot]};
              r.inv_cos = queue.inv_cos[slot]; r.ener = queue.ener[slot]; r.rng = queue.rng[slot];
              r.idx = queue.idx[slot]; r.state = queue.state[slot]; r.iter = 0;
            }
            qn -= take;
            __syncwarp();
          }
          if (__ballot_sync(FULL_MASK, r.state != ST_DEAD) == 0u) {
            if (!run_open) break;  // queue is empty too: every queued ray found a free lane above
            continue;
          }
          if (COUNT && lane == 0) cn.c[WGRT_CNT_WARP_STEPS]++;
          walk_step<COUNT>(p, sh, tab, run_begin, lm, m, n, r, &cn);
        }
      }
      run_begin = run_end;
    }
  }
  if (COUNT) cn.flush(counters);
}

// ---------------------------------------------------------------------------------------------
// tile size: one work tile should hold whole runs.  A single warp measures the first run.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pick_tile_kernel(const __grid_constant__ wgrt_problem_t p, int* tile_size,
                                                         int* work_counter) {
  __shared__ int s_run;
  if (threadIdx.x == 0) {
    *work_counter = 0;
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
    // a tile should be a whole run (or whole runs): long runs are cut into equal pieces of at
    // most 5120 rays, short ones are grouped up to at least 2560 (~20-40 rays per lane keeps the
    // end-of-run tail small)
    const int64_t t_min = 2560, t_max = 5120;
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

// ---------------------------------------------------------------------------------------------
// region index construction (see wgrt_region.cuh)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int total_ring_verts(const RegionStatic& st) {
  return st.offsets ? static_cast<int>(st.offsets[st.npoly]) : st.nverts;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// The region index depends only on the vertex / offset CONTENT and the grid geometry.  Hash them on
// the device at every launch (a few KB) and rebuild the index only when the hash changed: the
// runner launches the kernel num_iter times on the same design (gpu_ray_tracing_pro_fullColor.py:
// 169-177).  state[0] = hash of the index currently built, state[1] = dirty flag for this launch.
__global__ void region_hash_kernel(const __grid_constant__ RegionSet rs, unsigned long long* state, int force) {
  __shared__ unsigned long long s_acc[256];
  unsigned long long acc = 0;
  for (int r = 0; r < NUM_REGIONS; ++r) {
    const RegionStatic& st = rs.st[r];
    const unsigned long long tag = mix64(0x1000ull * (r + 1));
    if (threadIdx.x == 0)
      acc += mix64(tag ^ mix64((static_cast<unsigned long long>(st.nverts) << 32) ^ st.npoly) ^
                   mix64((static_cast<unsigned long long>(st.n) << 32) ^ (st.nc << 8) ^ st.shift));
    const unsigned long long* v = reinterpret_cast<const unsigned long long*>(st.verts);
    for (int i = threadIdx.x; i < 2 * st.nverts; i += blockDim.x) acc += mix64(v[i] ^ mix64(tag + i));
    if (st.offsets)
      for (int i = threadIdx.x; i <= st.npoly; i += blockDim.x)
        acc += mix64(static_cast<unsigned long long>(st.offsets[i]) ^ mix64(tag + 0x80000000ull + i));
  }
  s_acc[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_acc[threadIdx.x] += s_acc[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const unsigned long long h = s_acc[0] | 1ull;  // never 0, the "nothing built" value
    state[1] = (force || state[0] != h) ? 1ull : 0ull;
    state[0] = h;
  }
}

__global__ void region_bbox_kernel(const __grid_constant__ RegionSet rs) {
  if (!*rs.dirty) return;
  const RegionStatic& st = rs.st[blockIdx.x];
  __shared__ double s_min[2][256], s_max[2][256];
  const int nv = min(total_ring_verts(st), st.nverts);
  double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const double vx = st.verts[2 * i], vy = st.verts[2 * i + 1];
    xmin = fmin(xmin, vx); xmax = fmax(xmax, vx);
    ymin = fmin(ymin, vy); ymax = fmax(ymax, vy);
  }
  s_min[0][threadIdx.x] = xmin; s_max[0][threadIdx.x] = xmax;
  s_min[1][threadIdx.x] = ymin; s_max[1][threadIdx.x] = ymax;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_min[0][threadIdx.x] = fmin(s_min[0][threadIdx.x], s_min[0][threadIdx.x + o]);
      s_max[0][threadIdx.x] = fmax(s_max[0][threadIdx.x], s_max[0][threadIdx.x + o]);
      s_min[1][threadIdx.x] = fmin(s_min[1][threadIdx.x], s_min[1][threadIdx.x + o]);
      s_max[1][threadIdx.x] = fmax(s_max[1][threadIdx.x], s_max[1][threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    RegionDyn d;
    xmin = s_min[0][0]; xmax = s_max[0][0]; ymin = s_min[1][0]; ymax = s_max[1][0];
    if (!(xmax >= xmin) || !(ymax >= ymin) || !isfinite(xmax - xmin) || !isfinite(ymax - ymin)) {
      xmin = ymin = 0.0; xmax = ymax = 1.0;  // empty or non-finite ring set: every cell ends up NONE/AMBIG
    }
    const double pad_x = 1e-3 * (xmax - xmin) + 1e-9, pad_y = 1e-3 * (ymax - ymin) + 1e-9;
    d.x0 = xmin - pad_x;
    d.y0 = ymin - pad_y;
    d.cell_dx = (xmax - xmin + 2.0 * pad_x) / st.n;
    d.cell_dy = (ymax - ymin + 2.0 * pad_y) / st.n;
    d.inv_dx = 1.0 / d.cell_dx;
    d.inv_dy = 1.0 / d.cell_dy;
    rs.dyn[blockIdx.x] = d;
  }
}

__device__ __forceinline__ double margin_of(double cell) { return 0.02 * cell + 1e-11; }

// Row masks: bit i of row r is set when the edge (prev(i) -> i) can matter to a point whose y lies in
// row r (plus margin).  blockIdx.z = 0 builds the fine rows, 1 the coarse rows.
__global__ void region_rowmask_kernel(const __grid_constant__ RegionSet rs) {
  if (!*rs.dirty) return;
  const RegionStatic& st = rs.st[blockIdx.y];
  const RegionDyn d = rs.dyn[blockIdx.y];
  const bool coarse = blockIdx.z == 1;
  const int nrows = coarse ? st.nc : st.n;
  const double row_h = coarse ? d.cell_dy * (1 << st.shift) : d.cell_dy;
  uint32_t* out = coarse ? st.rowmask_coarse : st.rowmask;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nrows * st.words) return;
  const int row = t / st.words, word = t - row * st.words;
  const double mrg = margin_of(row_h);
  const double row_lo = d.y0 + row * row_h - mrg, row_hi = d.y0 + (row + 1) * row_h + mrg;
  const int nv = min(total_ring_verts(st), st.nverts);
  uint32_t bits = 0;
  int k = 0;
  for (int b = 0; b < 32; ++b) {
    const int i = word * 32 + b;
    if (i >= nv) break;
    while (k < st.npoly && ring_begin(st.offsets, st.nverts, k + 1) <= i) ++k;
    if (k >= st.npoly) break;
    const int s = ring_begin(st.offsets, st.nverts, k), e = ring_begin(st.offsets, st.nverts, k + 1);
    const int j = (i == s) ? e - 1 : i - 1;
    const double yi = st.verts[2 * i + 1], yj = st.verts[2 * j + 1];
    if (!(fmax(yi, yj) < row_lo || fmin(yi, yj) > row_hi)) bits |= 1u << b;
  }
  out[t] = bits;
}

// Classify one grid cell [x_lo, x_hi] x [y_lo, y_hi] (already inflated by the safety margin) with
// centre (cx, cy), looking only at the edges in `mask`.
__device__ __forceinline__ uint8_t classify_cell(const RegionStatic& st, const uint32_t* __restrict__ mask, double x_lo,
                                                 double x_hi, double y_lo, double y_hi, double cx, double cy,
                                                 uint32_t& detail) {
  int first_unc = -1, last_unc = -1, hit_ring = -1;
  for (int k = 0; k < st.npoly && hit_ring < 0; ++k) {
    const int s = ring_begin(st.offsets, st.nverts, k), e = ring_begin(st.offsets, st.nverts, k + 1);
    if (e <= s) continue;
    bool near_edge = false, inside = false;
    for (int w = s >> 5; w <= (e - 1) >> 5 && !near_edge; ++w) {
      uint32_t bits = mask ? mask[w] : 0xffffffffu;   // no mask: every edge of the ring
      if (w == (s >> 5)) bits &= 0xffffffffu << (s & 31);
      if (w == ((e - 1) >> 5) && (e & 31)) bits &= 0xffffffffu >> (32 - (e & 31));
      while (bits) {
        const int i = (w << 5) + __ffs(bits) - 1;
        bits &= bits - 1;
        const int j = (i == s) ? e - 1 : i - 1;
        const double xi = st.verts[2 * i], yi = st.verts[2 * i + 1];
        const double xj = st.verts[2 * j], yj = st.verts[2 * j + 1];
        // conservative segment / inflated-cell intersection: bounding boxes meet and the four
        // corners are not strictly on one side of the supporting line (NaNs fall through to "near")
        if (!(fmax(xi, xj) < x_lo || fmin(xi, xj) > x_hi || fmax(yi, yj) < y_lo || fmin(yi, yj) > y_hi)) {
          const double ex = xi - xj, ey = yi - yj;
          const double c0 = ex * (y_lo - yj) - ey * (x_lo - xj);
          const double c1 = ex * (y_lo - yj) - ey * (x_hi - xj);
          const double c2 = ex * (y_hi - yj) - ey * (x_lo - xj);
          const double c3 = ex * (y_hi - yj) - ey * (x_hi - xj);
          const bool all_pos = c0 > 0 && c1 > 0 && c2 > 0 && c3 > 0;
          const bool all_neg = c0 < 0 && c1 < 0 && c2 < 0 && c3 < 0;
          if (!(all_pos || all_neg)) { near_edge = true; break; }
        }
        if ((yi > cy) != (yj > cy))
          if (cx < (xj - xi) * (cy - yi) / (yj - yi + 1e-20) + xi) inside = !inside;
      }
    }
    if (near_edge) {
      if (first_unc < 0) first_unc = k;
      last_unc = k;
    } else if (inside) {
      hit_ring = k;   // certainly inside ring k: later rings can never be the first hit
    }
  }
  detail = 0;
  if (first_unc >= 0) {
    detail = hit_ring >= 0 ? pack_detail(first_unc, hit_ring, hit_ring) : pack_detail(first_unc, last_unc + 1, 255);
    return CELL_AMBIG;
  }
  return hit_ring >= 0 ? static_cast<uint8_t>(hit_ring) : CELL_NONE;
}

__global__ void region_coarse_kernel(const __grid_constant__ RegionSet rs) {
  if (!*rs.dirty) return;
  const RegionStatic& st = rs.st[blockIdx.y];
  const RegionDyn d = rs.dyn[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= st.nc * st.nc) return;
  const int iy = t / st.nc, ix = t - iy * st.nc;
  const double w = d.cell_dx * (1 << st.shift), h = d.cell_dy * (1 << st.shift);
  const double mx = margin_of(w), my = margin_of(h);
  uint32_t detail;
  st.coarse[t] = classify_cell(st, st.rowmask_coarse + static_cast<size_t>(iy) * st.words, d.x0 + ix * w - mx,
                               d.x0 + (ix + 1) * w + mx, d.y0 + iy * h - my, d.y0 + (iy + 1) * h + my,
                               d.x0 + (ix + 0.5) * w, d.y0 + (iy + 0.5) * h, detail);
}

// One block per coarse cell; only MIXED coarse cells get their (1 << shift)^2 fine cells classified.
__global__ void region_fine_kernel(const __grid_constant__ RegionSet rs) {
  if (!*rs.dirty) return;
  const RegionStatic& st = rs.st[blockIdx.y];
  if (static_cast<int>(blockIdx.x) >= st.nc * st.nc) return;
  if (st.coarse[blockIdx.x] != CELL_AMBIG) return;
  const RegionDyn d = rs.dyn[blockIdx.y];
  const int cy = blockIdx.x / st.nc, cx = blockIdx.x - cy * st.nc;
  const int S = 1 << st.shift;
  const double mx = margin_of(d.cell_dx), my = margin_of(d.cell_dy);
  for (int t = threadIdx.x; t < S * S; t += blockDim.x) {
    const int iy = (cy << st.shift) + t / S, ix = (cx << st.shift) + (t & (S - 1));
    uint32_t detail;
    const uint8_t code = classify_cell(st, st.rowmask + static_cast<size_t>(iy) * st.words, d.x0 + ix * d.cell_dx - mx,
                                       d.x0 + (ix + 1) * d.cell_dx + mx, d.y0 + iy * d.cell_dy - my,
                                       d.y0 + (iy + 1) * d.cell_dy + my, d.x0 + (ix + 0.5) * d.cell_dx,
                                       d.y0 + (iy + 0.5) * d.cell_dy, detail);
    const size_t cell = static_cast<size_t>(iy) * st.n + ix;
    st.cells[cell] = code;
    if (code == CELL_AMBIG) st.detail[cell] = detail;
  }
}

// The atlas (wgrt_region.cuh).  Level 1: one thread per cell classifies the cell against all five
// region sets, looking at every edge; thread 0 publishes the atlas geometry.  Level 2: one block per
// level-1 cell with a MIXED field re-classifies its 64 x 64 sub-cells for the MIXED sets only.
__device__ __forceinline__ void atlas_bbox(const RegionSet& rs, double& xmin, double& ymin, double& w, double& h) {
  double xmax = -INFINITY, ymax = -INFINITY;
  xmin = INFINITY; ymin = INFINITY;
  for (int r = 0; r < NUM_REGIONS; ++r) {
    if (min(total_ring_verts(rs.st[r]), rs.st[r].nverts) <= 0) continue;   // empty set: always "outside"
    const RegionDyn d = rs.dyn[r];
    xmin = fmin(xmin, d.x0); xmax = fmax(xmax, d.x0 + d.cell_dx * rs.st[r].n);
    ymin = fmin(ymin, d.y0); ymax = fmax(ymax, d.y0 + d.cell_dy * rs.st[r].n);
  }
  if (!(xmax > xmin) || !(ymax > ymin) || !isfinite(xmax - xmin) || !isfinite(ymax - ymin)) {
    xmin = ymin = 0.0; xmax = ymax = 1.0;
  }
  w = (xmax - xmin) / ATLAS_N;
  h = (ymax - ymin) / ATLAS_N;
}

__device__ __forceinline__ uint32_t atlas_field(int r, uint8_t code) {
  if (r == REG_FC) return static_cast<uint32_t>(code) << ATLAS_SHIFT_FC;
  if (r == REG_OC) return static_cast<uint32_t>(code) << ATLAS_SHIFT_OC;
  const uint32_t c = code == CELL_AMBIG ? 2u : (code == CELL_NONE ? 0u : 1u);
  return c << (r == REG_IC ? ATLAS_SHIFT_IC : r == REG_R1 ? ATLAS_SHIFT_R1 : ATLAS_SHIFT_R2);
}
__device__ __forceinline__ uint32_t atlas_field_mask(int r) {
  return r == REG_FC ? 0xffu << ATLAS_SHIFT_FC : r == REG_OC ? 0xffu << ATLAS_SHIFT_OC
       : 3u << (r == REG_IC ? ATLAS_SHIFT_IC : r == REG_R1 ? ATLAS_SHIFT_R1 : ATLAS_SHIFT_R2);
}
__device__ __forceinline__ bool atlas_field_mixed(int r, uint32_t word) {
  const uint32_t f = word & atlas_field_mask(r);
  return f == atlas_field(r, CELL_AMBIG);
}

__global__ void region_atlas_kernel(const __grid_constant__ RegionSet rs) {
  // the region descriptors the walk's rare exact path reads: refreshed at EVERY build (the vertex
  // pointers may have changed even when the content hash, and with it the index, did not)
  if (blockIdx.x == 0 && threadIdx.x < NUM_REGIONS)
    region_load(static_cast<Region*>(rs.regions)[threadIdx.x], rs.st[threadIdx.x], rs.dyn[threadIdx.x]);
  if (!*rs.dirty) return;
  double xmin, ymin, w, h;
  atlas_bbox(rs, xmin, ymin, w, h);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) *rs.atlas_dyn = AtlasDyn{xmin, ymin, 1.0 / w, 1.0 / h};
  if (t >= ATLAS_N * ATLAS_N) return;
  const int iy = t / ATLAS_N, ix = t - iy * ATLAS_N;
  const double mx = margin_of(w), my = margin_of(h);
  uint32_t word = 0;
  for (int r = 0; r < NUM_REGIONS; ++r) {
    uint32_t detail;
    const uint8_t code = classify_cell(rs.st[r], nullptr, xmin + ix * w - mx, xmin + (ix + 1) * w + mx, ymin + iy * h - my,
                                       ymin + (iy + 1) * h + my, xmin + (ix + 0.5) * w, ymin + (iy + 0.5) * h, detail);
    word |= atlas_field(r, code);
    if (code == CELL_AMBIG) word |= ATLAS_ANY_MIXED;
  }
  rs.atlas[t] = word;
}

__global__ void __launch_bounds__(256) region_atlas2_kernel(const __grid_constant__ RegionSet rs) {
  if (!*rs.dirty) return;
  const uint32_t coarse = rs.atlas[blockIdx.x];
  if (!(coarse & ATLAS_ANY_MIXED)) return;
  double xmin, ymin, w, h;
  atlas_bbox(rs, xmin, ymin, w, h);
  constexpr int S = 1 << ATLAS_SUB_SHIFT;
  const double w2 = w / S, h2 = h / S;   // level-2 index of a point: int(fx * S) with fx in level-1 cells
  const int cy = blockIdx.x / ATLAS_N, cx = blockIdx.x - cy * ATLAS_N;
  uint32_t* out = rs.atlas + ATLAS_N * ATLAS_N;
  const double mx = margin_of(w2), my = margin_of(h2);
  for (int t = threadIdx.x; t < S * S; t += blockDim.x) {
    const int iy = (cy << ATLAS_SUB_SHIFT) + t / S, ix = (cx << ATLAS_SUB_SHIFT) + (t & (S - 1));
    uint32_t word = coarse & ~ATLAS_ANY_MIXED;
    for (int r = 0; r < NUM_REGIONS; ++r) {
      if (!atlas_field_mixed(r, coarse)) continue;   // certain for the whole level-1 cell
      uint32_t detail;
      const uint8_t code = classify_cell(rs.st[r], nullptr, xmin + ix * w2 - mx, xmin + (ix + 1) * w2 + mx,
                                         ymin + iy * h2 - my, ymin + (iy + 1) * h2 + my, xmin + (ix + 0.5) * w2,
                                         ymin + (iy + 0.5) * h2, detail);
      word = (word & ~atlas_field_mask(r)) | atlas_field(r, code);
      if (code == CELL_AMBIG) word |= ATLAS_ANY_MIXED;
    }
    out[static_cast<size_t>(iy) * ATLAS_N2 + ix] = word;
  }
}

// gpu_ray_tracing_pro_fullColor.py:158: rng_states[i] = 0x9E3779B9 * (i + 1) mod 2^32
__global__ void seed_rng_kernel(uint32_t* __restrict__ states, int64_t n, int64_t first_index) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) states[i] = 0x9E3779B9u * static_cast<uint32_t>(first_index + i + 1);
}

template <bool COUNT>
__global__ void locate_grid_kernel(const __grid_constant__ RegionSet rs, int region, const double* px,
                                   const double* py, int64_t n, int32_t* out, unsigned long long* counters,
                                   int via_atlas) {
  __shared__ Region reg;
  __shared__ Atlas atlas;
  if (threadIdx.x == 0) {
    region_load(reg, rs.st[region], rs.dyn[region]);
    const AtlasDyn ad = *rs.atlas_dyn;
    atlas.x0 = ad.x0; atlas.y0 = ad.y0; atlas.inv_dx = ad.inv_dx; atlas.inv_dy = ad.inv_dy; atlas.words = rs.atlas;
    atlas.words2 = rs.atlas + ATLAS_N * ATLAS_N;
  }
  __syncthreads();
  Counts cn;
  if (COUNT) cn.clear();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    if (via_atlas) {
      const uint32_t word = atlas_lookup(atlas, px[i], py[i]);
      out[i] = (region == REG_FC || region == REG_OC)
                   ? atlas_hit<COUNT>(word, region == REG_FC ? ATLAS_SHIFT_FC : ATLAS_SHIFT_OC, reg, px[i], py[i], &cn)
                   : (atlas_inside<COUNT>(word, region == REG_IC ? ATLAS_SHIFT_IC : region == REG_R1 ? ATLAS_SHIFT_R1 : ATLAS_SHIFT_R2,
                                          reg, px[i], py[i], &cn) ? 0 : -1);
    } else {
      out[i] = region_locate<COUNT>(reg, px[i], py[i], &cn);
    }
  }
  if (COUNT) cn.flush(counters);
}

// pupil-mask sums (AR_system_evaluation_functions.py:68-109) and per-cell totals
// (gpu_ray_tracing_pro_fullColor.py:186): one CTA per (lambda, FoV-y, FoV-x) bin tile.
__global__ void __launch_bounds__(256) pupil_sums_kernel(const float* __restrict__ EB, int64_t tiles, int EBy,
                                                         int EBx, int mask, int step_y, int step_x, int n_epy,
                                                         int n_epx, float* __restrict__ out,
                                                         float* __restrict__ cell_sums) {
  extern __shared__ float s_tile[];
  const int64_t tile = blockIdx.x;
  if (tile >= tiles) return;
  const int npix = EBy * EBx;
  const float* src = EB + tile * npix;
  float local = 0.f;
  for (int i = threadIdx.x; i < npix; i += blockDim.x) {
    const float v = __ldg(src + i);
    s_tile[i] = v;
    local += v;
  }
  // counts are non-negative integers < 2^24 per tile in practice; float sum of a tile stays exact
  __shared__ float s_red[8];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL_MASK, local, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0 && cell_sums) {
    float tot = 0.f;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) tot += s_red[w];
    cell_sums[tile] = tot;
  }
  if (!out) return;
  const float radius = mask * 0.5f, ctr = radius - 0.5f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int pos = warp; pos < n_epy * n_epx; pos += nwarps) {
    const int y0 = (pos / n_epx) * step_y, x0 = (pos % n_epx) * step_x;
    float acc = 0.f;
    for (int q = lane; q < mask * mask; q += 32) {
      const int my = q / mask, mx = q - my * mask;
      const float dx = mx - ctr, dy = my - ctr;
      if (sqrtf(dx * dx + dy * dy) <= radius) acc += s_tile[(y0 + my) * EBx + (x0 + mx)];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
    if (lane == 0) out[tile * (n_epy * n_epx) + pos] = acc;
  }
}

// Dense pupil sampling (down to the full pupil convolution the reference comments out as "super
// long", AR_system_evaluation_functions.py:75-89) and eyebox tiles too large for shared memory
// (BASELINE config 4: 320 x 480 bins): one CTA per (tile, block of output positions).  The CTA stages
// the window of bins its outputs need as ROW PREFIX SUMS; the disc mask is a contiguous column range
// [a_r, b_r] in each of its rows, so an output is sum_r (P[y0+r][x0+b_r+1] - P[y0+r][x0+a_r]):
// 2 * mask loads instead of ~0.785 * mask^2.  Bins are integer counts, float32 prefix sums of fewer than
// 2^24 counts are exact, so the result equals the direct sum bit for bit.
__global__ void __launch_bounds__(256) pupil_window_kernel(const float* __restrict__ EB, int EBy, int EBx, int mask,
                                                           int step_y, int step_x, int n_epy, int n_epx, int boy, int box,
                                                           int blocks_x, float* __restrict__ out) {
  extern __shared__ float s_pre[];              // [win_rows][win_cols + 1]
  __shared__ short s_a[256], s_b[256];          // column range of the disc in mask row r (mask <= 256)
  const int64_t tile = blockIdx.x;
  const int by = blockIdx.y / blocks_x, bx = blockIdx.y - by * blocks_x;
  const int oy0 = by * boy, ox0 = bx * box;
  const int ny = min(boy, n_epy - oy0), nx = min(box, n_epx - ox0);
  const int wy0 = oy0 * step_y, wx0 = ox0 * step_x;
  const int win_rows = (ny - 1) * step_y + mask, win_cols = (nx - 1) * step_x + mask;
  const int pitch = (box - 1) * step_x + mask + 1;
  const float radius = mask * 0.5f, ctr = radius - 0.5f;
  for (int r = threadIdx.x; r < mask; r += blockDim.x) {
    int a = mask, b = -1;
    const float dy = r - ctr;
    for (int c = 0; c < mask; ++c) {
      const float dx = c - ctr;
      if (sqrtf(dx * dx + dy * dy) <= radius) { a = min(a, c); b = c; }   // same rule as pupil_sums_kernel / EVAL:68-73
    }
    s_a[r] = static_cast<short>(a);
    s_b[r] = static_cast<short>(b);
  }
  const float* src = EB + tile * static_cast<int64_t>(EBy) * EBx;
  // one warp per window row: inclusive scan of the row into s_pre[row][1..], s_pre[row][0] = 0
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int row = warp; row < win_rows; row += nwarps) {
    const float* g = src + static_cast<int64_t>(wy0 + row) * EBx + wx0;
    float* pr = s_pre + row * pitch;
    if (lane == 0) pr[0] = 0.f;
    float carry = 0.f;
    for (int c0 = 0; c0 < win_cols; c0 += 32) {
      const int c = c0 + lane;
      float v = c < win_cols ? __ldg(g + c) : 0.f;
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(FULL_MASK, v, o);
        if (lane >= o) v += t;
      }
      v += carry;
      if (c < win_cols) pr[c + 1] = v;
      carry = __shfl_sync(FULL_MASK, v, 31);
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < ny * nx; q += blockDim.x) {
    const int iy = q / nx, ix = q - iy * nx;
    const float* base = s_pre + (iy * step_y) * pitch + ix * step_x;
    float acc = 0.f;
    for (int r = 0; r < mask; ++r) {
      const int a = s_a[r], b = s_b[r];
      if (b >= a) acc += base[r * pitch + b + 1] - base[r * pitch + a];
    }
    out[(tile * n_epy + (oy0 + iy)) * n_epx + (ox0 + ix)] = acc;
  }
}

// Bin tensor <-> uint8 for the exact narrow all-reduce (multi_gpu.reduce_bins): one pass that converts,
// and reports the largest entry and whether any entry is not an integer in [0, 255].
__global__ void __launch_bounds__(256) bins_pack_u8_kernel(const float4* __restrict__ in, int64_t n4, uint32_t* __restrict__ out,
                                                           unsigned* __restrict__ stats, float limit) {
  float vmax = 0.f;
  bool bad = false;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(in + i);
    const float f[4] = {v.x, v.y, v.z, v.w};
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float c = fminf(fmaxf(f[k], 0.f), 255.f);
      const uint32_t q = static_cast<uint32_t>(c);
      bad |= !(static_cast<float>(q) == f[k]) || c > limit;   // negative, > limit, fractional or NaN
      vmax = fmaxf(vmax, c);
      w |= q << (8 * k);
    }
    out[i] = w;
  }
  for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULL_MASK, vmax, o));
  const unsigned any_bad = __ballot_sync(FULL_MASK, bad);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(stats, __float_as_uint(vmax));          // non-negative floats order like their bit patterns
    if (any_bad) atomicOr(stats + 1, 1u);
  }
}

__global__ void __launch_bounds__(256) bins_unpack_u8_kernel(const uint32_t* __restrict__ in, int64_t n4, float4* __restrict__ out) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t w = __ldg(in + i);
    out[i] = make_float4(static_cast<float>(w & 255u), static_cast<float>((w >> 8) & 255u),
                         static_cast<float>((w >> 16) & 255u), static_cast<float>(w >> 24));
  }
}

// per-cell totals straight from global memory (tiles of any size)
__global__ void __launch_bounds__(256) cell_sums_kernel(const float* __restrict__ EB, int64_t npix, float* __restrict__ cell_sums) {
  const float* src = EB + static_cast<int64_t>(blockIdx.x) * npix;
  float local = 0.f;
  for (int64_t i = threadIdx.x; i < npix; i += blockDim.x) local += __ldg(src + i);
  __shared__ float s_red[8];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL_MASK, local, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) tot += s_red[w];
    cell_sums[blockIdx.x] = tot;
  }
}

}  // namespace

cudaError_t launch_region_build(const RegionSet& rs, bool force, cudaStream_t s) {
  region_hash_kernel<<<1, 256, 0, s>>>(rs, rs.hash_state, force ? 1 : 0);
  region_bbox_kernel<<<NUM_REGIONS, 256, 0, s>>>(rs);
  int max_rw = 1, max_coarse = 1;
  for (int r = 0; r < NUM_REGIONS; ++r) {
    max_rw = max(max_rw, rs.st[r].n * rs.st[r].words);
    max_coarse = max(max_coarse, rs.st[r].nc * rs.st[r].nc);
  }
  region_rowmask_kernel<<<dim3((max_rw + 127) / 128, NUM_REGIONS, 2), 128, 0, s>>>(rs);
  region_coarse_kernel<<<dim3((max_coarse + 127) / 128, NUM_REGIONS), 128, 0, s>>>(rs);
  region_fine_kernel<<<dim3(max_coarse, NUM_REGIONS), 256, 0, s>>>(rs);
  region_atlas_kernel<<<(ATLAS_N * ATLAS_N + 127) / 128, 128, 0, s>>>(rs);
  region_atlas2_kernel<<<ATLAS_N * ATLAS_N, 256, 0, s>>>(rs);
  return cudaGetLastError();
}

cudaError_t launch_walk_fast(const wgrt_problem_t& p, const RegionSet& rs, int* work_counter,
                             unsigned long long* counters, int num_sms, cudaStream_t s) {
  if (p.num_rays == 0) return cudaSuccess;
  int* tile_size = work_counter + 1;  // workspace layout: {tile counter, tile size}
  pick_tile_kernel<<<1, 1024, 0, s>>>(p, tile_size, work_counter);
  const int rows = 6 + 4 * static_cast<int>(p.n_FC) + 6 * static_cast<int>(p.n_OC);
  const size_t smem = walk_smem_queue_offset(rows) + (WALK_THREADS / 32) * sizeof(WarpQueue);
  const bool count = (p.flags & WGRT_FLAG_COUNTERS) != 0;
  const bool implicit = p.runner_points > 0;
  auto kern = count ? (implicit ? walk_fast_kernel<true, true> : walk_fast_kernel<true, false>)
                    : (implicit ? walk_fast_kernel<false, true> : walk_fast_kernel<false, false>);
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  int per_sm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WALK_THREADS, smem);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  const int64_t min_tiles = (p.num_rays + 31) / 32;
  const int64_t resident = static_cast<int64_t>(num_sms) * per_sm;
  const int grid = static_cast<int>(resident < min_tiles ? resident : (min_tiles > 1 ? min_tiles : 1));
  kern<<<grid, WALK_THREADS, smem, s>>>(p, rs, work_counter, tile_size, counters);
  return cudaGetLastError();
}

cudaError_t launch_seed_rng(uint32_t* states, int64_t n, int64_t first_index, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  seed_rng_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(states, n, first_index);
  return cudaGetLastError();
}

cudaError_t launch_debug_locate_grid(const RegionSet& rs, int region, const double* px, const double* py, int64_t n,
                                     int32_t* out, unsigned long long* counters, int via_atlas, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((n + 127) / 128);
  if (counters) locate_grid_kernel<true><<<blocks, 128, 0, s>>>(rs, region, px, py, n, out, counters, via_atlas);
  else locate_grid_kernel<false><<<blocks, 128, 0, s>>>(rs, region, px, py, n, out, counters, via_atlas);
  return cudaGetLastError();
}

cudaError_t launch_bins_pack_u8(const float* bins, int64_t n, uint8_t* out, unsigned* stats, float limit, int num_sms,
                                cudaStream_t s) {
  cudaError_t err = cudaMemsetAsync(stats, 0, 2 * sizeof(unsigned), s);
  if (err != cudaSuccess || n == 0) return err;
  bins_pack_u8_kernel<<<num_sms * 8, 256, 0, s>>>(reinterpret_cast<const float4*>(bins), n / 4,
                                                  reinterpret_cast<uint32_t*>(out), stats, limit);
  return cudaGetLastError();
}

cudaError_t launch_bins_unpack_u8(const uint8_t* in, int64_t n, float* bins, int num_sms, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  bins_unpack_u8_kernel<<<num_sms * 8, 256, 0, s>>>(reinterpret_cast<const uint32_t*>(in), n / 4,
                                                    reinterpret_cast<float4*>(bins));
  return cudaGetLastError();
}

cudaError_t launch_pupil_sums(const float* EB, int64_t L, int64_t Yf, int64_t Xf, int64_t EBy, int64_t EBx,
                              int mask, int step_y, int step_x, float* out, float* cell_sums, cudaStream_t s) {
  const int64_t tiles = L * Yf * Xf;
  if (tiles == 0) return cudaSuccess;
  const int n_epy = EBy >= mask ? static_cast<int>((EBy - mask) / step_y + 1) : 0;
  const int n_epx = EBx >= mask ? static_cast<int>((EBx - mask) / step_x + 1) : 0;
  const size_t tile_smem = static_cast<size_t>(EBy * EBx) * sizeof(float);
  const bool sparse_sampling = static_cast<int64_t>(n_epy) * n_epx <= 256 && tile_smem <= 200 * 1024;
  if (sparse_sampling) {
    // the reference's sampled eye positions (7 x 8 at the default size): the whole tile in shared memory
    cudaError_t err = cudaFuncSetAttribute(pupil_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(tile_smem));
    if (err != cudaSuccess) return err;
    pupil_sums_kernel<<<static_cast<unsigned>(tiles), 256, tile_smem, s>>>(EB, tiles, static_cast<int>(EBy),
                                                                          static_cast<int>(EBx), mask, step_y, step_x,
                                                                          n_epy, n_epx, (n_epy && n_epx) ? out : nullptr,
                                                                          cell_sums);
    return cudaGetLastError();
  }
  if (cell_sums) cell_sums_kernel<<<static_cast<unsigned>(tiles), 256, 0, s>>>(EB, EBy * EBx, cell_sums);
  if (out && n_epy && n_epx) {
    if (mask > 256) return cudaErrorInvalidValue;
    // block of output positions per CTA: its window of row prefix sums must fit ~96 KB
    int boy = n_epy < 16 ? n_epy : 16, box = n_epx < 64 ? n_epx : 64;
    auto window_bytes = [&](int by_, int bx_) {
      return static_cast<size_t>((by_ - 1) * step_y + mask) * ((bx_ - 1) * step_x + mask + 1) * sizeof(float);
    };
    while (window_bytes(boy, box) > 96 * 1024 && (boy > 1 || box > 1)) {
      if (box >= boy && box > 1) box = (box + 1) / 2; else boy = (boy + 1) / 2;
    }
    const size_t smem = window_bytes(boy, box);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;   // a pupil mask wider than ~220 bins
    cudaError_t err = cudaFuncSetAttribute(pupil_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    const int blocks_y = (n_epy + boy - 1) / boy, blocks_x = (n_epx + box - 1) / box;
    if (static_cast<int64_t>(blocks_y) * blocks_x > 65535) return cudaErrorInvalidValue;
    pupil_window_kernel<<<dim3(static_cast<unsigned>(tiles), static_cast<unsigned>(blocks_y * blocks_x)), 256, smem, s>>>(
        EB, static_cast<int>(EBy), static_cast<int>(EBx), mask, step_y, step_x, n_epy, n_epx, boy, box, blocks_x, out);
  }
  return cudaGetLastError();
}

}  // namespace wgrt
