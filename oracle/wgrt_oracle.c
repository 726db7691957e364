/*
 * wgrt_oracle.c -- CPU restatement of the reference ray walk.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for libwgrt.so.  Nothing in the product path may call it:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg do.
 *
 * It restates, operation for operation, the algorithm of
 *   /root/reference/GPU_ray_tracing_functions.py   (GRTF below)
 *     get_uniform_random_number      GRTF:25-34
 *     is_inside_polygon              GRTF:36-50
 *     point_on_segment               GRTF:52-61
 *     is_inside_or_on_edge           GRTF:63-71
 *     *_4d variants                  GRTF:73-108
 *     _wrap_minus_pi_to_pi           GRTF:124-130
 *     E_field_cal                    GRTF:132-152
 *     add_to_EB_atomic_val           GRTF:154-165
 *     process_rays_kernel_pro_fullColor  GRTF:833-1246
 * in IEEE double arithmetic with glibc's libm (CPython's own math.hypot restated) and no FMA contraction (build with
 * -ffp-contract=off), which is what the reference's only CPU path -- its Numba kernels under
 * NUMBA_ENABLE_CUDASIM=1, i.e. CPython floats and `math.*` -- evaluates.
 *
 * Parity pin: tests/golden/ holds matrix_EB / rng_states produced by running the reference file
 * itself under the simulator (oracle/make_golden.py); tests/test_oracle_golden.py requires this
 * restatement to reproduce them bit for bit.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/wgrt.h"

typedef struct { double re, im; } cplx;

static inline cplx c_mul(cplx a, cplx b) {
  /* CPython / NumPy complex product: (ac - bd) + (ad + bc)i */
  cplx r;
  r.re = a.re * b.re - a.im * b.im;
  r.im = a.re * b.im + a.im * b.re;
  return r;
}
static inline cplx c_add(cplx a, cplx b) { cplx r = {a.re + b.re, a.im + b.im}; return r; }

typedef struct {
  uint64_t c[WGRT_NUM_COUNTERS];
} counters_t;

/* Per-decision trace of ONE ray (wgrt_oracle_trace_events, used by tools/triage_mismatch.py): every
 * draw with the thresholds it is compared with.  8 doubles per event:
 * state (-1 = in-coupling from air), u, e1, e1+e2, e1+e2+e3 (NaN for two-order events), x, y, ener. */
typedef struct {
  double* out;
  int64_t cap, n;
} event_trace_t;
static __thread event_trace_t* g_trace = NULL;
#define TRACE_EVENT(st, e3v)                                                                   \
  do {                                                                                         \
    if (g_trace) {                                                                             \
      if (g_trace->n < g_trace->cap) {                                                         \
        double* t_ = g_trace->out + 8 * g_trace->n;                                            \
        t_[0] = (double)(st); t_[1] = u; t_[2] = e1; t_[3] = e1 + e2; t_[4] = (e3v);           \
        t_[5] = x; t_[6] = y; t_[7] = ener;                                                    \
      }                                                                                        \
      g_trace->n++;                                                                            \
    }                                                                                          \
  } while (0)

/* GRTF:25-34 */
static inline double draw_uniform_at(uint32_t* rng_states, int64_t index, int64_t index_base, counters_t* cn) {
  uint32_t s = rng_states[index];
  /* GRTF:28-29; index_base = wgrt_problem_t.ray_index_base (0 for a plain launch) */
  if (s == 0u) s = 0x6D2B79F5u ^ (uint32_t)(index_base + index + 1);
  s ^= s << 13;
  s ^= s >> 17;
  s ^= s << 5;
  rng_states[index] = s;
  cn->c[WGRT_CNT_DRAWS]++;
  return (double)s * (1.0 / 4294967296.0);
}

/* GRTF:52-61 (and :73-82 with a different vertex source) */
static inline int on_segment(double px, double py, double x1, double y1, double x2, double y2,
                             double tol, counters_t* cn) {
  if ((px < fmin(x1, x2) - tol) || (px > fmax(x1, x2) + tol) ||
      (py < fmin(y1, y2) - tol) || (py > fmax(y1, y2) + tol))
    return 0;
  cn->c[WGRT_CNT_CROSS]++;
  return fabs((x2 - x1) * (py - y1) - (y2 - y1) * (px - x1)) <= tol;
}

/* GRTF:63-71 followed by GRTF:36-50; poly = [V,2] */
static int inside_or_on_edge(double px, double py, const double* poly, int64_t start, int64_t end,
                             counters_t* cn) {
  int64_t nv = end - start;
  cn->c[WGRT_CNT_POLY_TESTS]++;
  int64_t j = nv - 1;
  for (int64_t i = 0; i < nv; ++i) {
    const double* a = poly + 2 * (start + j);
    const double* b = poly + 2 * (start + i);
    cn->c[WGRT_CNT_EDGE_VISITS]++;
    if (on_segment(px, py, a[0], a[1], b[0], b[1], 1e-12, cn)) return 1;
    j = i;
  }
  int inside = 0;
  j = nv - 1;
  for (int64_t i = 0; i < nv; ++i) {
    double xi = poly[2 * (start + i)], yi = poly[2 * (start + i) + 1];
    double xj = poly[2 * (start + j)], yj = poly[2 * (start + j) + 1];
    cn->c[WGRT_CNT_EDGE_VISITS]++;
    if ((yi > py) != (yj > py)) {
      cn->c[WGRT_CNT_STRADDLE]++;
      if (px < (xj - xi) * (py - yi) / (yj - yi + 1e-20) + xi) inside = !inside;
    }
    j = i;
  }
  return inside;
}

/*
 * math.hypot as CPython >= 3.11 evaluates it (Modules/mathmodule.c, vector_norm with n = 2): the
 * simulator calls math.hypot at GRTF:145-146, and CPython does not forward to libm's hypot but
 * runs a compensated sum of squares with one Newton correction.  Restated so that the oracle is
 * bit-equal to the reference's CPU path at the unit level too (tests/golden/units.npz).
 */
typedef struct { double hi, lo; } dl_t;
static inline dl_t dl_fast_sum(double a, double b) {
  double x = a + b;
  double y = (a - x) + b;
  dl_t r = {x, y};
  return r;
}
static inline dl_t dl_mul(double x, double y) {
  double z = x * y;
  double zz = fma(x, y, -z);
  dl_t r = {z, zz};
  return r;
}
static double py_hypot(double a, double b) {
  double v[2] = {fabs(a), fabs(b)};
  double mx = v[0] > v[1] ? v[0] : v[1];
  if (isinf(v[0]) || isinf(v[1])) return INFINITY;
  if (isnan(v[0]) || isnan(v[1])) return NAN;
  if (mx == 0.0) return 0.0;
  int max_e;
  frexp(mx, &max_e);
  if (max_e < -1023) {
    const double dmin = 2.2250738585072014e-308;
    return dmin * py_hypot(v[0] / dmin, v[1] / dmin);
  }
  double scale = ldexp(1.0, -max_e);
  double csum = 1.0, frac1 = 0.0, frac2 = 0.0;
  for (int i = 0; i < 2; ++i) {
    double x = v[i] * scale;
    dl_t pr = dl_mul(x, x);
    dl_t sm = dl_fast_sum(csum, pr.hi);
    csum = sm.hi;
    frac1 += pr.lo;
    frac2 += sm.lo;
  }
  double h = sqrt(csum - 1.0 + (frac1 + frac2));
  dl_t pr = dl_mul(-h, h);
  dl_t sm = dl_fast_sum(csum, pr.hi);
  csum = sm.hi;
  frac1 += pr.lo;
  frac2 += sm.lo;
  double x = csum - 1.0 + (frac1 + frac2);
  h += x / (2.0 * h);
  return h / scale;
}

/* GRTF:124-130 */
static inline double wrap_pi(double x) {
  double two_pi = 2.0 * M_PI;
  x = x + M_PI;
  x = x - two_pi * floor(x / two_pi);
  x = x - M_PI;
  return x;
}

/* GRTF:132-152.  q = the four LUT entries in CALL order (E_te_te, E_te_tm, E_tm_te, E_tm_tm). */
static inline void efield(double Ete_abs, double Etm_abs, double delta, const cplx q[4],
                          double* o_te, double* o_tm, double* o_delta, counters_t* cn) {
  cplx phase = {cos(delta), sin(delta)};
  cplx te_in = {Ete_abs, 0.0};
  cplx etm = {Etm_abs, 0.0};
  cplx tm_in = c_mul(phase, etm);
  cplx a = q[0], b = q[2], c = q[1], d = q[3];
  cplx Ete_out = c_add(c_mul(a, te_in), c_mul(b, tm_in));
  cplx Etm_out = c_add(c_mul(c, te_in), c_mul(d, tm_in));
  double te_abs = py_hypot(Ete_out.re, Ete_out.im);
  double tm_abs = py_hypot(Etm_out.re, Etm_out.im);
  const double eps = 1e-20;
  double phi_te = te_abs >= eps ? atan2(Ete_out.im, Ete_out.re) : 0.0;
  double phi_tm = tm_abs >= eps ? atan2(Etm_out.im, Etm_out.re) : 0.0;
  *o_te = te_abs;
  *o_tm = tm_abs;
  *o_delta = wrap_pi(phi_tm - phi_te);
  cn->c[WGRT_CNT_EFIELD]++;
}

static inline cplx lut_at(const double* lut, int64_t entry, int32_t C, int ch) {
  const double* p = lut + 2 * (entry * (int64_t)C + ch);
  cplx r = {p[0], p[1]};
  return r;
}

typedef struct {
  double te, tm, dl; /* E_field_cal outputs of one order */
} order_t;

static inline void jones4(const double* lut, int64_t entry, int32_t C, int c0, int c1, int c2, int c3,
                          double Ete, double Etm, double dl, order_t* o, counters_t* cn) {
  cplx q[4] = {lut_at(lut, entry, C, c0), lut_at(lut, entry, C, c1), lut_at(lut, entry, C, c2),
               lut_at(lut, entry, C, c3)};
  efield(Ete, Etm, dl, q, &o->te, &o->tm, &o->dl, cn);
}

/* GRTF:154-165 with the lambda slice of GRTF:1168; EB is [L, Yf, Xf, EBy, EBx]. */
static inline void deposit(const wgrt_problem_t* p, int64_t lm, int64_t m, int64_t n, double x, double y,
                           counters_t* cn) {
  const double* r = p->eff_reg_FOV_range + 4 * (m * p->Y + n);
  double xmin = r[0], xmax = r[1], ymin = r[2], ymax = r[3];
  double dx = (xmax - xmin) / (double)p->EBx;
  double dy = (ymax - ymin) / (double)p->EBy;
  int64_t ix = (int64_t)floor((x - xmin) / dx);
  int64_t iy = (int64_t)floor((y - ymin) / dy);
  int64_t flat = (((lm * p->Y + n) * p->X + m) * p->EBy + iy) * p->EBx + ix;
  int64_t total = p->L * p->Y * p->X * p->EBy * p->EBx;
  if (flat >= 0 && flat < total) {
    /* float atomic add (worker threads share the bins); counts are exact in float32 */
    uint32_t* slot = (uint32_t*)(p->matrix_EB + flat);
    uint32_t old = __atomic_load_n(slot, __ATOMIC_RELAXED), want;
    do {
      float f;
      memcpy(&f, &old, 4);
      f += 1.0f;
      memcpy(&want, &f, 4);
    } while (!__atomic_compare_exchange_n(slot, &old, want, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  }
  cn->c[WGRT_CNT_DEPOSITS]++;
}

/* GRTF:100-108 on eff_reg_FOV[m, n, 0:4, :] */
static int inside_eyebox_rect(const wgrt_problem_t* p, int64_t m, int64_t n, double x, double y,
                              counters_t* cn) {
  const double* rect = p->eff_reg_FOV + 8 * (m * p->Y + n);
  return inside_or_on_edge(x, y, rect, 0, 4, cn);
}

static int first_hit(double x, double y, const double* poly, const int64_t* off, int64_t npoly,
                     counters_t* cn) {
  for (int64_t i = 0; i < npoly; ++i)
    if (inside_or_on_edge(x, y, poly, off[i], off[i + 1], cn)) return (int)i;
  return -1;
}

/* One ray, GRTF:842-1246. */
static void walk_ray(const wgrt_problem_t* p, int64_t idx, counters_t* cn) {
  double x = (double)p->x[idx];
  double y = (double)p->y[idx];
  int64_t m = (int64_t)p->m[idx];
  int64_t n = (int64_t)p->n[idx];
  int64_t lm = p->lmd_num ? (int64_t)p->lmd_num[idx] : 0;
  double Ete = (double)p->te[idx];
  double Etm = (double)p->tm[idx];
  double dl = (double)p->delta_phase[idx];
  double ener = 1.0;
  const double threshold = p->threshold; /* GRTF:859 (0) or GRTF:444 (1e-15) */
  double gap_x = 0.0, gap_y = 0.0;
  double cos_theta = 0.0; /* cos(theta.real) of the current direction */
  int state;
  cn->c[WGRT_CNT_RAYS]++;

  const int64_t cell = (lm * p->X + m) * p->Y + n; /* entry index of [L,X,Y,*] tables */
  const int64_t cells_per_poly = p->L * p->X * p->Y;
  const double* T = p->lut_TIR + 4 * cell;
  const double* G = p->lut_gap + 8 * cell;
  const int32_t Ci = p->C_ic, Cf = p->C_fc, Co = p->C_oc;
  const double c_ic1 = cos(lut_at(p->lut_ic1, cell, Ci, 0).re);
  const double c_ic2 = cos(lut_at(p->lut_ic2, cell, Ci, 0).re);
  const double c_ic3 = cos(lut_at(p->lut_ic3, cell, Ci, 0).re);
  order_t o1, o2, o3;
  double e1, e2, e3, u, norm;

  /* in-coupling from air, GRTF:860-904 */
  jones4(p->lut_ic1, cell, Ci, 13, 18, 33, 38, Ete, Etm, dl, &o1, cn);
  jones4(p->lut_ic1, cell, Ci, 15, 20, 35, 40, Ete, Etm, dl, &o2, cn);
  e1 = (o1.te * o1.te + o1.tm * o1.tm) * c_ic2 / c_ic1 * p->n_g;
  e2 = (o2.te * o2.te + o2.tm * o2.tm) * c_ic3 / c_ic1 * p->n_g;
  u = draw_uniform_at(p->rng_states, idx, p->ray_index_base, cn);
  cn->c[WGRT_CNT_DRAW2]++;
  TRACE_EVENT(-1, NAN);
#define TAKE(o, e, tir, gx, gy, costh)                 \
  do {                                                 \
    norm = sqrt((o).te * (o).te + (o).tm * (o).tm);    \
    Ete = (o).te / norm;                               \
    Etm = (o).tm / norm;                               \
    dl = (o).dl + T[tir];                              \
    gap_x = G[gx];                                     \
    gap_y = G[gy];                                     \
    x += gap_x;                                        \
    y += gap_y;                                        \
    cos_theta = (costh);                               \
    cn->c[WGRT_CNT_BOUNCES]++;                         \
  } while (0)
  if (u <= e1) {
    TAKE(o1, e1, 0, 0, 1, c_ic2);
    ener *= e1;
    state = inside_or_on_edge(x, y, p->IC, 0, p->IC_n, cn) ? 0 : 2;
  } else if (u <= e1 + e2) {
    TAKE(o2, e2, 2, 4, 5, c_ic3);
    ener *= e2;
    if (!inside_or_on_edge(x, y, p->IC, 0, p->IC_n, cn)) return;
    state = 1;
  } else {
    return;
  }

  for (int64_t it = 0; it < 100000; ++it) { /* GRTF:905 */
    cn->c[WGRT_CNT_ITERS]++;
    if (!inside_or_on_edge(x, y, p->eff_reg1, 0, p->eff_reg1_n, cn)) return; /* GRTF:906 */
    if (state == 0 || state == 1) {
      /* GRTF:908-999 */
      if (state == 0) {
        jones4(p->lut_ic2, cell, Ci, 4, 9, 24, 29, Ete, Etm, dl, &o1, cn);
        jones4(p->lut_ic2, cell, Ci, 6, 11, 26, 31, Ete, Etm, dl, &o2, cn);
      } else {
        jones4(p->lut_ic3, cell, Ci, 2, 22, 7, 27, Ete, Etm, dl, &o1, cn); /* GRTF:957-958 order */
        jones4(p->lut_ic3, cell, Ci, 4, 9, 24, 29, Ete, Etm, dl, &o2, cn);
      }
      e1 = (o1.te * o1.te + o1.tm * o1.tm) * c_ic2 / cos_theta;
      e2 = (o2.te * o2.te + o2.tm * o2.tm) * c_ic3 / cos_theta;
      u = draw_uniform_at(p->rng_states, idx, p->ray_index_base, cn);
      cn->c[WGRT_CNT_DRAW2]++;
      TRACE_EVENT(state, NAN);
      if (u <= e1) {
        TAKE(o1, e1, 0, 0, 1, c_ic2);
        ener *= e1;
        state = inside_or_on_edge(x, y, p->IC, 0, p->IC_n, cn) ? 0 : 2;
      } else if (u <= e1 + e2) {
        TAKE(o2, e2, 2, 4, 5, c_ic3);
        ener *= e2;
        if (!inside_or_on_edge(x, y, p->IC, 0, p->IC_n, cn)) return;
        state = 1;
      } else {
        return;
      }
    } else if (state == 2 || state == 3) {
      /* GRTF:1000-1108 */
      int i = first_hit(x, y, p->FC, p->FC_offset, p->n_FC, cn);
      if (i >= 0) {
        int64_t e = (int64_t)i * cells_per_poly + cell;
        if (state == 2) {
          jones4(p->lut_fc1, e, Cf, 3, 6, 15, 18, Ete, Etm, dl, &o1, cn);
          jones4(p->lut_fc1, e, Cf, 2, 5, 14, 17, Ete, Etm, dl, &o2, cn);
        } else {
          jones4(p->lut_fc2, e, Cf, 4, 7, 16, 19, Ete, Etm, dl, &o1, cn);
          jones4(p->lut_fc2, e, Cf, 3, 6, 15, 18, Ete, Etm, dl, &o2, cn);
        }
        double c_fc1 = cos(lut_at(p->lut_fc1, e, Cf, 0).re);
        double c_fc2 = cos(lut_at(p->lut_fc2, e, Cf, 0).re);
        e1 = (o1.te * o1.te + o1.tm * o1.tm) * c_fc1 / cos_theta;
        e2 = (o2.te * o2.te + o2.tm * o2.tm) * c_fc2 / cos_theta;
        double ener1 = ener * e1, ener2 = ener * e2;
        u = draw_uniform_at(p->rng_states, idx, p->ray_index_base, cn);
        cn->c[WGRT_CNT_DRAW2]++;
        TRACE_EVENT(state, NAN);
        if (u <= e1 && ener1 > threshold) {
          TAKE(o1, e1, 0, 0, 1, c_fc1);
          ener = ener1 * 1.0;
          state = 2;
        } else if (u <= e1 + e2 && ener2 > threshold) {
          TAKE(o2, e2, 1, 2, 3, c_fc2);
          ener = ener2 * 1.0;
          state = 3;
        } else {
          return;
        }
      } else if (state == 2) {
        x += gap_x; /* GRTF:1049-1052 */
        y += gap_y;
        dl += 2 * T[0];
        cn->c[WGRT_CNT_BOUNCES]++;
      } else {
        /* GRTF:1102-1108 */
        if (!inside_or_on_edge(x, y, p->eff_reg2, 0, p->eff_reg2_n, cn)) {
          state = 4;
        } else {
          x += gap_x;
          y += gap_y;
          dl += 2 * T[1];
          cn->c[WGRT_CNT_BOUNCES]++;
        }
      }
    } else {
      /* states 4 and 5, GRTF:1110-1246 */
      int i = first_hit(x, y, p->OC, p->OC_offset, p->n_OC, cn);
      if (i >= 0) {
        int64_t e = (int64_t)i * cells_per_poly + cell;
        if (state == 4) {
          jones4(p->lut_oc1, e, Co, 4, 9, 24, 29, Ete, Etm, dl, &o1, cn);
          jones4(p->lut_oc1, e, Co, 2, 7, 22, 27, Ete, Etm, dl, &o2, cn);
          jones4(p->lut_oc1, e, Co, 13, 18, 33, 38, Ete, Etm, dl, &o3, cn);
        } else {
          jones4(p->lut_oc2, e, Co, 6, 11, 26, 31, Ete, Etm, dl, &o1, cn);
          jones4(p->lut_oc2, e, Co, 4, 9, 24, 29, Ete, Etm, dl, &o2, cn);
          jones4(p->lut_oc2, e, Co, 15, 20, 35, 40, Ete, Etm, dl, &o3, cn);
        }
        double c_oc1 = cos(lut_at(p->lut_oc1, e, Co, 0).re);
        double c_oc2 = cos(lut_at(p->lut_oc2, e, Co, 0).re);
        e1 = (o1.te * o1.te + o1.tm * o1.tm) * c_oc1 / cos_theta;
        e2 = (o2.te * o2.te + o2.tm * o2.tm) * c_oc2 / cos_theta;
        e3 = (o3.te * o3.te + o3.tm * o3.tm) * c_ic1 / cos_theta / p->n_g;
        double ener1 = ener * e1, ener2 = ener * e2, ener3 = ener * e3;
        u = draw_uniform_at(p->rng_states, idx, p->ray_index_base, cn);
        cn->c[WGRT_CNT_DRAW3]++;
        TRACE_EVENT(state, e1 + e2 + e3);
        if (u <= e1 && ener1 > threshold) {
          TAKE(o1, e1, 1, 2, 3, c_oc1);
          ener = ener1 * 1.0;
          state = 4;
        } else if (u <= e1 + e2 && ener2 > threshold) {
          TAKE(o2, e2, 3, 6, 7, c_oc2);
          ener = ener2 * 1.0;
          state = 5;
        } else if (u <= e1 + e2 + e3 && ener3 > threshold) {
          if (inside_eyebox_rect(p, m, n, x, y, cn)) deposit(p, lm, m, n, x, y, cn);
          return;
        } else {
          return;
        }
      } else if (state == 4) {
        x += gap_x; /* GRTF:1175-1178 */
        y += gap_y;
        dl += 2 * T[1];
        cn->c[WGRT_CNT_BOUNCES]++;
      } else {
        return; /* GRTF:1244-1246 */
      }
    }
  }
#undef TAKE
}

static int check_problem(const wgrt_problem_t* p) {
  if (!p || p->num_rays < 0) return WGRT_ERR_INVALID;
  if (p->num_rays == 0) return WGRT_OK;
  if (!p->x || !p->y || !p->m || !p->n || (!p->lmd_num && p->L != 1) || !p->te || !p->tm || !p->delta_phase ||
      !p->rng_states || !p->IC || !p->FC || !p->FC_offset || !p->OC || !p->OC_offset ||
      !p->eff_reg1 || !p->eff_reg2 || !p->eff_reg_FOV || !p->eff_reg_FOV_range || !p->lut_ic1 ||
      !p->lut_ic2 || !p->lut_ic3 || !p->lut_fc1 || !p->lut_fc2 || !p->lut_oc1 || !p->lut_oc2 ||
      !p->lut_TIR || !p->lut_gap || !p->matrix_EB)
    return WGRT_ERR_INVALID;
  if (p->C_ic < 41 || p->C_fc < 20 || p->C_oc < 41) return WGRT_ERR_INVALID;
  return WGRT_OK;
}

typedef struct {
  const wgrt_problem_t* p;
  int64_t first, end;
  int64_t* next; /* shared chunk cursor */
  int* bad;
  counters_t local;
} worker_t;

static void* worker_main(void* arg) {
  worker_t* w = (worker_t*)arg;
  const wgrt_problem_t* p = w->p;
  const int64_t chunk = 256;
  for (;;) {
    int64_t lo = __atomic_fetch_add(w->next, chunk, __ATOMIC_RELAXED);
    if (lo >= w->end) break;
    int64_t hi = lo + chunk < w->end ? lo + chunk : w->end;
    for (int64_t i = lo; i < hi; ++i) {
      int64_t m = (int64_t)p->m[i], n = (int64_t)p->n[i], lm = p->lmd_num ? (int64_t)p->lmd_num[i] : 0;
      if (m < 0 || m >= p->X || n < 0 || n >= p->Y || lm < 0 || lm >= p->L) {
        __atomic_store_n(w->bad, 1, __ATOMIC_RELAXED);
        continue;
      }
      walk_ray(p, i, &w->local);
    }
  }
  return NULL;
}

/* Trace rays [first, first+count) of a HOST problem on `num_threads` threads (results do not
 * depend on the thread count: RNG state is per ray and bins are additive).  counters may be NULL. */
int wgrt_oracle_trace(const wgrt_problem_t* p, int64_t first, int64_t count, uint64_t* counters,
                      int num_threads) {
  int rc = check_problem(p);
  if (rc != WGRT_OK) return rc;
  if (first < 0 || count < 0 || first + count > p->num_rays) return WGRT_ERR_INVALID;
  if (num_threads < 1) num_threads = 1;
  if (num_threads > 256) num_threads = 256;
  int64_t next = first;
  int bad = 0;
  worker_t* ws = (worker_t*)calloc((size_t)num_threads, sizeof(worker_t));
  pthread_t* th = (pthread_t*)calloc((size_t)num_threads, sizeof(pthread_t));
  if (!ws || !th) { free(ws); free(th); return WGRT_ERR_INVALID; }
  for (int t = 0; t < num_threads; ++t) {
    ws[t].p = p; ws[t].first = first; ws[t].end = first + count; ws[t].next = &next; ws[t].bad = &bad;
  }
  if (num_threads == 1) {
    worker_main(&ws[0]);
  } else {
    for (int t = 0; t < num_threads; ++t) pthread_create(&th[t], NULL, worker_main, &ws[t]);
    for (int t = 0; t < num_threads; ++t) pthread_join(th[t], NULL);
  }
  if (counters)
    for (int t = 0; t < num_threads; ++t)
      for (int k = 0; k < WGRT_NUM_COUNTERS; ++k) counters[k] += ws[t].local.c[k];
  free(ws); free(th);
  return bad ? WGRT_ERR_INVALID : WGRT_OK;
}

/* Walk ray `idx` of a HOST problem alone and record its decisions (see event_trace_t).  Mutates the
 * ray's RNG state and the bins like a launch over that one ray.  Returns the number of events
 * (which may exceed `cap`; only the first `cap` are stored) or a negative error code. */
int64_t wgrt_oracle_trace_events(const wgrt_problem_t* p, int64_t idx, double* events, int64_t cap) {
  int rc = check_problem(p);
  if (rc != WGRT_OK) return rc;
  if (idx < 0 || idx >= p->num_rays || !events || cap < 0) return WGRT_ERR_INVALID;
  int64_t m = (int64_t)p->m[idx], n = (int64_t)p->n[idx], lm = p->lmd_num ? (int64_t)p->lmd_num[idx] : 0;
  if (m < 0 || m >= p->X || n < 0 || n >= p->Y || lm < 0 || lm >= p->L) return WGRT_ERR_INVALID;
  counters_t cn;
  memset(&cn, 0, sizeof cn);
  event_trace_t tr = {events, cap, 0};
  g_trace = &tr;
  walk_ray(p, idx, &cn);
  g_trace = NULL;
  return tr.n;
}


/* ================================================================================================
 * Legacy deterministic energy-splitting tracer: process_rays_kernel (GRTF:192-417) and
 * pack_active_to_front (GRTF:178-190), restated.  Rows are processed in index order by ONE thread, so
 * children are appended in a deterministic order (the reference's order depends on thread scheduling;
 * tests compare the rows as a multiset).
 * ================================================================================================ */
static inline void legacy_efield(const double* lut, int64_t entry, int32_t C, int c0, int c1, int c2, int c3,
                                 double Ete, double Etm, double dl, double* o_te, double* o_tm, double* o_dl) {
  counters_t cn;
  cplx q[4] = {lut_at(lut, entry, C, c0), lut_at(lut, entry, C, c1), lut_at(lut, entry, C, c2), lut_at(lut, entry, C, c3)};
  efield(Ete, Etm, dl, q, o_te, o_tm, o_dl, &cn);
}

/* GRTF:154-165: float32 atomic add of `value` at (n, m, iy, ix) */
static void legacy_deposit(const wgrt_legacy_problem_t* p, int64_t m, int64_t n, double x, double y, double value) {
  const double* r = p->eff_reg_FOV_range + 4 * (m * p->Y + n);
  double xmin = r[0], xmax = r[1], ymin = r[2], ymax = r[3];
  double dx = (xmax - xmin) / (double)p->EBx;
  double dy = (ymax - ymin) / (double)p->EBy;
  int64_t ix = (int64_t)floor((x - xmin) / dx);
  int64_t iy = (int64_t)floor((y - ymin) / dy);
  int64_t flat = ((n * p->X + m) * p->EBy + iy) * p->EBx + ix;
  int64_t total = p->Y * p->X * p->EBy * p->EBx;
  if (flat >= 0 && flat < total) p->matrix_EB[flat] += (float)value;
}

/* writes a row the way the split branches do (GRTF:252-263 ff.) */
static void legacy_write(double* v, double te, double tm, double dl, double x, double y, double gx, double gy,
                         double theta, double phi, int64_t m, int64_t n, double state) {
  v[8] = te; v[9] = tm; v[10] = dl;
  v[0] = x; v[1] = y; v[2] = gx; v[3] = gy; v[4] = theta; v[5] = phi;
  v[6] = (double)m; v[7] = (double)n; v[11] = state; v[12] = 1.0;
}

static void legacy_walk_row(const wgrt_legacy_problem_t* p, int64_t idx, uint64_t* dropped) {
  counters_t cn;
  double* v = p->vectors + WGRT_LEGACY_COLS * idx;
  if (v[12] == 0.0) return;
  double x = v[0], y = v[1], gap_x = v[2], gap_y = v[3], theta = v[4], phi = v[5];
  const int64_t m = (int64_t)v[6], n = (int64_t)v[7];
  double Ete = v[8], Etm = v[9], dl = v[10], region_state = v[11];
  if (m < 0 || m >= p->X || n < 0 || n >= p->Y) return;   /* outside every table (the reference would read out of bounds) */
  const int64_t cell = m * p->Y + n, cpp = p->X * p->Y;
  const double* T = p->lut_TIR + 4 * cell;
  const double* G = p->lut_gap + 8 * cell;
  const int32_t Ci = p->C_ic, Cf = p->C_fc, Co = p->C_oc;
  double te, tm, d2;

  if (region_state == 0.0) {   /* GRTF:222-233 */
    theta = lut_at(p->lut_ic2, cell, Ci, 0).re;   /* only .real is ever stored (GRTF:256-257) */
    phi = lut_at(p->lut_ic2, cell, Ci, 1).re;
    legacy_efield(p->lut_ic1, cell, Ci, 8, 11, 20, 23, Ete, Etm, dl, &Ete, &Etm, &dl);
    dl += T[0];
    gap_x = G[0]; gap_y = G[1];
    x += gap_x; y += gap_y;
    region_state = 1.0;
  }
#define LEGACY_CHILD(lut, e, c0, c1, c2, c3, tir, g0, dirlut, st)                                        \
  do {                                                                                                   \
    int64_t ni = (int64_t)(*p->total_ray_counter);                                                       \
    *p->total_ray_counter = (int32_t)(ni + 1);                                                           \
    if (ni < p->capacity) {                                                                              \
      legacy_efield(lut, e, Cf, c0, c1, c2, c3, Ete, Etm, dl, &te, &tm, &d2);                            \
      legacy_write(p->vectors + WGRT_LEGACY_COLS * ni, te, tm, d2 + T[tir], x + G[g0], y + G[(g0) + 1], \
                   G[g0], G[(g0) + 1], lut_at(dirlut, e, Cf, 0).re, lut_at(dirlut, e, Cf, 1).re, m, n, st); \
    } else {                                                                                             \
      ++*dropped;                                                                                        \
    }                                                                                                    \
  } while (0)

  if (region_state == 1.0) {   /* GRTF:235-286 */
    for (int64_t it = 0; it < p->max_steps; ++it) {
      if (!inside_or_on_edge(x, y, p->IC, 0, p->IC_n, &cn)) {
        for (int64_t i = 0; i < p->n_FC; ++i) {
          if (inside_or_on_edge(x, y, p->FC, p->FC_offset[i], p->FC_offset[i + 1], &cn)) {
            const int64_t e = i * cpp + cell;
            legacy_efield(p->lut_fc1, e, Cf, 3, 6, 15, 18, Ete, Etm, dl, &te, &tm, &d2);
            legacy_write(v, te, tm, d2 + T[0], x + gap_x, y + gap_y, gap_x, gap_y, theta, phi, m, n, 2.0);
            LEGACY_CHILD(p->lut_fc1, e, 4, 7, 16, 19, 1, 2, p->lut_fc2, 3.0);
            return;
          }
        }
        dl += 2 * T[0];
        x += gap_x; y += gap_y;
      } else {
        legacy_efield(p->lut_ic2, cell, Ci, 3, 6, 15, 18, Ete, Etm, dl, &Ete, &Etm, &dl);
        dl += T[0];
        x += gap_x; y += gap_y;
      }
    }
    v[12] = 0.0;   /* GRTF:284-286 */
    return;
  }

  if (region_state == 2.0 || region_state == 3.0) {   /* GRTF:288-377 */
    if (!inside_or_on_edge(x, y, p->eff_reg1, 0, p->eff_reg1_n, &cn)) { v[12] = 0.0; return; }
    for (int64_t it = 0; it < p->max_steps; ++it) {
      for (int64_t i = 0; i < p->n_FC; ++i) {
        if (inside_or_on_edge(x, y, p->FC, p->FC_offset[i], p->FC_offset[i + 1], &cn)) {
          const int64_t e = i * cpp + cell;
          if (region_state == 2.0) {
            legacy_efield(p->lut_fc1, e, Cf, 3, 6, 15, 18, Ete, Etm, dl, &te, &tm, &d2);
            legacy_write(v, te, tm, d2 + T[0], x + gap_x, y + gap_y, gap_x, gap_y, theta, phi, m, n, region_state);
            LEGACY_CHILD(p->lut_fc1, e, 4, 7, 16, 19, 1, 2, p->lut_fc2, 3.0);
          } else {
            legacy_efield(p->lut_fc2, e, Cf, 3, 6, 15, 18, Ete, Etm, dl, &te, &tm, &d2);
            legacy_write(v, te, tm, d2 + T[1], x + gap_x, y + gap_y, gap_x, gap_y, theta, phi, m, n, region_state);
            LEGACY_CHILD(p->lut_fc2, e, 2, 5, 14, 17, 0, 0, p->lut_fc1, 2.0);
          }
          return;
        }
      }
      if (!inside_or_on_edge(x, y, p->eff_reg2, 0, p->eff_reg2_n, &cn)) {
        if (region_state == 3.0) { region_state = 4.0; break; }
        v[12] = 0.0;
        return;
      }
      dl += 2 * T[0];   /* GRTF:375: T[.., 0] also for state 3 */
      x += gap_x; y += gap_y;
    }
  }
#undef LEGACY_CHILD

  if (region_state == 4.0) {   /* GRTF:378-417; nothing but the flag is ever written back from here */
    for (int64_t it = 0; it < p->max_steps; ++it) {
      if (!inside_or_on_edge(x, y, p->eff_reg1, 0, p->eff_reg1_n, &cn)) { v[12] = 0.0; return; }
      int hit = 0;
      for (int64_t i = 0; i < p->n_OC; ++i) {
        if (inside_or_on_edge(x, y, p->OC, p->OC_offset[i], p->OC_offset[i + 1], &cn)) {
          const int64_t e = i * cpp + cell;
          hit = 1;
          if (inside_or_on_edge(x, y, p->eff_reg_FOV + 8 * cell, 0, 4, &cn)) {
            legacy_efield(p->lut_oc, e, Co, 10, 13, 22, 25, Ete, Etm, dl, &te, &tm, &d2);
            const double efficiency = te * te + tm * tm;
            if (efficiency > 0) legacy_deposit(p, m, n, x, y, efficiency);
          }
          legacy_efield(p->lut_oc, e, Co, 3, 6, 15, 18, Ete, Etm, dl, &Ete, &Etm, &dl);
          dl += T[1];
          x += gap_x; y += gap_y;
          if (Ete * Ete + Etm * Etm < 0) { v[12] = 0.0; return; }
          break;
        }
      }
      if (!hit) {
        dl += 2 * T[1];
        x += gap_x; y += gap_y;
      }
    }
  }
}

int wgrt_oracle_legacy_step(const wgrt_legacy_problem_t* p, uint64_t* dropped) {
  if (!p || !p->vectors || !p->total_ray_counter || !p->matrix_EB || p->useful_count_in < 0 ||
      p->useful_count_in > p->capacity || p->C_ic < 24 || p->C_fc < 20 || p->C_oc < 26)
    return WGRT_ERR_INVALID;
  uint64_t d = 0;
  for (int64_t i = 0; i < p->useful_count_in; ++i) legacy_walk_row(p, i, &d);
  if (dropped) *dropped = d;
  return WGRT_OK;
}

/* GRTF:178-190, rows visited in index order */
int64_t wgrt_oracle_legacy_pack(const double* src, double* dst, int64_t src_len) {
  int64_t pos = 0;
  for (int64_t i = 0; i < src_len; ++i) {
    const double* r = src + WGRT_LEGACY_COLS * i;
    if (r[12] != 0.0 && r[8] * r[8] + r[9] * r[9] > 0.0) {
      memcpy(dst + WGRT_LEGACY_COLS * pos, r, sizeof(double) * WGRT_LEGACY_COLS);
      ++pos;
    }
  }
  return pos;
}

/* ---- unit-level entry points (same contracts as the wgrt_debug_* functions) --------------- */
int wgrt_oracle_locate(const double* verts, int64_t n_verts, const int64_t* offsets, int64_t n_polys,
                       const double* px, const double* py, int64_t n_points, int32_t* out) {
  (void)n_verts;
  counters_t cn;
  memset(&cn, 0, sizeof cn);
  for (int64_t i = 0; i < n_points; ++i) out[i] = first_hit(px[i], py[i], verts, offsets, n_polys, &cn);
  return WGRT_OK;
}

int wgrt_oracle_efield(const double* ete, const double* etm, const double* delta, const double* jones,
                       int64_t n, double* out) {
  counters_t cn;
  memset(&cn, 0, sizeof cn);
  for (int64_t i = 0; i < n; ++i) {
    cplx q[4];
    for (int k = 0; k < 4; ++k) {
      q[k].re = jones[8 * i + 2 * k];
      q[k].im = jones[8 * i + 2 * k + 1];
    }
    efield(ete[i], etm[i], delta[i], q, &out[3 * i], &out[3 * i + 1], &out[3 * i + 2], &cn);
  }
  return WGRT_OK;
}

int wgrt_oracle_hypot(const double* a, const double* b, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = py_hypot(a[i], b[i]);
  return WGRT_OK;
}

int wgrt_oracle_xorshift(uint32_t* states, int64_t n, int draws, double* out_last) {
  counters_t cn;
  memset(&cn, 0, sizeof cn);
  for (int64_t i = 0; i < n; ++i) {
    double u = 0.0;
    for (int d = 0; d < draws; ++d) u = draw_uniform_at(states, i, 0, &cn);
    if (out_last) out_last[i] = u;
  }
  return WGRT_OK;
}
