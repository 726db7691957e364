// wgrt_strict.cu -- literal thread-per-ray walk (WGRT_FLAG_STRICT) and the unit-level hooks.
//
// This is the parity anchor: one thread walks one ray through the six-state machine of
// process_rays_kernel_pro_fullColor (GRTF:842-1246) evaluating the reference's expressions in the
// reference's order, scanning every polygon edge, calling the same libdevice transcendentals.
// Compiled with -fmad=false so that, like the CPU oracle, no multiply-add is contracted.
// It exists to (a) pin the fast engine on inputs too large for the CPU oracle and (b) count the
// literal algorithm's work (edge visits, straddling edges, cross products) for the roofline.
#include "wgrt_device.cuh"

namespace wgrt {

namespace {

struct Order {
  double te, tm, dl;
};

__device__ __forceinline__ cplx lut_at(const double* __restrict__ lut, int64_t entry, int32_t C, int ch) {
  const double2 v = *reinterpret_cast<const double2*>(lut + 2 * (entry * static_cast<int64_t>(C) + ch));
  return cplx{v.x, v.y};
}

template <bool COUNT>
__device__ __forceinline__ Order jones4(const double* __restrict__ lut, int64_t entry, int32_t C, int c0, int c1,
                                        int c2, int c3, double Ete, double Etm, double dl, Counts* cn) {
  const cplx q[4] = {lut_at(lut, entry, C, c0), lut_at(lut, entry, C, c1), lut_at(lut, entry, C, c2),
                     lut_at(lut, entry, C, c3)};
  Order o;
  efield_literal(Ete, Etm, dl, q, o.te, o.tm, o.dl);
  if (COUNT) cn->c[WGRT_CNT_EFIELD]++;
  return o;
}

template <bool COUNT>
__device__ void walk_one_ray(const wgrt_problem_t& p, int64_t idx, Counts* cn) {
  double x, y, Ete, Etm, dl;
  int64_t m, n, lm;
  if (p.runner_points > 0) {  // runner layout, see include/wgrt.h
    const int64_t rpc = 2 * p.runner_points, k = idx % rpc, cell = p.runner_first_cell + idx / rpc;
    const int64_t pt = k < p.runner_points ? k : k - p.runner_points;
    x = static_cast<double>(p.x[pt]);
    y = static_cast<double>(p.y[pt]);
    Ete = k < p.runner_points ? 1.0 : 0.0;
    Etm = 1.0 - Ete;
    dl = 0.0;
    lm = cell % p.L; n = (cell / p.L) % p.Y; m = cell / (p.L * p.Y);
  } else {
    x = static_cast<double>(p.x[idx]);
    y = static_cast<double>(p.y[idx]);
    m = static_cast<int64_t>(p.m[idx]);
    n = static_cast<int64_t>(p.n[idx]);
    lm = p.lmd_num ? static_cast<int64_t>(p.lmd_num[idx]) : 0;
    Ete = static_cast<double>(p.te[idx]);
    Etm = static_cast<double>(p.tm[idx]);
    dl = static_cast<double>(p.delta_phase[idx]);
  }
  if (m < 0 || m >= p.X || n < 0 || n >= p.Y || lm < 0 || lm >= p.L) return;  // outside every table
  if (p.runner_points == 0 &&
      !(isfinite(p.m[idx]) && isfinite(p.n[idx]) && (!p.lmd_num || isfinite(p.lmd_num[idx])))) return;
  uint32_t rng = p.rng_states[idx];
  double ener = 1.0;
  const double threshold = p.threshold;  // GRTF:859 (0) or GRTF:444 (1e-15)
  double gap_x = 0.0, gap_y = 0.0, cos_theta = 0.0, norm;
  int state;
  if (COUNT) cn->c[WGRT_CNT_RAYS]++;

  const int64_t cell = (lm * p.X + m) * p.Y + n;
  const int64_t cells_per_poly = p.L * p.X * p.Y;
  const double* __restrict__ T = p.lut_TIR + 4 * cell;
  const double* __restrict__ G = p.lut_gap + 8 * cell;
  const int32_t Ci = p.C_ic, Cf = p.C_fc, Co = p.C_oc;
  const double c_ic1 = cos(lut_at(p.lut_ic1, cell, Ci, 0).re);
  const double c_ic2 = cos(lut_at(p.lut_ic2, cell, Ci, 0).re);
  const double c_ic3 = cos(lut_at(p.lut_ic3, cell, Ci, 0).re);
  Order o1, o2, o3;
  double e1, e2, e3, u;

#define WGRT_TAKE(o, tir, gx, gy, costh)            \
  do {                                              \
    norm = sqrt((o).te * (o).te + (o).tm * (o).tm); \
    Ete = (o).te / norm;                            \
    Etm = (o).tm / norm;                            \
    dl = (o).dl + T[tir];                           \
    gap_x = G[gx];                                  \
    gap_y = G[gy];                                  \
    x += gap_x;                                     \
    y += gap_y;                                     \
    cos_theta = (costh);                            \
    if (COUNT) cn->c[WGRT_CNT_BOUNCES]++;           \
  } while (0)
#define WGRT_DRAW(which)                     \
  do {                                       \
    u = xorshift_draw(rng, p.ray_index_base + idx); \
    if (COUNT) {                             \
      cn->c[WGRT_CNT_DRAWS]++;               \
      cn->c[which]++;                        \
    }                                        \
  } while (0)
#define WGRT_DONE()            \
  do {                         \
    p.rng_states[idx] = rng;   \
    return;                    \
  } while (0)

  // in-coupling from air, GRTF:860-904
  o1 = jones4<COUNT>(p.lut_ic1, cell, Ci, 13, 18, 33, 38, Ete, Etm, dl, cn);
  o2 = jones4<COUNT>(p.lut_ic1, cell, Ci, 15, 20, 35, 40, Ete, Etm, dl, cn);
  e1 = (o1.te * o1.te + o1.tm * o1.tm) * c_ic2 / c_ic1 * p.n_g;
  e2 = (o2.te * o2.te + o2.tm * o2.tm) * c_ic3 / c_ic1 * p.n_g;
  WGRT_DRAW(WGRT_CNT_DRAW2);
  if (u <= e1) {
    WGRT_TAKE(o1, 0, 0, 1, c_ic2);
    ener *= e1;
    state = inside_or_on_edge_literal<COUNT>(x, y, p.IC, 0, p.IC_n, cn) ? 0 : 2;
  } else if (u <= e1 + e2) {
    WGRT_TAKE(o2, 2, 4, 5, c_ic3);
    ener *= e2;
    if (!inside_or_on_edge_literal<COUNT>(x, y, p.IC, 0, p.IC_n, cn)) WGRT_DONE();
    state = 1;
  } else {
    WGRT_DONE();
  }

  for (int it = 0; it < 100000; ++it) {  // GRTF:905
    if (COUNT) cn->c[WGRT_CNT_ITERS]++;
    if (!inside_or_on_edge_literal<COUNT>(x, y, p.eff_reg1, 0, p.eff_reg1_n, cn)) WGRT_DONE();  // GRTF:906
    if (state <= 1) {  // GRTF:908-999
      if (state == 0) {
        o1 = jones4<COUNT>(p.lut_ic2, cell, Ci, 4, 9, 24, 29, Ete, Etm, dl, cn);
        o2 = jones4<COUNT>(p.lut_ic2, cell, Ci, 6, 11, 26, 31, Ete, Etm, dl, cn);
      } else {
        o1 = jones4<COUNT>(p.lut_ic3, cell, Ci, 2, 22, 7, 27, Ete, Etm, dl, cn);  // GRTF:957-958
        o2 = jones4<COUNT>(p.lut_ic3, cell, Ci, 4, 9, 24, 29, Ete, Etm, dl, cn);
      }
      e1 = (o1.te * o1.te + o1.tm * o1.tm) * c_ic2 / cos_theta;
      e2 = (o2.te * o2.te + o2.tm * o2.tm) * c_ic3 / cos_theta;
      WGRT_DRAW(WGRT_CNT_DRAW2);
      if (u <= e1) {
        WGRT_TAKE(o1, 0, 0, 1, c_ic2);
        ener *= e1;
        state = inside_or_on_edge_literal<COUNT>(x, y, p.IC, 0, p.IC_n, cn) ? 0 : 2;
      } else if (u <= e1 + e2) {
        WGRT_TAKE(o2, 2, 4, 5, c_ic3);
        ener *= e2;
        if (!inside_or_on_edge_literal<COUNT>(x, y, p.IC, 0, p.IC_n, cn)) WGRT_DONE();
        state = 1;
      } else {
        WGRT_DONE();
      }
    } else if (state <= 3) {  // GRTF:1000-1108
      const int i = first_hit_literal<COUNT>(x, y, p.FC, p.FC_offset, p.n_FC, cn);
      if (i >= 0) {
        const int64_t e = static_cast<int64_t>(i) * cells_per_poly + cell;
        if (state == 2) {
          o1 = jones4<COUNT>(p.lut_fc1, e, Cf, 3, 6, 15, 18, Ete, Etm, dl, cn);
          o2 = jones4<COUNT>(p.lut_fc1, e, Cf, 2, 5, 14, 17, Ete, Etm, dl, cn);
        } else {
          o1 = jones4<COUNT>(p.lut_fc2, e, Cf, 4, 7, 16, 19, Ete, Etm, dl, cn);
          o2 = jones4<COUNT>(p.lut_fc2, e, Cf, 3, 6, 15, 18, Ete, Etm, dl, cn);
        }
        const double c_fc1 = cos(lut_at(p.lut_fc1, e, Cf, 0).re);
        const double c_fc2 = cos(lut_at(p.lut_fc2, e, Cf, 0).re);
        e1 = (o1.te * o1.te + o1.tm * o1.tm) * c_fc1 / cos_theta;
        e2 = (o2.te * o2.te + o2.tm * o2.tm) * c_fc2 / cos_theta;
        const double ener1 = ener * e1, ener2 = ener * e2;
        WGRT_DRAW(WGRT_CNT_DRAW2);
        if (u <= e1 && ener1 > threshold) {
          WGRT_TAKE(o1, 0, 0, 1, c_fc1);
          ener = ener1 * 1.0;
          state = 2;
        } else if (u <= e1 + e2 && ener2 > threshold) {
          WGRT_TAKE(o2, 1, 2, 3, c_fc2);
          ener = ener2 * 1.0;
          state = 3;
        } else {
          WGRT_DONE();
        }
      } else if (state == 2) {  // GRTF:1049-1052
        x += gap_x;
        y += gap_y;
        dl += 2 * T[0];
        if (COUNT) cn->c[WGRT_CNT_BOUNCES]++;
      } else {  // GRTF:1102-1108
        if (!inside_or_on_edge_literal<COUNT>(x, y, p.eff_reg2, 0, p.eff_reg2_n, cn)) {
          state = 4;
        } else {
          x += gap_x;
          y += gap_y;
          dl += 2 * T[1];
          if (COUNT) cn->c[WGRT_CNT_BOUNCES]++;
        }
      }
    } else {  // states 4, 5: GRTF:1110-1246
      const int i = first_hit_literal<COUNT>(x, y, p.OC, p.OC_offset, p.n_OC, cn);
      if (i >= 0) {
        const int64_t e = static_cast<int64_t>(i) * cells_per_poly + cell;
        if (state == 4) {
          o1 = jones4<COUNT>(p.lut_oc1, e, Co, 4, 9, 24, 29, Ete, Etm, dl, cn);
          o2 = jones4<COUNT>(p.lut_oc1, e, Co, 2, 7, 22, 27, Ete, Etm, dl, cn);
          o3 = jones4<COUNT>(p.lut_oc1, e, Co, 13, 18, 33, 38, Ete, Etm, dl, cn);
        } else {
          o1 = jones4<COUNT>(p.lut_oc2, e, Co, 6, 11, 26, 31, Ete, Etm, dl, cn);
          o2 = jones4<COUNT>(p.lut_oc2, e, Co, 4, 9, 24, 29, Ete, Etm, dl, cn);
          o3 = jones4<COUNT>(p.lut_oc2, e, Co, 15, 20, 35, 40, Ete, Etm, dl, cn);
        }
        const double c_oc1 = cos(lut_at(p.lut_oc1, e, Co, 0).re);
        const double c_oc2 = cos(lut_at(p.lut_oc2, e, Co, 0).re);
        e1 = (o1.te * o1.te + o1.tm * o1.tm) * c_oc1 / cos_theta;
        e2 = (o2.te * o2.te + o2.tm * o2.tm) * c_oc2 / cos_theta;
        e3 = (o3.te * o3.te + o3.tm * o3.tm) * c_ic1 / cos_theta / p.n_g;
        const double ener1 = ener * e1, ener2 = ener * e2, ener3 = ener * e3;
        WGRT_DRAW(WGRT_CNT_DRAW3);
        if (u <= e1 && ener1 > threshold) {
          WGRT_TAKE(o1, 1, 2, 3, c_oc1);
          ener = ener1 * 1.0;
          state = 4;
        } else if (u <= e1 + e2 && ener2 > threshold) {
          WGRT_TAKE(o2, 3, 6, 7, c_oc2);
          ener = ener2 * 1.0;
          state = 5;
        } else if (u <= e1 + e2 + e3 && ener3 > threshold) {
          const double* rect = p.eff_reg_FOV + 8 * (m * p.Y + n);  // GRTF:100-108
          if (inside_or_on_edge_literal<COUNT>(x, y, rect, 0, 4, cn)) {
            const double* r = p.eff_reg_FOV_range + 4 * (m * p.Y + n);
            deposit_bin(p, lm, m, n, x, y, r[0], r[1], r[2], r[3]);
            if (COUNT) cn->c[WGRT_CNT_DEPOSITS]++;
          }
          WGRT_DONE();
        } else {
          WGRT_DONE();
        }
      } else if (state == 4) {  // GRTF:1175-1178
        x += gap_x;
        y += gap_y;
        dl += 2 * T[1];
        if (COUNT) cn->c[WGRT_CNT_BOUNCES]++;
      } else {
        WGRT_DONE();  // GRTF:1244-1246
      }
    }
  }
  WGRT_DONE();
#undef WGRT_TAKE
#undef WGRT_DRAW
#undef WGRT_DONE
}

template <bool COUNT>
__global__ void __launch_bounds__(256) walk_strict_kernel(const __grid_constant__ wgrt_problem_t p,
                                                          unsigned long long* counters) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  Counts cn;
  if (COUNT) cn.clear();
  if (idx < p.num_rays) walk_one_ray<COUNT>(p, idx, &cn);
  if (COUNT) cn.flush(counters);
}

// The rays the fast walk left undecided (near ties, see wgrt_walk.cu): walked from their start with the
// literal expressions.  Their RNG states and bins are untouched by the fast walk, so the result is what
// a strict launch gives for them.  Counted into WGRT_CNT_NEAR_TIE by every launch.
template <bool COUNT>
__global__ void __launch_bounds__(128) walk_redo_kernel(const __grid_constant__ wgrt_problem_t p,
                                                        const RedoList* __restrict__ redo,
                                                        unsigned long long* counters) {
  const unsigned n = min(redo->count, REDO_CAP);
  Counts cn;
  if (COUNT) cn.clear();
  for (unsigned j = threadIdx.x; j < n; j += blockDim.x) {
    const int64_t idx = redo->idx[j];
    if (idx >= 0 && idx < p.num_rays) walk_one_ray<COUNT>(p, idx, &cn);
  }
  if (COUNT) {
    cn.c[WGRT_CNT_RAYS] = 0;   // the fast walk counted the ray when it loaded it
    cn.flush(counters);
  }
  if (threadIdx.x == 0 && n) atomicAdd(counters + WGRT_CNT_NEAR_TIE, static_cast<unsigned long long>(n));
}

__global__ void locate_literal_kernel(const double* verts, const int64_t* off, int64_t npoly, const double* px,
                                      const double* py, int64_t n, int32_t* out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = first_hit_literal<false>(px[i], py[i], verts, off, npoly, nullptr);
}

__global__ void efield_kernel(const double* ete, const double* etm, const double* delta, const double* jones,
                              int64_t n, double* out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cplx q[4];
  for (int k = 0; k < 4; ++k) q[k] = cplx{jones[8 * i + 2 * k], jones[8 * i + 2 * k + 1]};
  efield_literal(ete[i], etm[i], delta[i], q, out[3 * i], out[3 * i + 1], out[3 * i + 2]);
}

__global__ void xorshift_kernel(uint32_t* states, int64_t n, int draws, double* out_last) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = states[i];
  double u = 0.0;
  for (int d = 0; d < draws; ++d) u = xorshift_draw(s, i);
  states[i] = s;
  if (out_last) out_last[i] = u;
}

// Peak probes: 8 independent FMA chains per thread, long enough to be issue bound.  Explicit
// fma intrinsics: this translation unit is built with -fmad=false.
__device__ __forceinline__ double fma_t(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* sink, int iters, T a, T b) {
  T acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = static_cast<T>(threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fma_t(acc[k], a, b);
  }
  T s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += acc[k];
  if (s == static_cast<T>(-1234.5)) sink[0] = s;
}

inline unsigned blocks_for(int64_t n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

}  // namespace

cudaError_t launch_walk_strict(const wgrt_problem_t& p, unsigned long long* counters, cudaStream_t s) {
  if (p.num_rays == 0) return cudaSuccess;
  const int threads = 256;
  if (p.flags & WGRT_FLAG_COUNTERS)
    walk_strict_kernel<true><<<blocks_for(p.num_rays, threads), threads, 0, s>>>(p, counters);
  else
    walk_strict_kernel<false><<<blocks_for(p.num_rays, threads), threads, 0, s>>>(p, counters);
  return cudaGetLastError();
}

cudaError_t launch_walk_redo(const wgrt_problem_t& p, const RedoList* redo, unsigned long long* counters, cudaStream_t s) {
  if (p.flags & WGRT_FLAG_COUNTERS) walk_redo_kernel<true><<<1, 128, 0, s>>>(p, redo, counters);
  else walk_redo_kernel<false><<<1, 128, 0, s>>>(p, redo, counters);
  return cudaGetLastError();
}

cudaError_t launch_debug_locate_literal(const double* verts, const int64_t* off, int64_t npoly, const double* px,
                                        const double* py, int64_t n, int32_t* out, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  locate_literal_kernel<<<blocks_for(n, 128), 128, 0, s>>>(verts, off, npoly, px, py, n, out);
  return cudaGetLastError();
}

cudaError_t launch_debug_efield(const double* ete, const double* etm, const double* delta, const double* jones,
                                int64_t n, double* out, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  efield_kernel<<<blocks_for(n, 128), 128, 0, s>>>(ete, etm, delta, jones, n, out);
  return cudaGetLastError();
}

template <typename T>
static cudaError_t time_fma(int num_sms, double* tflops) {
  T* sink = nullptr;
  cudaError_t e = cudaMalloc(&sink, sizeof(T));
  if (e != cudaSuccess) return e;
  const int blocks = num_sms * 8, threads = 256, iters = 8192;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a);
    fma_peak_kernel<T><<<blocks, threads>>>(sink, iters, static_cast<T>(1.0000001), static_cast<T>(1e-7));
    cudaEventRecord(b);
    e = cudaEventSynchronize(b);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(sink);
  if (e != cudaSuccess) return e;
  const double flops = 2.0 * 8.0 * iters * static_cast<double>(blocks) * threads;
  *tflops = flops / (best * 1e-3) / 1e12;
  return cudaGetLastError();
}

cudaError_t launch_fma_peak(int num_sms, double* fp64_tflops, double* fp32_tflops) {
  cudaError_t e = time_fma<double>(num_sms, fp64_tflops);
  if (e != cudaSuccess) return e;
  return time_fma<float>(num_sms, fp32_tflops);
}

cudaError_t launch_debug_xorshift(uint32_t* states, int64_t n, int draws, double* out_last, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  xorshift_kernel<<<blocks_for(n, 128), 128, 0, s>>>(states, n, draws, out_last);
  return cudaGetLastError();
}

}  // namespace wgrt
