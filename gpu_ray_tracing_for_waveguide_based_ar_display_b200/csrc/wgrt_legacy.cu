// wgrt_legacy.cu -- the legacy deterministic energy-splitting tracer.
//
// process_rays_kernel (GRTF:192-417): one thread advances one ray row to its next fold-coupler split (the
// zero order stays in the row, the diffracted order is appended as a new row) or to its end; in the
// out-coupler zone every hit deposits the out-coupled energy |E|^2 into the eyebox bin.  Restated here with
//  * the exact-equivalent region index of wgrt_region.cuh (one atlas word per position) instead of the reference's
//    edge scans,
//  * child rows claimed with warp-aggregated atomics (one atomic per warp and split site),
//  * ray queues in SoA form ([13][capacity]: every column read and written coalesced) for the generation
//    loop of wgrt_legacy_trace_host, and in the reference's AoS form ([N][13]) for the drop-in launch,
//  * pack_active_to_front (GRTF:178-190) as a ballot / prefix-sum compaction: one atomic per warp.
// The expressions that produce ray data are the literal ones (E_field_cal via efield_literal); this
// translation unit is compiled with -fmad=false like the strict walk.
#include "wgrt_region.cuh"

namespace wgrt {

namespace {

constexpr int LC = WGRT_LEGACY_COLS;

// a ray row in either layout: AoS = the reference's vectors[N, 13]; SoA = column c at base + c * cap
template <bool SOA>
struct Row {
  double* base;
  int64_t cap, i;
  __device__ __forceinline__ double get(int c) const { return SOA ? base[c * cap + i] : base[i * LC + c]; }
  __device__ __forceinline__ void set(int c, double v) const {
    if (SOA) base[c * cap + i] = v; else base[i * LC + c] = v;
  }
};

__device__ __forceinline__ cplx lut_at(const double* __restrict__ lut, int64_t entry, int32_t C, int ch) {
  const double2 v = *reinterpret_cast<const double2*>(lut + 2 * (entry * static_cast<int64_t>(C) + ch));
  return cplx{v.x, v.y};
}

__device__ __forceinline__ void jones(const double* __restrict__ lut, int64_t entry, int32_t C, int c0, int c1, int c2, int c3,
                                      double Ete, double Etm, double dl, double& te, double& tm, double& d) {
  const cplx q[4] = {lut_at(lut, entry, C, c0), lut_at(lut, entry, C, c1), lut_at(lut, entry, C, c2), lut_at(lut, entry, C, c3)};
  efield_literal(Ete, Etm, dl, q, te, tm, d);
}

// GRTF:252-263 ff.: what the split branches store
template <bool SOA>
__device__ __forceinline__ void write_row(const Row<SOA>& r, double te, double tm, double dl, double x, double y, double gx,
                                          double gy, double theta, double phi, int64_t m, int64_t n, double state) {
  r.set(8, te); r.set(9, tm); r.set(10, dl);
  r.set(0, x); r.set(1, y); r.set(2, gx); r.set(3, gy); r.set(4, theta); r.set(5, phi);
  r.set(6, static_cast<double>(m)); r.set(7, static_cast<double>(n)); r.set(11, state); r.set(12, 1.0);
}

// Row index for a child: the lanes of the warp that split at this site claim consecutive rows with ONE atomic.
__device__ __forceinline__ int64_t claim_row(int32_t* counter) {
  const unsigned am = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs(am) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(am));
  base = __shfl_sync(am, base, leader);
  return static_cast<int64_t>(base) + __popc(am & ((1u << lane) - 1u));
}

template <bool SOA>
__global__ void __launch_bounds__(128) legacy_step_kernel(const __grid_constant__ wgrt_legacy_problem_t p,
                                                          const __grid_constant__ RegionSet rs,
                                                          unsigned long long* __restrict__ dropped) {
  const Region* __restrict__ regions = static_cast<const Region*>(rs.regions);
  __shared__ Atlas atlas;
  if (threadIdx.x == 0) {
    const AtlasDyn ad = *rs.atlas_dyn;
    atlas.x0 = ad.x0; atlas.y0 = ad.y0; atlas.inv_dx = ad.inv_dx; atlas.inv_dy = ad.inv_dy;
    atlas.words = rs.atlas;
    atlas.words2 = rs.atlas + ATLAS_N * ATLAS_N;
  }
  __syncthreads();
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= p.useful_count_in) return;
  const Row<SOA> v{p.vectors, p.capacity, idx};
  if (v.get(12) == 0.0) return;
  double x = v.get(0), y = v.get(1), gap_x = v.get(2), gap_y = v.get(3), theta = v.get(4), phi = v.get(5);
  const int64_t m = static_cast<int64_t>(v.get(6)), n = static_cast<int64_t>(v.get(7));
  double Ete = v.get(8), Etm = v.get(9), dl = v.get(10), state = v.get(11);
  if (m < 0 || m >= p.X || n < 0 || n >= p.Y) return;   // outside every table
  const int64_t cell = m * p.Y + n, cpp = p.X * p.Y;
  const double* __restrict__ T = p.lut_TIR + 4 * cell;
  const double* __restrict__ G = p.lut_gap + 8 * cell;
  const int32_t Ci = p.C_ic, Cf = p.C_fc, Co = p.C_oc;
  double te, tm, d2;

  // all five region answers for the ray's position from one atlas word (two levels; fields still MIXED there go
  // to the per-set grids / the literal edge expressions): the same exact-equivalent index the Monte-Carlo walk uses
  auto look = [&](unsigned need) {
    uint32_t w = atlas_lookup(atlas, x, y);
    if (w & ATLAS_ANY_MIXED) w = atlas_resolve<false>(w, need, regions, x, y, nullptr);
    return w;
  };
  auto in_set = [](uint32_t w, int shift) { return ((w >> shift) & 3u) == 1u; };
  auto slice_of = [](uint32_t w, int shift) { const int c = (w >> shift) & 0xff; return c == CELL_NONE ? -1 : c; };
  constexpr unsigned N_IC = 1u << REG_IC, N_R1 = 1u << REG_R1, N_R2 = 1u << REG_R2, N_FC = 1u << REG_FC, N_OC = 1u << REG_OC;
  // the diffracted order of a fold-coupler split goes to a new row (GRTF:264-281, 317-333, 351-366)
  auto child = [&](const double* lut, int64_t e, int c0, int c1, int c2, int c3, int tir, int g0, const double* dirlut,
                   double st) {
    const int64_t ni = claim_row(p.total_ray_counter);
    if (ni < p.capacity) {
      jones(lut, e, Cf, c0, c1, c2, c3, Ete, Etm, dl, te, tm, d2);
      write_row(Row<SOA>{p.vectors, p.capacity, ni}, te, tm, d2 + T[tir], x + G[g0], y + G[g0 + 1], G[g0], G[g0 + 1],
                lut_at(dirlut, e, Cf, 0).re, lut_at(dirlut, e, Cf, 1).re, m, n, st);
    } else {
      atomicAdd(dropped, 1ull);
    }
  };

  if (state == 0.0) {   // GRTF:222-233
    theta = lut_at(p.lut_ic2, cell, Ci, 0).re;
    phi = lut_at(p.lut_ic2, cell, Ci, 1).re;
    jones(p.lut_ic1, cell, Ci, 8, 11, 20, 23, Ete, Etm, dl, Ete, Etm, dl);
    dl += T[0];
    gap_x = G[0]; gap_y = G[1];
    x += gap_x; y += gap_y;
    state = 1.0;
  }

  if (state == 1.0) {   // GRTF:235-286
    for (int64_t it = 0; it < p.max_steps; ++it) {
      const uint32_t w = look(N_IC | N_FC);
      if (!in_set(w, ATLAS_SHIFT_IC)) {
        const int i = slice_of(w, ATLAS_SHIFT_FC);
        if (i >= 0) {
          const int64_t e = static_cast<int64_t>(i) * cpp + cell;
          jones(p.lut_fc1, e, Cf, 3, 6, 15, 18, Ete, Etm, dl, te, tm, d2);
          write_row(v, te, tm, d2 + T[0], x + gap_x, y + gap_y, gap_x, gap_y, theta, phi, m, n, 2.0);
          child(p.lut_fc1, e, 4, 7, 16, 19, 1, 2, p.lut_fc2, 3.0);
          return;
        }
        dl += 2 * T[0];
        x += gap_x; y += gap_y;
      } else {
        jones(p.lut_ic2, cell, Ci, 3, 6, 15, 18, Ete, Etm, dl, Ete, Etm, dl);
        dl += T[0];
        x += gap_x; y += gap_y;
      }
    }
    v.set(12, 0.0);   // GRTF:284-286
    return;
  }

  if (state == 2.0 || state == 3.0) {   // GRTF:288-377
    if (!in_set(look(N_R1), ATLAS_SHIFT_R1)) { v.set(12, 0.0); return; }
    for (int64_t it = 0; it < p.max_steps; ++it) {
      const uint32_t w = look(N_FC | N_R2);
      const int i = slice_of(w, ATLAS_SHIFT_FC);
      if (i >= 0) {
        const int64_t e = static_cast<int64_t>(i) * cpp + cell;
        if (state == 2.0) {
          jones(p.lut_fc1, e, Cf, 3, 6, 15, 18, Ete, Etm, dl, te, tm, d2);
          write_row(v, te, tm, d2 + T[0], x + gap_x, y + gap_y, gap_x, gap_y, theta, phi, m, n, state);
          child(p.lut_fc1, e, 4, 7, 16, 19, 1, 2, p.lut_fc2, 3.0);
        } else {
          jones(p.lut_fc2, e, Cf, 3, 6, 15, 18, Ete, Etm, dl, te, tm, d2);
          write_row(v, te, tm, d2 + T[1], x + gap_x, y + gap_y, gap_x, gap_y, theta, phi, m, n, state);
          child(p.lut_fc2, e, 2, 5, 14, 17, 0, 0, p.lut_fc1, 2.0);
        }
        return;
      }
      if (!in_set(w, ATLAS_SHIFT_R2)) {
        if (state == 3.0) { state = 4.0; break; }
        v.set(12, 0.0);
        return;
      }
      dl += 2 * T[0];   // GRTF:375: lut_TIR[.., 0] also for state 3
      x += gap_x; y += gap_y;
    }
  }

  if (state == 4.0) {   // GRTF:378-417; from here only the flag is ever written back
    const double* rect = p.eff_reg_FOV + 8 * cell;
    const double* rg = p.eff_reg_FOV_range + 4 * cell;
    for (int64_t it = 0; it < p.max_steps; ++it) {
      const uint32_t w = look(N_R1 | N_OC);
      if (!in_set(w, ATLAS_SHIFT_R1)) { v.set(12, 0.0); return; }
      const int i = slice_of(w, ATLAS_SHIFT_OC);
      if (i >= 0) {
        const int64_t e = static_cast<int64_t>(i) * cpp + cell;
        if (inside_or_on_edge_literal<false>(x, y, rect, 0, 4, nullptr)) {
          jones(p.lut_oc, e, Co, 10, 13, 22, 25, Ete, Etm, dl, te, tm, d2);
          const double efficiency = te * te + tm * tm;
          if (efficiency > 0) {   // GRTF:154-165: bin (n, m, iy, ix), float32 atomic add of the energy
            const double dx = (rg[1] - rg[0]) / static_cast<double>(p.EBx), dy = (rg[3] - rg[2]) / static_cast<double>(p.EBy);
            const int64_t ix = static_cast<int64_t>(floor((x - rg[0]) / dx)), iy = static_cast<int64_t>(floor((y - rg[2]) / dy));
            const int64_t flat = ((n * p.X + m) * p.EBy + iy) * p.EBx + ix;
            if (flat >= 0 && flat < p.Y * p.X * p.EBy * p.EBx) atomicAdd(p.matrix_EB + flat, static_cast<float>(efficiency));
          }
        }
        jones(p.lut_oc, e, Co, 3, 6, 15, 18, Ete, Etm, dl, Ete, Etm, dl);
        dl += T[1];
        x += gap_x; y += gap_y;
        if (Ete * Ete + Etm * Etm < 0) { v.set(12, 0.0); return; }
      } else {
        dl += 2 * T[1];
        x += gap_x; y += gap_y;
      }
    }
  }
}

// GRTF:178-190 as a ballot / prefix-sum compaction: a warp's surviving rows take consecutive slots of dst,
// claimed with one atomic per warp.
template <bool SOA>
__global__ void __launch_bounds__(256) legacy_pack_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                                          int64_t src_len, int64_t src_cap, int64_t dst_cap,
                                                          int32_t* __restrict__ out_count) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool keep = false;
  const Row<SOA> r{const_cast<double*>(src), src_cap, i};
  if (i < src_len) {
    const double te = r.get(8), tm = r.get(9);
    keep = r.get(12) != 0.0 && te * te + tm * tm > 0.0;
  }
  const unsigned mask = __ballot_sync(FULL_MASK, keep);
  if (!mask) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(out_count, __popc(mask));
  base = __shfl_sync(FULL_MASK, base, 0);
  if (keep) {
    const Row<SOA> d{dst, dst_cap, static_cast<int64_t>(base) + __popc(mask & ((1u << lane) - 1u))};
#pragma unroll
    for (int c = 0; c < LC; ++c) d.set(c, r.get(c));
  }
}

// AoS [n, 13] <-> SoA [13][cap]
__global__ void __launch_bounds__(256) legacy_transpose_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n,
                                                               int64_t cap, int to_soa) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * LC) return;
  const int64_t i = t / LC;
  const int c = static_cast<int>(t - i * LC);
  if (to_soa) dst[c * cap + i] = src[t]; else dst[t] = src[c * cap + i];
}

}  // namespace

cudaError_t launch_legacy_step(const wgrt_legacy_problem_t& p, const RegionSet& rs, bool soa, unsigned long long* dropped,
                               cudaStream_t s) {
  if (p.useful_count_in <= 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((p.useful_count_in + 127) / 128);
  if (soa) legacy_step_kernel<true><<<blocks, 128, 0, s>>>(p, rs, dropped);
  else legacy_step_kernel<false><<<blocks, 128, 0, s>>>(p, rs, dropped);
  return cudaGetLastError();
}

cudaError_t launch_legacy_pack(const double* src, double* dst, int64_t src_len, int64_t src_cap, int64_t dst_cap, bool soa,
                               int32_t* out_count, cudaStream_t s) {
  if (src_len <= 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>((src_len + 255) / 256);
  if (soa) legacy_pack_kernel<true><<<blocks, 256, 0, s>>>(src, dst, src_len, src_cap, dst_cap, out_count);
  else legacy_pack_kernel<false><<<blocks, 256, 0, s>>>(src, dst, src_len, src_cap, dst_cap, out_count);
  return cudaGetLastError();
}

cudaError_t launch_legacy_transpose(const double* src, double* dst, int64_t n, int64_t cap, bool to_soa, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  legacy_transpose_kernel<<<static_cast<unsigned>((n * LC + 255) / 256), 256, 0, s>>>(src, dst, n, cap, to_soa ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace wgrt
