// wgrt_index.cu -- construction of the exact-equivalent region index and the atlas (wgrt_region.cuh),
// device-side RNG seeding (RUN:158) and the unit-level locate hook.
//
// The index replaces the reference's per-query polygon scans (GRTF:36-71) by table lookups whose answers
// are identical for every input point; this file builds those tables on the device from the caller's
// vertex arrays (rebuilt only when a content hash of the vertices changes).
#include "wgrt_region.cuh"

namespace wgrt {

namespace {

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


// ---------------------------------------------------------------------------------------------
// Zone form of the atlas (wgrt_device.cuh: ZoneSet): distinct atlas words -> 16-bit zone ids.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t ZONE_EMPTY = 0xFFFFFFFFu;   // never a valid atlas word (bits 24-30 are always 0)
constexpr uint32_t ZONE_SLOTS = 2u * ZONE_CAP;

__device__ __forceinline__ uint32_t zone_hash(uint32_t w) {
  w ^= w >> 16; w *= 0x7feb352dU; w ^= w >> 15; w *= 0x846ca68bU; w ^= w >> 16;
  return w & (ZONE_SLOTS - 1u);
}
__device__ __forceinline__ void zone_insert(uint32_t* __restrict__ keys, uint32_t w) {
  uint32_t h = zone_hash(w);
  for (uint32_t probe = 0; probe < ZONE_SLOTS; ++probe) {
    const uint32_t cur = keys[h];
    if (cur == w) return;
    if (cur == ZONE_EMPTY) {
      const uint32_t old = atomicCAS(keys + h, ZONE_EMPTY, w);
      if (old == ZONE_EMPTY || old == w) return;
    }
    h = (h + 1u) & (ZONE_SLOTS - 1u);
  }
}
__device__ __forceinline__ int zone_find(const uint32_t* __restrict__ keys, uint32_t w) {
  uint32_t h = zone_hash(w);
  for (uint32_t probe = 0; probe < ZONE_SLOTS; ++probe) {
    const uint32_t cur = keys[h];
    if (cur == w) return static_cast<int>(h);
    if (cur == ZONE_EMPTY) return -1;
    h = (h + 1u) & (ZONE_SLOTS - 1u);
  }
  return -1;
}

__global__ void zone_reset_kernel(const __grid_constant__ RegionSet rs) {
  if (!*rs.dirty) return;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ZONE_SLOTS) rs.zones.hash_keys[t] = ZONE_EMPTY;
}

// level 1 at the zone resolution: every cell classified against all five sets (as region_atlas_kernel does
// at the atlas resolution); the words go to a scratch grid and into the hash
__global__ void zone_level1_kernel(const __grid_constant__ RegionSet rs) {
  if (!*rs.dirty) return;
  double xmin, ymin, w, h;
  atlas_bbox(rs, xmin, ymin, w, h);
  const double w1 = w * (static_cast<double>(ATLAS_N) / ZONE_N1), h1 = h * (static_cast<double>(ATLAS_N) / ZONE_N1);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) zone_insert(rs.zones.hash_keys, ATLAS_OUTSIDE);
  if (t >= ZONE_N1 * ZONE_N1) return;
  const int iy = t / ZONE_N1, ix = t - iy * ZONE_N1;
  const double mx = margin_of(w1), my = margin_of(h1);
  uint32_t word = 0;
  for (int r = 0; r < NUM_REGIONS; ++r) {
    uint32_t detail;
    const uint8_t code = classify_cell(rs.st[r], nullptr, xmin + ix * w1 - mx, xmin + (ix + 1) * w1 + mx, ymin + iy * h1 - my,
                                       ymin + (iy + 1) * h1 + my, xmin + (ix + 0.5) * w1, ymin + (iy + 0.5) * h1, detail);
    word |= atlas_field(r, code);
    if (code == CELL_AMBIG) word |= ATLAS_ANY_MIXED;
  }
  rs.zones.words1[t] = word;
  zone_insert(rs.zones.hash_keys, word);
}

// level 2: the words the atlas already holds under its MIXED level-1 cells
__global__ void __launch_bounds__(256) zone_collect2_kernel(const __grid_constant__ RegionSet rs) {
  if (!*rs.dirty) return;
  if (!(rs.atlas[blockIdx.x] & ATLAS_ANY_MIXED)) return;
  constexpr int S = 1 << ATLAS_SUB_SHIFT;
  const int cy = blockIdx.x / ATLAS_N, cx = blockIdx.x - cy * ATLAS_N;
  const uint32_t* in = rs.atlas + ATLAS_N * ATLAS_N;
  uint32_t last = ZONE_EMPTY;
  for (int t = threadIdx.x; t < S * S; t += blockDim.x) {
    const int iy = (cy << ATLAS_SUB_SHIFT) + t / S, ix = (cx << ATLAS_SUB_SHIFT) + (t & (S - 1));
    const uint32_t wd = in[static_cast<size_t>(iy) * ATLAS_N2 + ix];
    if (wd != last) zone_insert(rs.zones.hash_keys, wd);
    last = wd;
  }
}

// dense zone ids in slot order (one block)
__global__ void __launch_bounds__(1024) zone_number_kernel(const __grid_constant__ RegionSet rs) {
  if (!*rs.dirty) return;
  __shared__ int s_cnt[1024];
  constexpr int PER = ZONE_SLOTS / 1024;
  const uint32_t* keys = rs.zones.hash_keys;
  int c = 0;
  for (int k = 0; k < PER; ++k) c += keys[threadIdx.x * PER + k] != ZONE_EMPTY;
  s_cnt[threadIdx.x] = c;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {   // inclusive scan
    const int v = threadIdx.x >= o ? s_cnt[threadIdx.x - o] : 0;
    __syncthreads();
    s_cnt[threadIdx.x] += v;
    __syncthreads();
  }
  int id = s_cnt[threadIdx.x] - c;
  const int total = s_cnt[1023];
  for (int k = 0; k < PER; ++k) {
    const int slot = threadIdx.x * PER + k;
    if (keys[slot] != ZONE_EMPTY) {
      if (id < ZONE_CAP) {
        rs.zones.hash_zone[slot] = static_cast<uint16_t>(id);
        rs.zones.words[id] = keys[slot];
      }
      ++id;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double xmin, ymin, w, h;
    atlas_bbox(rs, xmin, ymin, w, h);
    ZoneDyn d;
    d.x0 = xmin; d.y0 = ymin;
    // a power of two times the atlas' level-1 scale (exact): a point maps to the same level-2 cell in both forms
    d.inv_dx = static_cast<double>(1 << ZONE_REFINE) * (1.0 / w); d.inv_dy = static_cast<double>(1 << ZONE_REFINE) * (1.0 / h);
    d.num_zones = total;
    d.valid = total <= ZONE_CAP ? 1 : 0;
    const int so = zone_find(keys, ATLAS_OUTSIDE);
    d.outside_zone = (d.valid && so >= 0) ? rs.zones.hash_zone[so] : 0;
    d.pad_ = 0;
    *rs.zones.dyn = d;
  }
}

__global__ void zone_write1_kernel(const __grid_constant__ RegionSet rs) {
  if (!*rs.dirty || !rs.zones.dyn->valid) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ZONE_N1 * ZONE_N1) return;
  const int iy = t / ZONE_N1, ix = t - iy * ZONE_N1;
  const uint32_t wd = rs.zones.words1[t];
  // defer to level 2 only where the atlas populated it (under ITS MIXED level-1 cell); otherwise the (possibly
  // MIXED) word itself is the zone and the walk resolves it through the per-set grids
  const bool parent_mixed = (rs.atlas[(iy >> ZONE_REFINE) * ATLAS_N + (ix >> ZONE_REFINE)] & ATLAS_ANY_MIXED) != 0;
  uint16_t z = ZONE_MIXED;
  if (!((wd & ATLAS_ANY_MIXED) && parent_mixed)) z = rs.zones.hash_zone[zone_find(rs.zones.hash_keys, wd)];
  rs.zones.level1[t] = z;
}

__global__ void __launch_bounds__(256) zone_write2_kernel(const __grid_constant__ RegionSet rs) {
  if (!*rs.dirty || !rs.zones.dyn->valid) return;
  if (!(rs.atlas[blockIdx.x] & ATLAS_ANY_MIXED)) return;
  constexpr int S = 1 << ATLAS_SUB_SHIFT;
  const int cy = blockIdx.x / ATLAS_N, cx = blockIdx.x - cy * ATLAS_N;
  const uint32_t* in = rs.atlas + ATLAS_N * ATLAS_N;
  uint32_t last = ZONE_EMPTY;
  uint16_t lz = 0;
  for (int t = threadIdx.x; t < S * S; t += blockDim.x) {
    const int iy = (cy << ATLAS_SUB_SHIFT) + t / S, ix = (cx << ATLAS_SUB_SHIFT) + (t & (S - 1));
    const size_t at = static_cast<size_t>(iy) * ATLAS_N2 + ix;
    const uint32_t wd = in[at];
    if (wd != last) { lz = rs.zones.hash_zone[zone_find(rs.zones.hash_keys, wd)]; last = wd; }
    rs.zones.level2[at] = lz;
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
  __shared__ ZoneAtlas zones;
  if (threadIdx.x == 0) {
    region_load(reg, rs.st[region], rs.dyn[region]);
    const AtlasDyn ad = *rs.atlas_dyn;
    atlas.x0 = ad.x0; atlas.y0 = ad.y0; atlas.inv_dx = ad.inv_dx; atlas.inv_dy = ad.inv_dy; atlas.words = rs.atlas;
    atlas.words2 = rs.atlas + ATLAS_N * ATLAS_N;
    const ZoneDyn zd = *rs.zones.dyn;
    zones.x0 = zd.x0; zones.y0 = zd.y0; zones.inv_dx = zd.inv_dx; zones.inv_dy = zd.inv_dy;
    zones.level1 = rs.zones.level1; zones.level2 = rs.zones.level2; zones.trans = rs.zones.trans;
    zones.words = rs.zones.words; zones.outside_zone = zd.outside_zone; zones.valid = zd.valid;
  }
  __syncthreads();
  Counts cn;
  if (COUNT) cn.clear();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    if (via_atlas) {
      // via_atlas 1: the word atlas; 2: the zone grids (zone id -> word), as the production walk reads them
      const uint32_t word = (via_atlas == 2 && zones.valid) ? __ldg(zones.words + zone_lookup(zones, px[i], py[i]))
                                                            : atlas_lookup(atlas, px[i], py[i]);
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
  // the zone form of the atlas and the walk's transition table
  zone_reset_kernel<<<(ZONE_SLOTS + 255) / 256, 256, 0, s>>>(rs);
  zone_level1_kernel<<<(ZONE_N1 * ZONE_N1 + 127) / 128, 128, 0, s>>>(rs);
  zone_collect2_kernel<<<ATLAS_N * ATLAS_N, 256, 0, s>>>(rs);
  zone_number_kernel<<<1, 1024, 0, s>>>(rs);
  zone_write1_kernel<<<(ZONE_N1 * ZONE_N1 + 127) / 128, 128, 0, s>>>(rs);
  zone_write2_kernel<<<ATLAS_N * ATLAS_N, 256, 0, s>>>(rs);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  return launch_zone_transitions(rs, rs.st[REG_FC].npoly, rs.st[REG_OC].npoly, s);
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

}  // namespace wgrt
