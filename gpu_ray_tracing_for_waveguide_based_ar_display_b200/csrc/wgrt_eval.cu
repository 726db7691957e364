// wgrt_eval.cu -- consumer-side kernels on the bin tensor: the pupil-mask
// sums and per-cell totals of AR_system_evaluation_functions.py:68-109 / gpu_ray_tracing_pro_fullColor.py:186,
// and the exact uint8 packing of the bins used by the multi-GPU reduce.
#include "wgrt_device.cuh"

namespace wgrt {

namespace {

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
