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

// ---------------------------------------------------------------------------------------------
// The rest of evaluation() on the device (AR_system_evaluation_functions.py:110-160): per sampled eye
// position, the white image through the display model -- sRGB image (clip, gamma, brightness stretch),
// CIE XYZ / Lab, CIEDE2000 against D65, luminance min / max / mean over the FoV -- reduced to a handful of
// numbers per eye position.  One CTA per eye position; all arithmetic in double like the reference's NumPy.
// The colour constants arrive in wgrt_eval_params_t from the Python mirror (the reference's own matrices).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double srgb_gamma(double v) {   // EVAL:14-16
  return v <= 0.0031308 ? v * 12.92 : 1.055 * pow(v, 1.0 / 2.4) - 0.055;
}
__device__ __forceinline__ double lab_f(double t) {
  const double d = 6.0 / 29.0;
  return t > d * d * d ? cbrt(t) : t / (3.0 * d * d) + 4.0 / 29.0;
}
__device__ __forceinline__ double deg_atan2_360(double y, double x) {
  double h = atan2(y, x) * (180.0 / 3.14159265358979323846);
  h = fmod(h, 360.0);
  if (h < 0.0) h += 360.0;
  return h;
}
__device__ __forceinline__ double rad(double deg) { return deg * (3.14159265358979323846 / 180.0); }
__device__ __forceinline__ double pow7(double x) { const double x2 = x * x, x4 = x2 * x2; return x4 * x2 * x; }
// CIEDE2000 (Sharma, Wu, Dalal 2005), kL = kC = kH = 1 -- the formula colour.delta_E(..., 'CIE 2000') evaluates
__device__ double delta_e_2000(double L1, double a1, double b1, double L2, double a2, double b2) {
  const double C1 = hypot(a1, b1), C2 = hypot(a2, b2);
  const double Cm = 0.5 * (C1 + C2);
  const double c7 = pow7(Cm);
  const double G = 0.5 * (1.0 - sqrt(c7 / (c7 + 6103515625.0)));   // 25^7
  const double a1p = (1.0 + G) * a1, a2p = (1.0 + G) * a2;
  const double C1p = hypot(a1p, b1), C2p = hypot(a2p, b2);
  const double h1p = deg_atan2_360(b1, a1p), h2p = deg_atan2_360(b2, a2p);
  const double dLp = L2 - L1, dCp = C2p - C1p;
  double dh = h2p - h1p;
  const bool grey = C1p * C2p == 0.0;
  dh = grey ? 0.0 : (dh > 180.0 ? dh - 360.0 : (dh < -180.0 ? dh + 360.0 : dh));
  const double dHp = 2.0 * sqrt(C1p * C2p) * sin(rad(dh) * 0.5);
  const double Lm = 0.5 * (L1 + L2), Cpm = 0.5 * (C1p + C2p);
  const double hsum = h1p + h2p;
  const double hm = grey ? hsum : (fabs(h1p - h2p) <= 180.0 ? hsum * 0.5 : (hsum < 360.0 ? (hsum + 360.0) * 0.5 : (hsum - 360.0) * 0.5));
  const double T = 1.0 - 0.17 * cos(rad(hm - 30.0)) + 0.24 * cos(rad(2.0 * hm)) + 0.32 * cos(rad(3.0 * hm + 6.0)) -
                   0.20 * cos(rad(4.0 * hm - 63.0));
  const double q = (hm - 275.0) / 25.0;
  const double dth = 30.0 * exp(-q * q);
  const double p7 = pow7(Cpm);
  const double Rc = 2.0 * sqrt(p7 / (p7 + 6103515625.0));
  const double l50 = (Lm - 50.0) * (Lm - 50.0);
  const double Sl = 1.0 + 0.015 * l50 / sqrt(20.0 + l50);
  const double Sc = 1.0 + 0.045 * Cpm, Sh = 1.0 + 0.015 * Cpm * T;
  const double Rt = -sin(rad(2.0 * dth)) * Rc;
  const double x = dLp / Sl, y = dCp / Sc, z = dHp / Sh;
  return sqrt(x * x + y * y + z * z + Rt * y * z);
}

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T* scratch) {
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(FULL_MASK, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = scratch[0];
  for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) r = op(r, scratch[w]);
  return r;
}

// perceive: float32 [3, Yf, Xf, n_epy, n_epx] RAW pupil sums; scale = 1 / (num_rays_per_FoV * num_iter) (RUN:197).
// metrics [n_ep][WGRT_EVAL_NUM]: sum dE2000, min Y, max Y, sum Y, pixels with Y == 0, max V of the gamma image.
// image (optional): float32 [Yf, Xf, 3, n_epy, n_epx], the brightness-normalised sRGB view (EVAL:131-136).
__global__ void __launch_bounds__(256) eval_metrics_kernel(const float* __restrict__ perceive, int Yf, int Xf, int n_epy,
                                                           int n_epx, const __grid_constant__ wgrt_eval_params_t prm,
                                                           double* __restrict__ metrics, float* __restrict__ image) {
  __shared__ double scratch[8];
  const int ep = blockIdx.x, n_ep = n_epy * n_epx;
  const int npix = Yf * Xf;
  const size_t plane = static_cast<size_t>(npix) * n_ep;   // one wavelength
  double s_de = 0.0, y_min = INFINITY, y_max = -INFINITY, s_y = 0.0, zeros = 0.0, v_max = 0.0;
  for (int pix = threadIdx.x; pix < npix; pix += blockDim.x) {
    double px[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)   // channel c (R, G, B) is wavelength index 2 - c (EVAL:119-121)
      px[c] = prm.white_rgb[c] * (static_cast<double>(__ldg(perceive + (2 - c) * plane + static_cast<size_t>(pix) * n_ep + ep)) * prm.scale);
    double xyz[3], g[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double lin = prm.M[3 * r] * px[0] + prm.M[3 * r + 1] * px[1] + prm.M[3 * r + 2] * px[2];
      g[r] = srgb_gamma(fmin(fmax(lin, 0.0), 1.0));
      xyz[r] = prm.M_xyz[3 * r] * px[0] + prm.M_xyz[3 * r + 1] * px[1] + prm.M_xyz[3 * r + 2] * px[2];
    }
    v_max = fmax(v_max, fmax(g[0], fmax(g[1], g[2])));
    const double Y = xyz[1];
    double L = 0.0, a = 0.0, b = 0.0;
    if (Y != 0.0) {   // EVAL:141-147: normalise to Y = 100, then Lab; pixels with Y == 0 get Lab = 0
      const double k = 100.0 / fmax(Y, 1e-10);
      const double fx = lab_f(xyz[0] * k / prm.white_xyz[0]), fy = lab_f(xyz[1] * k / prm.white_xyz[1]),
                   fz = lab_f(xyz[2] * k / prm.white_xyz[2]);
      L = 116.0 * fy - 16.0; a = 500.0 * (fx - fy); b = 200.0 * (fy - fz);
    } else {
      zeros += 1.0;
    }
    s_de += delta_e_2000(L, a, b, prm.lab_d65[0], prm.lab_d65[1], prm.lab_d65[2]);
    y_min = fmin(y_min, Y); y_max = fmax(y_max, Y); s_y += Y;
  }
  auto add = [](double x, double y) { return x + y; };
  auto mn = [](double x, double y) { return fmin(x, y); };
  auto mx = [](double x, double y) { return fmax(x, y); };
  s_de = block_reduce(s_de, add, scratch);
  y_min = block_reduce(y_min, mn, scratch);
  y_max = block_reduce(y_max, mx, scratch);
  s_y = block_reduce(s_y, add, scratch);
  zeros = block_reduce(zeros, add, scratch);
  v_max = block_reduce(v_max, mx, scratch);
  if (threadIdx.x == 0) {
    double* o = metrics + static_cast<size_t>(ep) * WGRT_EVAL_NUM;
    o[WGRT_EVAL_SUM_DE] = s_de; o[WGRT_EVAL_Y_MIN] = y_min; o[WGRT_EVAL_Y_MAX] = y_max; o[WGRT_EVAL_Y_SUM] = s_y;
    o[WGRT_EVAL_Y_ZEROS] = zeros; o[WGRT_EVAL_V_MAX] = v_max;
  }
  if (!image) return;
  // EVAL:131-136 / 18-43: the HSV value stretch divides V, i.e. every channel, by the image's largest V
  const double inv = v_max > 0.0 ? 1.0 / v_max : 1.0;
  for (int pix = threadIdx.x; pix < npix; pix += blockDim.x) {
    double px[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      px[c] = prm.white_rgb[c] * (static_cast<double>(__ldg(perceive + (2 - c) * plane + static_cast<size_t>(pix) * n_ep + ep)) * prm.scale);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double lin = prm.M[3 * r] * px[0] + prm.M[3 * r + 1] * px[1] + prm.M[3 * r + 2] * px[2];
      image[(static_cast<size_t>(pix) * 3 + r) * n_ep + ep] = static_cast<float>(srgb_gamma(fmin(fmax(lin, 0.0), 1.0)) * inv);
    }
  }
}

}  // namespace

cudaError_t launch_eval_metrics(const float* perceive, int64_t Yf, int64_t Xf, int n_epy, int n_epx,
                                const wgrt_eval_params_t& prm, double* metrics, float* image, cudaStream_t s) {
  const int n_ep = n_epy * n_epx;
  if (n_ep <= 0 || Yf * Xf <= 0) return cudaSuccess;
  eval_metrics_kernel<<<n_ep, 256, 0, s>>>(perceive, static_cast<int>(Yf), static_cast<int>(Xf), n_epy, n_epx, prm, metrics, image);
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
