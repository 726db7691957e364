/*
 * wgrt.h -- C ABI of the B200 waveguide ray-propagation engine (libwgrt.so).
 *
 * Drop-in boundary for ONE path of yefuzhang/GPU_ray_tracing_for_waveguide_based_AR_display:
 * the Monte-Carlo ray walk `process_rays_kernel_pro_fullColor`
 * (reference: GPU_ray_tracing_functions.py:833-1246) as launched by the runner
 * (reference: gpu_ray_tracing_pro_fullColor.py:168-178).
 *
 * Plain pointers and sizes only; no torch / numba types.  All arrays are C-contiguous.
 * Complex LUT entries are complex128 stored as (re, im) double pairs, i.e. a NumPy
 * complex128 array can be passed as `const double*` unchanged.
 */
#ifndef WGRT_H_
#define WGRT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WGRT_VERSION 100 /* major*10000 + minor*100 + patch */

/* ---- error codes (negative); wgrt_last_error() gives the text ------------------------- */
#define WGRT_OK 0
#define WGRT_ERR_INVALID (-1)   /* bad argument: null pointer, inconsistent shapes, ...   */
#define WGRT_ERR_CUDA (-2)      /* a CUDA runtime call failed                              */
#define WGRT_ERR_NO_DEVICE (-3) /* no CUDA device / driver                                 */
#define WGRT_ERR_UNSUPPORTED (-4)

/* ---- flags ---------------------------------------------------------------------------- */
#define WGRT_FLAG_STRICT 0x1u   /* literal thread-per-ray walk (parity anchor; slow)       */
#define WGRT_FLAG_COUNTERS 0x2u /* accumulate the device event counters (see below)       */
#define WGRT_FLAG_BINS_ZERO 0x4u /* host entry only: matrix_EB starts at zero (as the runner
                                  * creates it, RUN:37) -> clear it on the device instead of
                                  * uploading it */
#define WGRT_FLAG_BINS_DEVICE 0x8u /* host entry only: matrix_EB is a DEVICE pointer (all other
                                   * buffers stay host pointers): the launches accumulate into the
                                   * caller's device tensor and nothing is downloaded -- for callers
                                   * that reduce the bins over several GPUs before reading them */

#define WGRT_FLAG_BINS_COLUMNS 0x10u /* host entry, runner layout only: move only the FoV-x columns of matrix_EB
                                    * that the launch's cell range touches (instead of the whole tensor); the
                                    * rest of the caller's host array is neither read nor written.  For
                                    * multi-GPU jobs partitioned by cell range: every rank passes a full-shape
                                    * host array and gets its own columns back (multi_gpu.trace_partitioned) */

/* ---- event counters (uint64 each) ------------------------------------------------------ */
enum {
  WGRT_CNT_RAYS = 0,     /* rays launched                                                   */
  WGRT_CNT_BOUNCES,      /* position advances x += gap (a "bounce")   */
  WGRT_CNT_DRAWS,        /* xorshift32 draws                                                */
  WGRT_CNT_DRAW2,        /* two-order decisions                                             */
  WGRT_CNT_DRAW3,        /* three-order decisions                                           */
  WGRT_CNT_EFIELD,       /* Jones-matrix applications (E_field_cal evaluations)             */
  WGRT_CNT_ITERS,        /* iterations of the walk loop                                     */
  WGRT_CNT_DEPOSITS,     /* rays counted into an eyebox bin                                 */
  WGRT_CNT_POLY_TESTS,   /* point-in-region queries (one per polygon ring)                  */
  WGRT_CNT_EDGE_VISITS,  /* polygon edges visited by the literal scan (strict/oracle only)  */
  WGRT_CNT_STRADDLE,     /* edges with (yi>y)!=(yj>y): intersection arithmetic needed       */
  WGRT_CNT_CROSS,        /* on-segment cross products evaluated                             */
  WGRT_CNT_EXACT_FALLBACK, /* region queries that left the cell grid for the exact edge scan */
  WGRT_CNT_WARP_STEPS,   /* fast walk: loop iterations executed per WARP (iters / this = live lanes) */
  WGRT_CNT_WARP_BATCHES, /* fast walk: 32-ray in-coupling batches */
  WGRT_CNT_NEAR_TIE,     /* fast walk, counted by EVERY launch (flag or not): rays whose draw came within 1e-10
                          * of a decision threshold and were therefore re-walked with the literal expressions */
  WGRT_NUM_COUNTERS = 16
};

/*
 * One launch of the full-colour ray walk.  Field order follows the 33 positional arguments of
 * the reference kernel (GPU_ray_tracing_functions.py:834-841); sizes follow each pointer.
 *
 * For wgrt_trace_fullcolor every pointer is a DEVICE pointer on the current device.
 * For wgrt_trace_fullcolor_host (and the CPU oracle, which reuses this struct) every pointer is
 * a HOST pointer.
 */
typedef struct wgrt_problem {
  /* rays: float32[num_rays] each.  gap_x, gap_y, pol, azi are accepted for signature
   * compatibility; the reference overwrites them before any use (GPU_ray_tracing_functions.py:
   * 848-851 vs 872-879), so they are never read and may be NULL. */
  const float* x;
  const float* y;
  const float* gap_x;
  const float* gap_y;
  const float* pol;
  const float* azi;
  const float* m;       /* FoV-x index, integer valued */
  const float* n;       /* FoV-y index, integer valued */
  const float* lmd_num; /* wavelength index, integer valued; NULL = 0 for every ray (needs L == 1) */
  const float* te;
  const float* tm;
  const float* delta_phase;
  uint32_t* rng_states; /* uint32[num_rays], read and written */
  int64_t num_rays;

  /* geometry: float64 [V,2] vertex rings; *_offset int64[n_polys+1] */
  const double* IC;
  int64_t IC_n;
  const double* FC;
  int64_t FC_n;
  const int64_t* FC_offset;
  int64_t n_FC; /* number of fold-coupler polygons = len(FC_offset)-1 */
  const double* OC;
  int64_t OC_n;
  const int64_t* OC_offset;
  int64_t n_OC;
  double n_g;
  const double* eff_reg1;
  int64_t eff_reg1_n;
  const double* eff_reg2;
  int64_t eff_reg2_n;
  const double* eff_reg_FOV;       /* [X, Y, 4, 2] */
  const double* eff_reg_FOV_range; /* [X, Y, 4] = xmin, xmax, ymin, ymax */

  /* RCWA tables, complex128 */
  const double* lut_ic1; /* [L, X, Y, C_ic] */
  const double* lut_ic2;
  const double* lut_ic3;
  const double* lut_fc1; /* [n_FC, L, X, Y, C_fc] */
  const double* lut_fc2;
  const double* lut_oc1; /* [n_OC, L, X, Y, C_oc] */
  const double* lut_oc2;
  int32_t C_ic; /* channels per entry: >= 41 */
  int32_t C_fc; /* >= 20 */
  int32_t C_oc; /* >= 41 */
  int32_t reserved0;
  const double* lut_TIR; /* [L, X, Y, 4] */
  const double* lut_gap; /* [L, X, Y, 8] */
  int64_t L;             /* wavelengths */
  int64_t X;             /* FoV-x cells */
  int64_t Y;             /* FoV-y cells */

  /* bins: float32 [L, Y, X, EBy, EBx], accumulated in place */
  float* matrix_EB;
  int64_t EBy;
  int64_t EBx;

  uint32_t flags;
  uint32_t tile_hint; /* 0 = automatic; otherwise rays per work tile of the fast path */

  /*
   * Runner layout (optional).  The reference runner builds its twelve
   * ray arrays from P = num_rays_per_FoV/2 start points by a fixed rule
   * (gpu_ray_tracing_pro_fullColor.py:82-115): cells in the order FoV-x outer, FoV-y, wavelength
   * inner; per cell P TE rays (te=1, tm=0) then P TM rays (te=0, tm=1), all with delta_phase = 0,
   * ray k of either half starting at point k.  With runner_points = P > 0 the engine derives every
   * ray from that rule instead of reading arrays: x and y hold the P start points (float32, as the
   * runner stores them), m / n / lmd_num / te / tm / delta_phase are ignored (may be NULL), ray i
   * of this launch belongs to cell runner_first_cell + i / (2P), and num_rays must be a multiple of
   * 2P.  Results are bit-identical to a launch on the materialised arrays.
   * In wgrt_trace_fullcolor_host, rng_states may then be NULL: the states are seeded on the device
   * as the runner does (RUN:158, global ray index = runner_first_cell * 2P + i) and not returned.
   */
  int64_t runner_points;
  int64_t runner_first_cell;

  /*
   * Energy gate of the fold / out-coupler branches, `ener_k > threshold`
   * (GPU_ray_tracing_functions.py:1020 ff.).  0 for process_rays_kernel_pro_fullColor (GRTF:859);
   * 1e-15 for its single-wavelength twin process_rays_kernel_pro (GRTF:444), which is otherwise
   * the same walk with L = 1: pass lmd_num = NULL (every ray has wavelength index 0), LUTs
   * [X, Y, C] as [1, X, Y, C] and bins [Y, X, EBy, EBx] as [1, Y, X, EBy, EBx].
   */
  double threshold;

  /*
   * Index of ray 0 of this launch in the caller's numbering of the whole job.  Only the reference's
   * zero-state rule reads a ray's index: `if s == 0: s = 0x6D2B79F5 ^ (idx + 1)`
   * (GPU_ray_tracing_functions.py:28-29, idx = thread index of the launch).  A launch that covers
   * rays [b, b + num_rays) of a larger job (a pipeline chunk, a multi-GPU shard) passes b here so
   * that a zero RNG state reseeds exactly as in the single launch over the whole job.  Normally 0.
   */
  int64_t ray_index_base;

  /*
   * Host entry, runner layout, rng_states == NULL only: added to the ray index in the device-side
   * seeding rule, state[i] = 0x9E3779B9 * (runner_first_cell * 2P + i + rng_seed_offset + 1) -- so that
   * replicas of one job (weak scaling: more Monte-Carlo samples per FoV) walk independent streams
   * without uploading RNG arrays.  Normally 0 (the runner's seeds, RUN:158).
   */
  int64_t rng_seed_offset;
} wgrt_problem_t;

/* Library / runtime ------------------------------------------------------------------------ */
int wgrt_version(void);
/* sizeof(wgrt_problem_t) as compiled into the library: ABI guard for FFI bindings */
int wgrt_problem_size(void);
const char* wgrt_last_error(void);
/* number of CUDA devices visible, or WGRT_ERR_NO_DEVICE */
int wgrt_device_count(void);
/* free cached device workspaces of the calling thread's current device */
int wgrt_release(void);

/*
 * Replaces: GRTF.process_rays_kernel_pro_fullColor[grid, block](...33 args...)
 * (reference call site gpu_ray_tracing_pro_fullColor.py:170-177).  Asynchronous on `stream`
 * (a cudaStream_t; NULL = legacy default stream, which is what Numba launches on, so the
 * runner's cuda.synchronize() at gpu_ray_tracing_pro_fullColor.py:178 still fences it).
 * Mutates only rng_states and matrix_EB, like the reference kernel.
 */
int wgrt_trace_fullcolor(const wgrt_problem_t* dev_problem, void* stream);

/*
 * Replaces the runner's region gpu_ray_tracing_pro_fullColor.py:145-185 for HOST buffers:
 * H2D of rays/geometry/LUTs/bins, `num_iter` launches (RUN:169-177), synchronize, D2H of
 * matrix_EB and rng_states.  Synchronous.
 *
 * The three steps run as a pipeline over chunks of the job on three internal streams, so the PCIe
 * transfers hide under the walk: with the runner layout a chunk is a range of FoV-x columns (their
 * LUT / table slices go up, their rays are walked num_iter times, their matrix_EB slice comes down
 * while the next columns are walked); with explicit ray arrays a chunk is a range of rays and the
 * bins come down after the last one.  Rays are independent and own their RNG stream, so the result
 * is bit-identical to num_iter launches over the whole ray set.  WGRT_HOST_CHUNKS=<n> in the
 * environment forces the number of chunks (default: by job size, at most 25).
 * `timings_ms`, if not NULL, receives the device-event SPANS {first H2D start -> last H2D end,
 * first launch start -> last launch end, first D2H start -> last D2H end} in milliseconds; the
 * spans overlap, their sum exceeds the wall time.
 */
int wgrt_trace_fullcolor_host(const wgrt_problem_t* host_problem, int num_iter, float* timings_ms);

/*
 * wgrt_trace_fullcolor_host followed, on the device, by the evaluation reductions of
 * wgrt_eval_pupil_sums (below) -- the runner's region gpu_ray_tracing_pro_fullColor.py:145-198 up to
 * the pupil-mask sums (AR_system_evaluation_functions.py:68-109) and per-cell totals (RUN:186) in one
 * call.  `perceive` [L, Y, X, n_epy, n_epx] and `cell_sums` [L, Y, X] are HOST float32 arrays (either
 * may be NULL) and hold sums of the RAW counts (the reference divides the bins by num_rays_per_FoV *
 * num_iter first, RUN:197; the sums are linear).  host_problem->matrix_EB may be NULL: the bins then
 * start at zero, never leave the device, and only the reductions (a few MB instead of the 864 MB
 * bin tensor at the default size) come back.
 */
int wgrt_trace_evaluate_host(const wgrt_problem_t* host_problem, int num_iter, int mask_size, int step_y,
                             int step_x, float* perceive, float* cell_sums, float* timings_ms);

/*
 * The runner's RNG seeding rule on the device (gpu_ray_tracing_pro_fullColor.py:158):
 * dev_states[i] = 0x9E3779B9 * (first_index + i + 1) mod 2^32, i in [0, n).  Asynchronous on `stream`.
 */
int wgrt_seed_rng(uint32_t* dev_states, int64_t n, int64_t first_index, void* stream);

/* Counters accumulated by launches that carried WGRT_FLAG_COUNTERS (device-wide, since reset). */
int wgrt_counters_read(uint64_t* out, int n);
int wgrt_counters_reset(void);

/*
 * Unit-level entry points used by the parity tests (each mirrors one reference device function).
 * All pointers are HOST pointers; the call copies, runs on the GPU, and copies back.
 */

/* is_inside_or_on_edge over a set of rings (GPU_ray_tracing_functions.py:36-71): out[i] = index
 * of the first ring containing point i, or -1.  mode 0 = literal scan, mode 1 = cell-grid index
 * + exact row-masked fallback, mode 2 = through the word atlas, mode 3 = through the zone grids (what the
 * production walk reads). */
int wgrt_debug_locate(const double* verts, int64_t n_verts, const int64_t* offsets, int64_t n_polys,
                      const double* px, const double* py, int64_t n_points, int32_t* out, int mode);

/* E_field_cal (GPU_ray_tracing_functions.py:132-152) on n inputs: jones = [n,4] complex128 in
 * the CALL order (E_te_te, E_te_tm, E_tm_te, E_tm_tm); out = [n,3] (Ete_abs, Etm_abs, delta). */
int wgrt_debug_efield(const double* ete, const double* etm, const double* delta, const double* jones,
                      int64_t n, double* out);

/* get_uniform_random_number (GPU_ray_tracing_functions.py:25-34): advance each state `draws`
 * times; out_last[i] = last uniform of stream i. */
int wgrt_debug_xorshift(uint32_t* states, int64_t n, int draws, double* out_last);

/* is_inside_or_on_edge_4d on one eyebox rectangle (GPU_ray_tracing_functions.py:73-108; rect = the 4 x 2
 * vertices of eff_reg_FOV[m, n]): out[i] bit 0 = point i is inside or on an edge.  mode 0 = the literal
 * two-pass test, mode 1 = the walk's production test (accept / reject shortcut for points clearly inside /
 * outside an exactly axis-aligned rectangle, literal otherwise); with mode 1, bit 1 of out[i] tells that
 * the shortcut was armed for this rectangle. */
int wgrt_debug_deposit_inside(const double* rect, const double* px, const double* py, int64_t n_points,
                              int32_t* out, int mode);

/* Checked build only (libwgrt_checked.so, compiled with -DWGRT_CHECKED; WGRT_ERR_UNSUPPORTED otherwise):
 * number of violated bounds assertions in the production walk since the last reset -- every index it forms
 * into shared memory, the Jones scratch, the atlas / region grids and the ray arrays is asserted there.
 * Synchronises the device.  (Stand-in for compute-sanitizer memcheck, which the GPU pool does not allow.) */
int wgrt_debug_check_failures(uint64_t* out, int reset);

/* Near-tie tolerance of the fast walk (default 1e-10; negative restores it): a ray whose uniform draw
 * lands within `tol` of a threshold it is compared with is re-walked with the reference's literal
 * expressions instead of being decided by the reformulated arithmetic.  Tests widen it (e.g. 0.05) to
 * push a large share of the rays through that path; results must not change.  Process-wide. */
int wgrt_debug_set_tie_tolerance(double tol);

/*
 * Roofline denominators measured on the current device: achieved FP64 FMA rate of a pure
 * dependent-chain DFMA kernel that fills every SM (TFLOP/s, 2 flops per FMA) and the same for FP32.
 * Used by bench.py because MEASURED_PEAKS.json records HBM and tensor peaks only.
 */
int wgrt_debug_fma_peak(double* fp64_tflops, double* fp32_tflops);

/*
 * Evaluation reductions on the bin tensor:
 *   out[l, fy, fx, iy, ix] = sum over the disc pupil mask of EB[l, fy, fx, y0:y0+mask, x0:x0+mask],
 *   y0 = iy*step_y, x0 = ix*step_x   (AR_system_evaluation_functions.py:68-109; the reference uses
 *   mask_size 30, step_y 8, step_x 12), and
 *   cell_sums[l, fy, fx] = sum of EB[l, fy, fx, :, :]   (gpu_ray_tracing_pro_fullColor.py:186).
 * EB is float32 [L, Yf, Xf, EBy, EBx]; out is float32 [L, Yf, Xf, n_epy, n_epx] with
 * n_ep* = (EB* - mask_size) / step_* + 1; either output may be NULL.
 * Any step down to 1 (the full pupil convolution the reference leaves commented out as too slow,
 * AR_system_evaluation_functions.py:75-89) and eyebox tiles of any size (BASELINE config 4: 320 x 480
 * bins) are supported: dense sampling runs on row prefix sums, which are exact for the integer counts
 * the walk produces (fewer than 2^24 per tile); mask_size <= 220.
 * wgrt_eval_pupil_sums takes DEVICE pointers and is asynchronous on `stream`;
 * wgrt_eval_pupil_sums_host takes HOST pointers and is synchronous.
 */
int wgrt_eval_pupil_sums(const float* dev_EB, int64_t L, int64_t Yf, int64_t Xf, int64_t EBy, int64_t EBx,
                         int mask_size, int step_y, int step_x, float* dev_out, float* dev_cell_sums,
                         void* stream);
int wgrt_eval_pupil_sums_host(const float* EB, int64_t L, int64_t Yf, int64_t Xf, int64_t EBy, int64_t EBx,
                              int mask_size, int step_y, int step_x, float* out, float* cell_sums);

/*
 * The remainder of evaluation() on the device (AR_system_evaluation_functions.py:110-160; SURVEY.md section 8,
 * row f1): for every sampled eye position, the white image through the display model -- sRGB view (clip, gamma,
 * brightness stretch), CIE XYZ -> Lab, CIEDE2000 against D65, luminance min / max / mean over the FoV --
 * reduced on the device to WGRT_EVAL_NUM doubles per eye position, from which the caller forms
 *   delta_e = mean_ep(sum_dE / (Yf * Xf)),  U_fov = sum over eye positions without a zero-luminance pixel of
 *   Y_min / Y_max, divided by n_ep,  U_EB = min_ep / max_ep of (zero pixel ? 0 : Y_sum / (Yf * Xf)).
 * The colour constants are the reference's own (EVAL:47-63) and arrive from the host mirror; CIE Lab and
 * CIEDE2000 are restated from the published formulas (colour-science is the reference's dependency for them,
 * unpinned there and absent here: "parity unpinned" for delta_e).  L must be 3 (R, G, B = wavelength 2, 1, 0).
 */
enum {
  WGRT_EVAL_SUM_DE = 0, /* sum over the FoV of the CIEDE2000 colour difference to D65               */
  WGRT_EVAL_Y_MIN,      /* min over the FoV of the luminance Y                                       */
  WGRT_EVAL_Y_MAX,
  WGRT_EVAL_Y_SUM,
  WGRT_EVAL_Y_ZEROS,    /* number of FoV pixels with Y == 0                                          */
  WGRT_EVAL_V_MAX,      /* largest gamma-encoded channel value of the view (its brightness stretch)  */
  WGRT_EVAL_NUM = 8
};
typedef struct wgrt_eval_params {
  double scale;        /* multiplies the raw pupil sums: 1 / (num_rays_per_FoV * num_iter) (RUN:197)            */
  double white_rgb[3]; /* M^-1 @ linearised white (EVAL:112-116)                                               */
  double M[9];         /* wavelength weights -> linear sRGB, row major (EVAL:47-49)                             */
  double M_xyz[9];     /* wavelength weights -> CIE XYZ (EVAL:55-57)                                            */
  double white_xyz[3]; /* Lab reference white                                                                  */
  double lab_d65[3];   /* Lab of the D65 illuminant (EVAL:60-63)                                               */
} wgrt_eval_params_t;
/* dev_perceive: float32 [3, Yf, Xf, n_epy, n_epx] raw pupil sums (wgrt_eval_pupil_sums output); dev_metrics:
 * double [n_epy * n_epx, WGRT_EVAL_NUM]; dev_image: float32 [Yf, Xf, 3, n_epy, n_epx] or NULL.  Asynchronous. */
int wgrt_eval_metrics(const float* dev_perceive, int64_t Yf, int64_t Xf, int n_epy, int n_epx,
                      const wgrt_eval_params_t* params, double* dev_metrics, float* dev_image, void* stream);
/* The same for HOST arrays (synchronous). */
int wgrt_eval_metrics_host(const float* perceive, int64_t Yf, int64_t Xf, int n_epy, int n_epx,
                           const wgrt_eval_params_t* params, double* metrics, float* image);
/*
 * wgrt_trace_evaluate_host with the evaluation finished on the device: K launches, pupil sums, per-cell
 * totals and wgrt_eval_metrics; only `metrics` [n_ep, WGRT_EVAL_NUM] (doubles), `cell_sums` [L, Y, X] and, if
 * not NULL, `perceive` / `image` come back.  params->scale is used as given.
 */
int wgrt_trace_evaluate_metrics_host(const wgrt_problem_t* host_problem, int num_iter, int mask_size, int step_y,
                                     int step_x, const wgrt_eval_params_t* params, double* metrics, float* cell_sums,
                                     float* perceive, float* image, float* timings_ms);

/*
 * Multi-GPU reduce of the bins (new; the reference is single GPU).  The bins are small integer counts in
 * float32; summing them over the ranks as uint8 (four to an int32 word) is exact whenever every entry of
 * every rank is an integer in [0, limit] with world_size * limit <= 255, and moves a quarter of the bytes
 * over NVLink.  wgrt_bins_pack_u8 converts n device floats (n % 4 == 0) to uint8 and writes stats[0] = bit
 * pattern of the largest entry (as float), stats[1] = 1 if any entry is not an integer in [0, limit];
 * wgrt_bins_unpack_u8 converts back.  Asynchronous on `stream`.  The NCCL call and the fallback are the
 * caller's (multi_gpu.reduce_bins carries the flag in a trailing word of the same all-reduce).
 */
int wgrt_bins_pack_u8(const float* dev_bins, int64_t n, uint8_t* dev_out, uint32_t* dev_stats, float limit,
                      void* stream);
int wgrt_bins_unpack_u8(const uint8_t* dev_in, int64_t n, float* dev_bins, void* stream);

/* ======================================================================================================
 * Legacy deterministic energy-splitting tracer
 *
 * Replaces GRTF.process_rays_kernel (GPU_ray_tracing_functions.py:192-417) and its support kernels
 * pack_active_to_front / zero_out_kernel / reset_counter_kernel (GRTF:167-190).  Instead of drawing one order
 * per grating hit, a ray that hits a fold-coupler slice SPLITS: the zero order continues in the ray's own
 * row, the diffracted order is appended as a new row at an atomically incremented index, and the launch
 * ends for both (GRTF:247-282, 300-366); in the out-coupler zone every hit deposits the out-coupled energy
 * |E|^2 (a float, not a count) into the eyebox bin and the ray continues with the zero order (GRTF:384-410).
 * The reference ships no host driver and no LUT files for this kernel; the launch-level contract below is
 * what its signature implies, with `vectors` as float64 rows and single-wavelength tables.
 * ====================================================================================================== */
#define WGRT_LEGACY_COLS 13 /* x, y, gap_x, gap_y, theta, phi, m, n, Ete, Etm, delta_phase, region_state, flag */

typedef struct wgrt_legacy_problem {
  double* vectors;            /* [capacity, 13] float64 ray rows, read and written                              */
  int64_t capacity;           /* rows allocated: children beyond it are dropped and reported (the reference
                               * would write out of bounds)                                                   */
  int64_t useful_count_in;    /* rows [0, useful_count_in) are processed (GRTF:202-205)                        */
  int32_t* total_ray_counter; /* [1]: row index of the next appended child (d_total_ray_counter)               */
  int64_t max_steps;          /* MAX_STEPS                                                                    */
  const double* IC;  int64_t IC_n;
  const double* FC;  int64_t FC_n;  const int64_t* FC_offset;  int64_t n_FC;
  const double* OC;  int64_t OC_n;  const int64_t* OC_offset;  int64_t n_OC;
  const double* eff_reg1;  int64_t eff_reg1_n;
  const double* eff_reg2;  int64_t eff_reg2_n;
  const double* eff_reg_FOV;       /* [X, Y, 4, 2] */
  const double* eff_reg_FOV_range; /* [X, Y, 4]    */
  const double* lut_ic1; /* complex128 [X, Y, C_ic]       (channels 8, 11, 20, 23)                             */
  const double* lut_ic2; /* complex128 [X, Y, C_ic]       (channels 0, 1, 3, 6, 15, 18)                        */
  const double* lut_fc1; /* complex128 [n_FC, X, Y, C_fc] (channels 0, 1, 3, 4, 6, 7, 15, 16, 18, 19)          */
  const double* lut_fc2; /* complex128 [n_FC, X, Y, C_fc] (channels 0, 1, 2, 3, 5, 6, 14, 15, 17, 18)          */
  const double* lut_oc;  /* complex128 [n_OC, X, Y, C_oc] (channels 3, 6, 10, 13, 15, 18, 22, 25)              */
  int32_t C_ic, C_fc, C_oc, reserved0; /* >= 24, >= 20, >= 26 */
  const double* lut_TIR; /* [X, Y, 4] */
  const double* lut_gap; /* [X, Y, 8] */
  int64_t X, Y;
  float* matrix_EB;      /* float32 [Y, X, EBy, EBx], accumulated in place (GRTF:154-165: index (n, m, iy, ix)) */
  int64_t EBy, EBx;
} wgrt_legacy_problem_t;

int wgrt_legacy_problem_size(void);

/* One launch of GRTF.process_rays_kernel[grid, block](vectors, useful_count_in, d_total_ray_counter, MAX_STEPS,
 * ...) on DEVICE pointers, asynchronous on `stream`.  Children whose row index would be >= capacity are
 * dropped; *total_ray_counter still counts them, so the caller sees the overflow as counter > capacity. */
int wgrt_legacy_step(const wgrt_legacy_problem_t* dev_problem, void* stream);

/* GRTF.pack_active_to_front[grid, block](src, dst, src_len, out_count) (GRTF:178-190): copies the rows of
 * src[0:src_len] with flag != 0 and Ete^2 + Etm^2 > 0 to the front of dst; *dev_out_count (int32, must be 0
 * before) receives their number.  Row slots are claimed per warp with a ballot / prefix sum and one atomic per
 * warp (the reference: one atomic per row); the order of the packed rows is unspecified in both.  DEVICE
 * pointers, asynchronous. */
int wgrt_legacy_pack_active(const double* dev_src, double* dev_dst, int64_t src_len, int32_t* dev_out_count,
                            void* stream);

/* Whole job on HOST arrays (the driver loop the reference never shipped): upload, then per generation one
 * wgrt_legacy_step over the live rows followed by wgrt_legacy_pack_active into the other buffer, until no row
 * is live or `max_generations` is reached; matrix_EB comes back.  host_problem->vectors holds the initial rows
 * [0, useful_count_in) and receives the rows still live at the end (stats[1] of them).
 * stats (uint64[8], may be NULL): 0 generations run, 1 rows live at the end, 2 rows processed (sum over the
 * generations), 3 children appended, 4 children dropped for lack of capacity, 5 largest live row count. */
int wgrt_legacy_trace_host(const wgrt_legacy_problem_t* host_problem, int max_generations, uint64_t* stats);

#ifdef __cplusplus
}
#endif
#endif /* WGRT_H_ */
