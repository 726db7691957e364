// wgrt_api.cu -- the C ABI of include/wgrt.h: validation, per-device workspace, launches.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "wgrt_device.cuh"

namespace {

using namespace wgrt;

thread_local char g_err[512] = "no error";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t err__ = (expr);                                                             \
    if (err__ != cudaSuccess)                                                               \
      return fail(WGRT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                  __FILE__, __LINE__);                                                      \
  } while (0)

// ---- per-device workspace --------------------------------------------------------------------
struct DeviceBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t want) {
    if (want <= bytes) return cudaSuccess;
    if (ptr) {
      cudaError_t e = cudaFree(ptr);
      ptr = nullptr;
      bytes = 0;
      if (e != cudaSuccess) return e;
    }
    size_t grown = want + want / 8 + 256;
    cudaError_t e = cudaMalloc(&ptr, grown);
    if (e == cudaSuccess) bytes = grown;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
};

struct Workspace {
  int device = -1;
  int num_sms = 0;
  DeviceBuf cells[NUM_REGIONS], detail[NUM_REGIONS], rowmask[NUM_REGIONS], coarse[NUM_REGIONS], rowmask_c[NUM_REGIONS];
  DeviceBuf atlas;   // uint32 [ATLAS_N * ATLAS_N] level 1, then [ATLAS_N2 * ATLAS_N2] level 2
  DeviceBuf zone2;   // uint16 [ATLAS_N2 * ATLAS_N2] level-2 zone ids
  DeviceBuf zone_small;   // level1 | words | trans | hash_keys | hash_zone | words1 | ZoneDyn (see setup_regions)
  DeviceBuf small;   // RegionDyn[NUM_REGIONS] | AtlasDyn @ 256 | tile counters @ 512 | hash state @ 768 | counters @ 1024 | Region[NUM_REGIONS] @ 1280
  bool index_stale = true;  // the index buffers were (re)allocated or used by a debug call
  DeviceBuf jones[3];   // per-warp Jones-matrix scratch of the warp walk, one per launch slot
  DeviceBuf redo[3];    // near-tie redo list of the warp walk, one per launch slot
  DeviceBuf arena;   // staging for the host entry points
  // Launches of one device share this workspace (tile counter, Jones scratch, region index): every
  // enqueue waits for the previous one when it goes to another stream (see order_after_last_use)
  cudaEvent_t last_use = nullptr;
  cudaStream_t last_stream = nullptr;
  bool has_last = false;
  // device polygon-offset arrays already validated: (pointer, entries, vertex count)
  struct SeenOffsets { const void* ptr; int64_t n, nv; };
  std::vector<SeenOffsets> seen_offsets;
  bool pipe_ready = false;                         // streams of the host entry's H2D / walk / D2H pipeline
  cudaStream_t s_in = nullptr, s_run[2] = {nullptr, nullptr}, s_out = nullptr;
  RegionDyn* dyn() { return static_cast<RegionDyn*>(small.ptr); }
  AtlasDyn* atlas_dyn() { return reinterpret_cast<AtlasDyn*>(static_cast<char*>(small.ptr) + 256); }
  // {tile counter, tile size} pairs, one per concurrent launch slot: slot 0 = wgrt_trace_fullcolor,
  // slots 1 and 2 = the two walk streams of the host entry's pipeline
  int* work(int slot) { return reinterpret_cast<int*>(static_cast<char*>(small.ptr) + 512 + 16 * slot); }
  unsigned long long* hash_state() {
    return reinterpret_cast<unsigned long long*>(static_cast<char*>(small.ptr) + 768);
  }
  unsigned long long* counters() {
    return reinterpret_cast<unsigned long long*>(static_cast<char*>(small.ptr) + 1024);
  }
  void release() {
    for (int r = 0; r < NUM_REGIONS; ++r) {
      cells[r].release(); detail[r].release(); rowmask[r].release(); coarse[r].release(); rowmask_c[r].release();
    }
    small.release();
    atlas.release();
    zone2.release();
    zone_small.release();
    for (auto& j : jones) j.release();
    for (auto& j : redo) j.release();
    arena.release();
    if (last_use) cudaEventDestroy(last_use);
    last_use = nullptr;
    has_last = false;
    seen_offsets.clear();
    if (pipe_ready) {
      cudaStreamDestroy(s_in); cudaStreamDestroy(s_run[0]); cudaStreamDestroy(s_run[1]); cudaStreamDestroy(s_out);
      pipe_ready = false;
    }
  }
};

std::mutex g_mu;
std::vector<Workspace> g_ws;

// Region grid geometry: nc x nc coarse cells, each split into 2^shift x 2^shift fine cells.
// WGRT_COARSE / WGRT_GRID override the defaults (64 coarse, 4096 fine cells per axis).
void grid_geometry(int* nc, int* shift) {
  static int g_nc = 0, g_shift = 0;
  if (!g_nc) {
    const char* ec = getenv("WGRT_COARSE");
    const char* ef = getenv("WGRT_GRID");
    int c = ec ? atoi(ec) : 64, f = ef ? atoi(ef) : 4096;
    if (c < 4) c = 4;
    if (c > 256) c = 256;
    int sh = 0;
    while ((c << (sh + 1)) <= f && sh < 8) ++sh;
    g_nc = c;
    g_shift = sh;
  }
  *nc = g_nc;
  *shift = g_shift;
}

int get_workspace(Workspace** out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(WGRT_ERR_NO_DEVICE, "cudaGetDevice: %s", cudaGetErrorString(e));
  for (auto& w : g_ws)
    if (w.device == dev) { *out = &w; return WGRT_OK; }
  g_ws.reserve(64);
  g_ws.emplace_back();
  Workspace& w = g_ws.back();
  w.device = dev;
  CUDA_TRY(cudaDeviceGetAttribute(&w.num_sms, cudaDevAttrMultiProcessorCount, dev));
  CUDA_TRY(w.small.reserve(2048));
  CUDA_TRY(cudaMemset(w.small.ptr, 0, 2048));
  *out = &w;
  return WGRT_OK;
}

int setup_regions(Workspace& w, RegionSet& rs, const double* const verts[NUM_REGIONS],
                  const int64_t* const offsets[NUM_REGIONS], const int64_t nverts[NUM_REGIONS],
                  const int64_t npoly[NUM_REGIONS]) {
  int nc, shift;
  grid_geometry(&nc, &shift);
  for (int r = 0; r < NUM_REGIONS; ++r) {
    RegionStatic& st = rs.st[r];
    if (nverts[r] > (1 << 24)) return fail(WGRT_ERR_UNSUPPORTED, "region %d: too many vertices", r);
    st.verts = verts[r];
    st.offsets = offsets[r];
    st.nverts = static_cast<int>(nverts[r]);
    st.npoly = static_cast<int>(npoly[r]);
    st.nc = nc;
    st.shift = shift;
    st.n = nc << shift;
    st.words = (st.nverts + 31) / 32;
    const size_t fine = static_cast<size_t>(st.n) * st.n, wd = st.words > 0 ? st.words : 1;
    const void* before[5] = {w.coarse[r].ptr, w.cells[r].ptr, w.detail[r].ptr, w.rowmask[r].ptr, w.rowmask_c[r].ptr};
    CUDA_TRY(w.coarse[r].reserve(static_cast<size_t>(nc) * nc));
    CUDA_TRY(w.cells[r].reserve(fine));
    CUDA_TRY(w.detail[r].reserve(fine * sizeof(uint32_t)));
    CUDA_TRY(w.rowmask[r].reserve(static_cast<size_t>(st.n) * wd * sizeof(uint32_t)));
    CUDA_TRY(w.rowmask_c[r].reserve(static_cast<size_t>(nc) * wd * sizeof(uint32_t)));
    st.coarse = static_cast<uint8_t*>(w.coarse[r].ptr);
    st.cells = static_cast<uint8_t*>(w.cells[r].ptr);
    st.detail = static_cast<uint32_t*>(w.detail[r].ptr);
    st.rowmask = static_cast<uint32_t*>(w.rowmask[r].ptr);
    st.rowmask_coarse = static_cast<uint32_t*>(w.rowmask_c[r].ptr);
    const void* after[5] = {w.coarse[r].ptr, w.cells[r].ptr, w.detail[r].ptr, w.rowmask[r].ptr, w.rowmask_c[r].ptr};
    for (int k = 0; k < 5; ++k)
      if (before[k] != after[k]) w.index_stale = true;
  }
  {
    const void* before = w.atlas.ptr;
    CUDA_TRY(w.atlas.reserve(sizeof(uint32_t) * (static_cast<size_t>(ATLAS_N) * ATLAS_N + static_cast<size_t>(ATLAS_N2) * ATLAS_N2)));
    if (before != w.atlas.ptr) w.index_stale = true;
  }
  {
    const void* b2 = w.zone2.ptr;
    const void* bs = w.zone_small.ptr;
    CUDA_TRY(w.zone2.reserve(sizeof(uint16_t) * static_cast<size_t>(ATLAS_N2) * ATLAS_N2));
    const size_t o_words = sizeof(uint16_t) * ZONE_N1 * ZONE_N1, o_trans = o_words + sizeof(uint32_t) * ZONE_CAP,
                 o_keys = o_trans + sizeof(uint32_t) * ZONE_STATES * ZONE_CAP, o_hz = o_keys + sizeof(uint32_t) * 2 * ZONE_CAP,
                 o_w1 = o_hz + sizeof(uint16_t) * 2 * ZONE_CAP, o_dyn = o_w1 + sizeof(uint32_t) * ZONE_N1 * ZONE_N1;
    CUDA_TRY(w.zone_small.reserve(o_dyn + sizeof(ZoneDyn)));
    if (b2 != w.zone2.ptr || bs != w.zone_small.ptr) w.index_stale = true;
    char* base = static_cast<char*>(w.zone_small.ptr);
    rs.zones.level1 = reinterpret_cast<uint16_t*>(base);
    rs.zones.level2 = static_cast<uint16_t*>(w.zone2.ptr);
    rs.zones.words = reinterpret_cast<uint32_t*>(base + o_words);
    rs.zones.trans = reinterpret_cast<uint32_t*>(base + o_trans);
    rs.zones.hash_keys = reinterpret_cast<uint32_t*>(base + o_keys);
    rs.zones.hash_zone = reinterpret_cast<uint16_t*>(base + o_hz);
    rs.zones.words1 = reinterpret_cast<uint32_t*>(base + o_w1);
    rs.zones.dyn = reinterpret_cast<ZoneDyn*>(base + o_dyn);
  }
  rs.atlas = static_cast<uint32_t*>(w.atlas.ptr);
  rs.atlas_dyn = w.atlas_dyn();
  rs.regions = static_cast<char*>(w.small.ptr) + 1280;
  rs.dyn = w.dyn();
  rs.hash_state = w.hash_state();
  rs.dirty = w.hash_state() + 1;
  return WGRT_OK;
}

int validate(const wgrt_problem_t* p, bool need_bins = true) {
  if (!p) return fail(WGRT_ERR_INVALID, "null problem");
  if (p->num_rays < 0) return fail(WGRT_ERR_INVALID, "num_rays < 0");
  if (!(p->threshold >= 0.0)) return fail(WGRT_ERR_INVALID, "threshold must be >= 0");
  if (p->L <= 0 || p->X <= 0 || p->Y <= 0 || p->EBx <= 0 || p->EBy <= 0)
    return fail(WGRT_ERR_INVALID, "L, X, Y, EBy, EBx must be positive");
  if (p->n_FC < 0 || p->n_OC < 0 || p->n_FC > 250 || p->n_OC > 250)
    return fail(WGRT_ERR_INVALID, "n_FC / n_OC must be within [0, 250]");
  if (p->C_ic < 41 || p->C_fc < 20 || p->C_oc < 41)
    return fail(WGRT_ERR_INVALID, "LUT channel counts too small (need C_ic>=41, C_fc>=20, C_oc>=41)");
  if (p->IC_n < 0 || p->FC_n < 0 || p->OC_n < 0 || p->eff_reg1_n < 0 || p->eff_reg2_n < 0)
    return fail(WGRT_ERR_INVALID, "negative vertex count");
#define NEED(f) \
  if (!p->f) return fail(WGRT_ERR_INVALID, "null pointer: " #f)
  if (p->runner_points < 0 || p->runner_first_cell < 0) return fail(WGRT_ERR_INVALID, "negative runner layout field");
  if (p->runner_points > 0) {
    if (p->num_rays % (2 * p->runner_points) != 0)
      return fail(WGRT_ERR_INVALID, "runner layout: num_rays must be a multiple of 2 * runner_points");
    if (p->runner_first_cell + p->num_rays / (2 * p->runner_points) > p->L * p->X * p->Y)
      return fail(WGRT_ERR_INVALID, "runner layout: cell range exceeds L * X * Y");
    if (p->num_rays > 0) { NEED(x); NEED(y); }
  } else if (p->num_rays > 0) {
    NEED(x); NEED(y); NEED(m); NEED(n); NEED(te); NEED(tm); NEED(delta_phase);
    if (!p->lmd_num && p->L != 1) return fail(WGRT_ERR_INVALID, "lmd_num may be NULL only when L == 1");
  }
  NEED(IC); NEED(FC); NEED(FC_offset); NEED(OC); NEED(OC_offset); NEED(eff_reg1); NEED(eff_reg2);
  NEED(eff_reg_FOV); NEED(eff_reg_FOV_range); NEED(lut_ic1); NEED(lut_ic2); NEED(lut_ic3); NEED(lut_fc1);
  NEED(lut_fc2); NEED(lut_oc1); NEED(lut_oc2); NEED(lut_TIR); NEED(lut_gap);
  if (need_bins) NEED(matrix_EB);
#undef NEED
  return WGRT_OK;
}

int region_set_of(Workspace& w, const wgrt_problem_t& p, RegionSet& rs) {
  const double* verts[NUM_REGIONS] = {p.IC, p.eff_reg1, p.eff_reg2, p.FC, p.OC};
  const int64_t* offs[NUM_REGIONS] = {nullptr, nullptr, nullptr, p.FC_offset, p.OC_offset};
  const int64_t nv[NUM_REGIONS] = {p.IC_n, p.eff_reg1_n, p.eff_reg2_n, p.FC_n, p.OC_n};
  const int64_t np[NUM_REGIONS] = {1, 1, 1, p.n_FC, p.n_OC};
  return setup_regions(w, rs, verts, offs, nv, np);
}

// (Re)build the region index for p's polygons on `stream` if their content changed.
int build_region_index(Workspace& w, const wgrt_problem_t& p, cudaStream_t stream) {
  RegionSet rs;
  int rc = region_set_of(w, p, rs);
  if (rc != WGRT_OK) return rc;
  CUDA_TRY(launch_region_build(rs, w.index_stale, stream));
  w.index_stale = false;
  return WGRT_OK;
}

// The workspace is shared by every launch of the device.  Work enqueued on `stream` that touches it must
// run after the previous enqueue's work when that went to a different stream (same-stream work is
// ordered already).  Launches of one device therefore execute in enqueue order, whatever their streams.
int order_after_last_use(Workspace& w, cudaStream_t stream) {
  if (w.has_last && w.last_stream != stream) CUDA_TRY(cudaStreamWaitEvent(stream, w.last_use, 0));
  return WGRT_OK;
}
int record_last_use(Workspace& w, cudaStream_t stream) {
  if (!w.last_use) CUDA_TRY(cudaEventCreateWithFlags(&w.last_use, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(w.last_use, stream));
  w.last_stream = stream;
  w.has_last = true;
  return WGRT_OK;
}

int check_offsets(const int64_t* off, int64_t np, int64_t nv, const char* name) {
  if (off[0] != 0 || off[np] > nv) return fail(WGRT_ERR_INVALID, "%s: must start at 0 and end within the vertex array", name);
  for (int64_t k = 0; k < np; ++k)
    if (off[k + 1] < off[k]) return fail(WGRT_ERR_INVALID, "%s: must be non-decreasing", name);
  return WGRT_OK;
}

// Device path: the polygon offsets index the vertex arrays inside the kernels, so malformed offsets would
// read out of bounds.  Each (pointer, size) pair is copied to the host and checked the first time it is
// seen on a device (at most 251 entries; one small synchronous copy on the caller's stream).
int validate_device_offsets(Workspace& w, const wgrt_problem_t& p, cudaStream_t stream) {
  for (int s = 0; s < 2; ++s) {
    const int64_t* off = s ? p.OC_offset : p.FC_offset;
    const int64_t np = s ? p.n_OC : p.n_FC, nv = s ? p.OC_n : p.FC_n;
    bool seen = false;
    for (auto& e : w.seen_offsets) seen = seen || (e.ptr == off && e.n == np && e.nv == nv);
    if (seen) continue;
    int64_t host[256];
    CUDA_TRY(cudaMemcpyAsync(host, off, sizeof(int64_t) * (np + 1), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    int rc = check_offsets(host, np, nv, s ? "OC_offset" : "FC_offset");
    if (rc != WGRT_OK) return rc;
    if (w.seen_offsets.size() >= 64) w.seen_offsets.clear();
    w.seen_offsets.push_back({off, np, nv});
  }
  return WGRT_OK;
}

// One launch of the walk.  `slot` selects the tile-counter pair (launches that may run concurrently
// need different slots); `build_index = false` when the caller built the region index already.
int trace_device(Workspace& w, const wgrt_problem_t& p, cudaStream_t stream, int slot = 0, bool build_index = true) {
  if (p.num_rays == 0) return WGRT_OK;
  if (!p.rng_states) return fail(WGRT_ERR_INVALID, "null pointer: rng_states");
  if (p.flags & WGRT_FLAG_STRICT) {
    CUDA_TRY(launch_walk_strict(p, w.counters(), stream));
    return WGRT_OK;
  }
  RegionSet rs;
  int rc = region_set_of(w, p, rs);
  if (rc != WGRT_OK) return rc;
  if (build_index) {
    CUDA_TRY(launch_region_build(rs, w.index_stale, stream));
    w.index_stale = false;
  }
  // (a reallocation frees the old scratch with cudaFree, which waits for work still using it)
  CUDA_TRY(w.jones[slot].reserve(walk_warp_scratch_bytes(p, w.num_sms)));
  CUDA_TRY(w.redo[slot].reserve(sizeof(RedoList)));
  CUDA_TRY(launch_walk_warp(p, rs, w.work(slot), w.counters(), w.num_sms, static_cast<double*>(w.jones[slot].ptr),
                            static_cast<RedoList*>(w.redo[slot].ptr), stream));
  return WGRT_OK;
}

// bump allocator over the workspace arena
struct Arena {
  char* base;
  size_t cap, used = 0;
  void* take(size_t bytes) {
    size_t at = (used + 255) & ~size_t(255);
    used = at + bytes;
    return used <= cap ? base + at : nullptr;
  }
};
size_t padded(size_t b) { return ((b + 255) & ~size_t(255)) + 256; }

}  // namespace

extern "C" {

int wgrt_version(void) { return WGRT_VERSION; }

int wgrt_problem_size(void) { return static_cast<int>(sizeof(wgrt_problem_t)); }

const char* wgrt_last_error(void) { return g_err; }

int wgrt_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(WGRT_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}

int wgrt_release(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(WGRT_ERR_NO_DEVICE, "no CUDA device");
  for (auto& w : g_ws)
    if (w.device == dev) {
      cudaDeviceSynchronize();
      w.release();
      w.device = -1;
    }
  return WGRT_OK;
}

int wgrt_trace_fullcolor(const wgrt_problem_t* p, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  int rc = validate(p);
  if (rc != WGRT_OK) return rc;
  Workspace* w = nullptr;
  rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->num_rays == 0) return WGRT_OK;
  rc = validate_device_offsets(*w, *p, st);
  if (rc != WGRT_OK) return rc;
  rc = order_after_last_use(*w, st);
  if (rc != WGRT_OK) return rc;
  rc = trace_device(*w, *p, st);
  if (rc != WGRT_OK) return rc;
  return record_last_use(*w, st);
}

}  // extern "C" (the host entry below needs helpers from an unnamed namespace)

// The host entry is a three-stage pipeline over CHUNKS of the job (H2D stream -> walk stream -> D2H
// stream, chained by events), so that the PCIe transfers of one chunk run under the walk of its
// neighbours.  Rays are independent and carry their own RNG stream (GRTF:25-34, RUN:158), so walking
// the job chunk by chunk -- and a chunk num_iter times before the next one starts -- is bit-identical
// to num_iter launches over the whole ray set.
//   runner layout : a chunk is a range of FoV-x columns (the runner's outermost loop, RUN:82-84).  The
//                   LUT / TIR / gap / eyebox tables of those columns go up as strided 2-D copies into
//                   the full-shape device arrays, the chunk's rays are walked, and the columns' slice
//                   matrix_EB[:, :, m0:m1] comes down while the next chunk is walked.
//   ray arrays    : a chunk is a range of rays; the shared tables go up first, the bins come down last.
namespace {

struct Copy2D {
  void* dst; const void* src; size_t dpitch, spitch, width, height;
};

cudaError_t copy2d(const Copy2D& c, cudaMemcpyKind kind, cudaStream_t st) {
  if (!c.width || !c.height) return cudaSuccess;
  if (c.height == 1 || (c.width == c.dpitch && c.width == c.spitch))
    return cudaMemcpyAsync(c.dst, c.src, c.width * c.height, kind, st);
  return cudaMemcpy2DAsync(c.dst, c.dpitch, c.src, c.spitch, c.width, c.height, kind, st);
}

int host_chunk_target(int64_t num_rays, int num_iter) {
  const char* e = getenv("WGRT_HOST_CHUNKS");   // read at every call: tests force odd chunkings
  const int forced = e ? atoi(e) : 0;
  if (forced > 0) return forced;
  // enough chunks to hide the transfers of the first and last one, few enough that a chunk still
  // fills the GPU: measured on C2 (B200, warp walk) for num_iter = 1: 8 chunks 19.3 ms, 16: 18.7,
  // 25: 18.2; for num_iter = 4: 4 chunks 41.2 ms, 8: 37.9, 12: 37.4, 16: 38.1, 25: 43.1
  // The walk kernel fills every SM with one CTA, so the launches of the two walk streams do not overlap and every
  // chunk launch ends in a tail of half-empty SMs: a long job, whose time is the walk's, wants few, large chunks
  // (num_iter = 10: 3 chunks 91 ms, 6: 94, 12: 99; num_iter = 4: 3 to 12 chunks 41 - 42 ms; num_iter = 2: 12: 23.0).
  const int64_t by_size = num_rays / 4500000;
  const int64_t cap = num_iter >= 8 ? 3 : num_iter >= 3 ? 6 : num_iter == 2 ? 12 : 25;
  return static_cast<int>(by_size < 1 ? 1 : (by_size > cap ? cap : by_size));
}

// Work tile of a chunk launch: chosen by the walk itself from the launch size (launch_walk_warp);
// WGRT_HOST_TILE=<rays> overrides it for experiments.
uint32_t host_chunk_tile() {
  const char* e = getenv("WGRT_HOST_TILE");
  return static_cast<uint32_t>(e && atoi(e) > 0 ? atoi(e) : 0);
}

// Optional evaluation stage of the host entry (wgrt_trace_evaluate_host): the bins stay on the device,
// the pupil sums / per-cell totals (row f1) are reduced there and only those come down.
struct EvalSpec {
  int mask_size, step_y, step_x;
  float* perceive;    // host [L, Y, X, n_epy, n_epx] or NULL
  float* cell_sums;   // host [L, Y, X] or NULL
  const wgrt_eval_params_t* params = nullptr;   // not NULL: finish evaluation() on the device (wgrt_eval_metrics)
  double* metrics = nullptr;                    // host [n_ep, WGRT_EVAL_NUM]
  float* image = nullptr;                       // host [Y, X, 3, n_epy, n_epx] or NULL
};

struct HostChunk {
  int64_t ray0 = 0, rays = 0;        // launch-relative ray range walked by this chunk
  int64_t cell0 = 0;                 // runner layout: first cell of the chunk (global cell index)
  int64_t in_m0 = 0, in_m1 = 0;      // runner layout: FoV-x columns whose tables this chunk uploads
  int64_t out_m0 = 0, out_m1 = 0;    // runner layout: FoV-x columns of matrix_EB this chunk moves
};

}  // namespace

namespace {
int trace_host_impl(const wgrt_problem_t* hp, int num_iter, float* timings_ms, const EvalSpec* ev);
}

extern "C" int wgrt_trace_fullcolor_host(const wgrt_problem_t* hp, int num_iter, float* timings_ms) {
  return trace_host_impl(hp, num_iter, timings_ms, nullptr);
}

extern "C" int wgrt_trace_evaluate_host(const wgrt_problem_t* hp, int num_iter, int mask_size, int step_y, int step_x,
                                        float* perceive, float* cell_sums, float* timings_ms) {
  if (mask_size <= 0 || step_y <= 0 || step_x <= 0) return fail(WGRT_ERR_INVALID, "bad pupil mask / steps");
  if (mask_size > 220) return fail(WGRT_ERR_UNSUPPORTED, "pupil mask wider than 220 bins");
  const EvalSpec ev{mask_size, step_y, step_x, perceive, cell_sums};
  return trace_host_impl(hp, num_iter, timings_ms, &ev);
}

extern "C" int wgrt_trace_evaluate_metrics_host(const wgrt_problem_t* hp, int num_iter, int mask_size, int step_y,
                                                int step_x, const wgrt_eval_params_t* params, double* metrics,
                                                float* cell_sums, float* perceive, float* image, float* timings_ms) {
  if (mask_size <= 0 || step_y <= 0 || step_x <= 0) return fail(WGRT_ERR_INVALID, "bad pupil mask / steps");
  if (mask_size > 220) return fail(WGRT_ERR_UNSUPPORTED, "pupil mask wider than 220 bins");
  if (!params || !metrics) return fail(WGRT_ERR_INVALID, "null pointer: params / metrics");
  if (hp && hp->L != 3) return fail(WGRT_ERR_UNSUPPORTED, "the colour evaluation needs exactly 3 wavelengths");
  EvalSpec ev{mask_size, step_y, step_x, perceive, cell_sums};
  ev.params = params; ev.metrics = metrics; ev.image = image;
  return trace_host_impl(hp, num_iter, timings_ms, &ev);
}

namespace {
int trace_host_impl(const wgrt_problem_t* hp, int num_iter, float* timings_ms, const EvalSpec* ev) {
  std::lock_guard<std::mutex> lk(g_mu);
  int rc = validate(hp, ev == nullptr);
  if (rc != WGRT_OK) return rc;
  if (num_iter < 0) return fail(WGRT_ERR_INVALID, "num_iter < 0");
  rc = check_offsets(hp->FC_offset, hp->n_FC, hp->FC_n, "FC_offset");
  if (rc != WGRT_OK) return rc;
  rc = check_offsets(hp->OC_offset, hp->n_OC, hp->OC_n, "OC_offset");
  if (rc != WGRT_OK) return rc;
  Workspace* w = nullptr;
  rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;

  const size_t N = static_cast<size_t>(hp->num_rays);
  const size_t L = hp->L, X = hp->X, Y = hp->Y;
  const size_t cells = L * X * Y, fov = X * Y;
  const size_t tile_b = static_cast<size_t>(hp->EBy * hp->EBx) * 4;   // one FoV cell's bins
  const size_t eb_b = cells * tile_b;
  const bool runner = hp->runner_points > 0;
  const bool seed_rng = runner && hp->rng_states == nullptr;
  // with an evaluation stage and no host bin array the bins live and die on the device
  const bool bins_on_device = (hp->flags & WGRT_FLAG_BINS_DEVICE) != 0;   // caller's device tensor, used in place
  if (bins_on_device && !hp->matrix_EB) return fail(WGRT_ERR_INVALID, "WGRT_FLAG_BINS_DEVICE needs matrix_EB");
  const bool bins_to_host = hp->matrix_EB != nullptr && !bins_on_device;
  const bool zero_bins = (hp->flags & WGRT_FLAG_BINS_ZERO) != 0 || (!bins_to_host && !bins_on_device);
  if (!runner && N && !hp->rng_states) return fail(WGRT_ERR_INVALID, "null pointer: rng_states");

  // ---- device staging: full-shape copies of every array (arena, grow-only) --------------------
  struct Item { const void* src; void** dst; size_t bytes; bool upfront; };
  wgrt_problem_t dp = *hp;
  dp.gap_x = dp.gap_y = dp.pol = dp.azi = nullptr;  // never read by the walk
  dp.flags &= ~(WGRT_FLAG_BINS_ZERO | WGRT_FLAG_BINS_DEVICE | WGRT_FLAG_BINS_COLUMNS);
  const bool columns_only = (hp->flags & WGRT_FLAG_BINS_COLUMNS) != 0;
  if (columns_only && !runner) return fail(WGRT_ERR_INVALID, "WGRT_FLAG_BINS_COLUMNS needs the runner layout");
  const size_t ray_b = N * 4, pts_b = static_cast<size_t>(hp->runner_points) * 4;
  const size_t ic_b = fov * hp->C_ic * 16, fc_b = fov * hp->C_fc * 16, oc_b = fov * hp->C_oc * 16;  // per wavelength (and slice)
  if (runner) dp.m = dp.n = dp.lmd_num = dp.te = dp.tm = dp.delta_phase = nullptr;
  std::vector<Item> items =
      runner ? std::vector<Item>{{hp->x, (void**)&dp.x, pts_b, true}, {hp->y, (void**)&dp.y, pts_b, true}}
             : std::vector<Item>{{hp->x, (void**)&dp.x, ray_b, false}, {hp->y, (void**)&dp.y, ray_b, false},
                                 {hp->m, (void**)&dp.m, ray_b, false}, {hp->n, (void**)&dp.n, ray_b, false},
                                 {hp->lmd_num, (void**)&dp.lmd_num, hp->lmd_num ? ray_b : 0, false},
                                 {hp->te, (void**)&dp.te, ray_b, false}, {hp->tm, (void**)&dp.tm, ray_b, false},
                                 {hp->delta_phase, (void**)&dp.delta_phase, ray_b, false}};
  const size_t n_ray_items = items.size();
  items.push_back({nullptr, (void**)&dp.rng_states, ray_b, false});
  const bool tables_upfront = !runner;   // runner layout: the per-cell tables go up column range by column range
  const std::vector<Item> shared_items = {
      {hp->IC, (void**)&dp.IC, (size_t)hp->IC_n * 16, true}, {hp->FC, (void**)&dp.FC, (size_t)hp->FC_n * 16, true},
      {hp->FC_offset, (void**)&dp.FC_offset, (size_t)(hp->n_FC + 1) * 8, true},
      {hp->OC, (void**)&dp.OC, (size_t)hp->OC_n * 16, true},
      {hp->OC_offset, (void**)&dp.OC_offset, (size_t)(hp->n_OC + 1) * 8, true},
      {hp->eff_reg1, (void**)&dp.eff_reg1, (size_t)hp->eff_reg1_n * 16, true},
      {hp->eff_reg2, (void**)&dp.eff_reg2, (size_t)hp->eff_reg2_n * 16, true},
      {hp->eff_reg_FOV, (void**)&dp.eff_reg_FOV, fov * 64, tables_upfront},
      {hp->eff_reg_FOV_range, (void**)&dp.eff_reg_FOV_range, fov * 32, tables_upfront},
      {hp->lut_ic1, (void**)&dp.lut_ic1, L * ic_b, tables_upfront}, {hp->lut_ic2, (void**)&dp.lut_ic2, L * ic_b, tables_upfront},
      {hp->lut_ic3, (void**)&dp.lut_ic3, L * ic_b, tables_upfront},
      {hp->lut_fc1, (void**)&dp.lut_fc1, L * hp->n_FC * fc_b, tables_upfront},
      {hp->lut_fc2, (void**)&dp.lut_fc2, L * hp->n_FC * fc_b, tables_upfront},
      {hp->lut_oc1, (void**)&dp.lut_oc1, L * hp->n_OC * oc_b, tables_upfront},
      {hp->lut_oc2, (void**)&dp.lut_oc2, L * hp->n_OC * oc_b, tables_upfront},
      {hp->lut_TIR, (void**)&dp.lut_TIR, cells * 32, tables_upfront}, {hp->lut_gap, (void**)&dp.lut_gap, cells * 64, tables_upfront},
      {bins_on_device ? nullptr : hp->matrix_EB, (void**)&dp.matrix_EB, bins_on_device ? 0 : eb_b, !zero_bins && !runner},
  };
  items.insert(items.end(), shared_items.begin(), shared_items.end());
  float *d_perceive = nullptr, *d_cells = nullptr, *d_image = nullptr;
  double* d_metrics = nullptr;
  size_t perceive_b = 0, metrics_b = 0, image_b = 0, n_epy = 0, n_epx = 0;
  if (ev) {
    n_epy = hp->EBy >= ev->mask_size ? static_cast<size_t>((hp->EBy - ev->mask_size) / ev->step_y + 1) : 0;
    n_epx = hp->EBx >= ev->mask_size ? static_cast<size_t>((hp->EBx - ev->mask_size) / ev->step_x + 1) : 0;
    perceive_b = cells * n_epy * n_epx * 4;
    items.push_back({nullptr, (void**)&d_perceive, perceive_b, false});
    items.push_back({nullptr, (void**)&d_cells, cells * 4, false});
    if (ev->params) {
      metrics_b = n_epy * n_epx * WGRT_EVAL_NUM * sizeof(double);
      image_b = ev->image ? fov * 3 * n_epy * n_epx * 4 : 0;
      items.push_back({nullptr, (void**)&d_metrics, metrics_b, false});
      items.push_back({nullptr, (void**)&d_image, image_b, false});
    }
  }
  size_t total = 0;
  for (auto& it : items) total += padded(it.bytes);
  CUDA_TRY(w->arena.reserve(total));
  Arena ar{static_cast<char*>(w->arena.ptr), w->arena.bytes};
  for (auto& it : items) {
    *it.dst = (it.bytes || it.src) ? ar.take(it.bytes) : nullptr;
    if (!*it.dst && (it.bytes || it.src)) return fail(WGRT_ERR_CUDA, "arena overflow");
  }
  if (bins_on_device) dp.matrix_EB = hp->matrix_EB;

  // ---- chunk plan ------------------------------------------------------------------------------
  std::vector<HostChunk> chunks;
  const int want = host_chunk_target(hp->num_rays, num_iter);
  if (runner) {
    const int64_t rpc = 2 * hp->runner_points, cpc = static_cast<int64_t>(Y * L);   // cells per FoV-x column
    const int64_t c0 = hp->runner_first_cell, c1 = c0 + hp->num_rays / rpc;
    const int64_t m_lo = N ? c0 / cpc : 0, m_hi = N ? (c1 - 1) / cpc + 1 : 0;
    const int64_t K = N ? std::max<int64_t>(1, std::min<int64_t>(want, m_hi - m_lo)) : 1;
    for (int64_t k = 0; k < K; ++k) {
      HostChunk c;
      c.in_m0 = m_lo + (m_hi - m_lo) * k / K;
      c.in_m1 = m_lo + (m_hi - m_lo) * (k + 1) / K;
      const int64_t a = std::max(c0, c.in_m0 * cpc), b = std::min(c1, c.in_m1 * cpc);
      c.cell0 = a;
      c.ray0 = (a - c0) * rpc;
      c.rays = b > a ? (b - a) * rpc : 0;
      // matrix_EB columns outside the cell range travel with the first / last chunk: the whole
      // tensor is uploaded (unless declared zero) and downloaded, as by a single full copy
      // (WGRT_FLAG_BINS_COLUMNS: only the columns the cell range touches move at all)
      c.out_m0 = (k == 0 && !columns_only) ? 0 : c.in_m0;
      c.out_m1 = (k == K - 1 && !columns_only) ? static_cast<int64_t>(X) : c.in_m1;
      chunks.push_back(c);
    }
  } else {
    const char* forced = getenv("WGRT_HOST_CHUNKS");
    const int64_t grain = (forced && atoi(forced) > 0) ? 64 : 1048576;   // rays per chunk, at least
    const int64_t K = std::max<int64_t>(1, std::min<int64_t>(want, (hp->num_rays + grain - 1) / grain));
    for (int64_t k = 0; k < K; ++k) {
      HostChunk c;
      c.ray0 = hp->num_rays * k / K;
      c.rays = hp->num_rays * (k + 1) / K - c.ray0;
      chunks.push_back(c);
    }
  }
  const size_t K = chunks.size();

  // ---- streams and events ----------------------------------------------------------------------
  if (!w->pipe_ready) {
    CUDA_TRY(cudaStreamCreateWithFlags(&w->s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&w->s_run[0], cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&w->s_run[1], cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&w->s_out, cudaStreamNonBlocking));
    w->pipe_ready = true;
  }
  // two walk streams, chunks alternate: the next chunk's CTAs fill the SMs the previous chunk's
  // persistent CTAs leave as they run out of tiles (no idle tail at chunk boundaries)
  cudaStream_t s_in = w->s_in, s_run2[2] = {w->s_run[0], w->s_run[1]}, s_out = w->s_out;
  std::vector<cudaEvent_t> ev_in(K), ev_run(K);
  cudaEvent_t span[8];   // begin / end of the work on each of the three stages; [6], [7]: shared tables up, index built
  struct EventGuard {
    std::vector<cudaEvent_t>*a, *b; cudaEvent_t* c;
    ~EventGuard() {
      for (auto e : *a) if (e) cudaEventDestroy(e);
      for (auto e : *b) if (e) cudaEventDestroy(e);
      for (int i = 0; i < 8; ++i) if (c[i]) cudaEventDestroy(c[i]);
    }
  } guard{&ev_in, &ev_run, span};
  for (auto& e : span) e = nullptr;
  for (auto& e : ev_in) e = nullptr;
  for (auto& e : ev_run) e = nullptr;
  for (auto& e : span) CUDA_TRY(cudaEventCreate(&e));
  for (auto& e : ev_in) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : ev_run) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  // a failure below must not leave copies running into buffers the caller is about to free
  struct SyncGuard {
    cudaStream_t a, b, c, d;
    ~SyncGuard() { cudaStreamSynchronize(a); cudaStreamSynchronize(b); cudaStreamSynchronize(c); cudaStreamSynchronize(d); }
  } sync_guard{s_in, s_run2[0], s_run2[1], s_out};

  const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;
  // an asynchronous wgrt_trace_fullcolor launch may still be using the workspace (region index)
  rc = order_after_last_use(*w, s_in);
  if (rc != WGRT_OK) return rc;
  rc = order_after_last_use(*w, s_run2[0]);
  if (rc != WGRT_OK) return rc;
  CUDA_TRY(cudaEventRecord(span[0], s_in));
  for (auto& it : items)
    if (it.upfront && it.bytes && it.src) CUDA_TRY(cudaMemcpyAsync(*it.dst, it.src, it.bytes, H2D, s_in));
  if (zero_bins && columns_only && !chunks.empty() && eb_b) {
    const size_t m0 = static_cast<size_t>(chunks.front().out_m0), m1 = static_cast<size_t>(chunks.back().out_m1);
    if (m1 > m0)
      CUDA_TRY(cudaMemset2DAsync((char*)dp.matrix_EB + m0 * tile_b, X * tile_b, 0, (m1 - m0) * tile_b, L * Y, s_in));
  } else if (zero_bins) {
    CUDA_TRY(cudaMemsetAsync(dp.matrix_EB, 0, eb_b, s_in));
  }
  // the region index is built once, on walk stream 0, as soon as the polygons are up
  CUDA_TRY(cudaEventRecord(span[6], s_in));
  CUDA_TRY(cudaStreamWaitEvent(s_run2[0], span[6], 0));
  CUDA_TRY(cudaEventRecord(span[2], s_run2[0]));
  if (N && !(dp.flags & WGRT_FLAG_STRICT)) {
    rc = build_region_index(*w, dp, s_run2[0]);
    if (rc != WGRT_OK) return rc;
  }
  CUDA_TRY(cudaEventRecord(span[7], s_run2[0]));
  CUDA_TRY(cudaStreamWaitEvent(s_run2[1], span[7], 0));

  for (size_t k = 0; k < K; ++k) {
    const HostChunk& c = chunks[k];
    cudaStream_t s_run = s_run2[k & 1];
    // ---- stage 1: this chunk's inputs ----------------------------------------------------------
    if (runner) {
      const size_t cols = static_cast<size_t>(c.in_m1 - c.in_m0), m0 = static_cast<size_t>(c.in_m0);
      auto table = [&](const double* dst, const double* src, size_t row_b, size_t planes) {
        // [planes, X, row] arrays: columns m0 .. m0+cols of every plane
        return copy2d({(char*)dst + m0 * row_b, (const char*)src + m0 * row_b, X * row_b, X * row_b, cols * row_b, planes},
                      H2D, s_in);
      };
      CUDA_TRY(table(dp.lut_ic1, hp->lut_ic1, Y * hp->C_ic * 16, L));
      CUDA_TRY(table(dp.lut_ic2, hp->lut_ic2, Y * hp->C_ic * 16, L));
      CUDA_TRY(table(dp.lut_ic3, hp->lut_ic3, Y * hp->C_ic * 16, L));
      CUDA_TRY(table(dp.lut_fc1, hp->lut_fc1, Y * hp->C_fc * 16, L * hp->n_FC));
      CUDA_TRY(table(dp.lut_fc2, hp->lut_fc2, Y * hp->C_fc * 16, L * hp->n_FC));
      CUDA_TRY(table(dp.lut_oc1, hp->lut_oc1, Y * hp->C_oc * 16, L * hp->n_OC));
      CUDA_TRY(table(dp.lut_oc2, hp->lut_oc2, Y * hp->C_oc * 16, L * hp->n_OC));
      CUDA_TRY(table(dp.lut_TIR, hp->lut_TIR, Y * 32, L));
      CUDA_TRY(table(dp.lut_gap, hp->lut_gap, Y * 64, L));
      CUDA_TRY(table(dp.eff_reg_FOV, hp->eff_reg_FOV, Y * 64, 1));
      CUDA_TRY(table(dp.eff_reg_FOV_range, hp->eff_reg_FOV_range, Y * 32, 1));
      if (!zero_bins && !bins_on_device) {
        const size_t o0 = static_cast<size_t>(c.out_m0), ocols = static_cast<size_t>(c.out_m1 - c.out_m0);
        CUDA_TRY(copy2d({(char*)dp.matrix_EB + o0 * tile_b, (const char*)hp->matrix_EB + o0 * tile_b, X * tile_b,
                         X * tile_b, ocols * tile_b, L * Y}, H2D, s_in));
      }
    } else {
      for (size_t i = 0; i < n_ray_items; ++i)
        if (items[i].bytes && items[i].src)
          CUDA_TRY(cudaMemcpyAsync((char*)*items[i].dst + c.ray0 * 4, (const char*)items[i].src + c.ray0 * 4,
                                   static_cast<size_t>(c.rays) * 4, H2D, s_in));
    }
    if (!seed_rng && c.rays)
      CUDA_TRY(cudaMemcpyAsync(dp.rng_states + c.ray0, hp->rng_states + c.ray0, static_cast<size_t>(c.rays) * 4, H2D, s_in));
    if (k == K - 1) CUDA_TRY(cudaEventRecord(span[1], s_in));
    CUDA_TRY(cudaEventRecord(ev_in[k], s_in));

    // ---- stage 2: walk the chunk num_iter times --------------------------------------------------
    CUDA_TRY(cudaStreamWaitEvent(s_run, ev_in[k], 0));
    if (c.rays) {
      wgrt_problem_t cp = dp;
      cp.num_rays = c.rays;
      cp.rng_states = dp.rng_states + c.ray0;
      cp.ray_index_base = hp->ray_index_base + c.ray0;
      if (!cp.tile_hint && K > 1) cp.tile_hint = host_chunk_tile();
      if (runner) {
        cp.runner_first_cell = c.cell0;
      } else {
        cp.x += c.ray0; cp.y += c.ray0; cp.m += c.ray0; cp.n += c.ray0; cp.te += c.ray0; cp.tm += c.ray0;
        cp.delta_phase += c.ray0;
        if (cp.lmd_num) cp.lmd_num += c.ray0;
      }
      if (seed_rng)
        CUDA_TRY(launch_seed_rng(cp.rng_states, c.rays, c.cell0 * 2 * hp->runner_points + hp->rng_seed_offset, s_run));
      for (int it = 0; it < num_iter; ++it) {
        rc = trace_device(*w, cp, s_run, 1 + static_cast<int>(k & 1), false);
        if (rc != WGRT_OK) return rc;
      }
    }
    CUDA_TRY(cudaEventRecord(ev_run[k], s_run));

    // ---- stage 3: results of the chunk -----------------------------------------------------------
    CUDA_TRY(cudaStreamWaitEvent(s_out, ev_run[k], 0));
    if (k == 0) CUDA_TRY(cudaEventRecord(span[4], s_out));
    if (c.rays && hp->rng_states)
      CUDA_TRY(cudaMemcpyAsync(hp->rng_states + c.ray0, dp.rng_states + c.ray0, static_cast<size_t>(c.rays) * 4, D2H, s_out));
    if (runner && bins_to_host) {
      const size_t o0 = static_cast<size_t>(c.out_m0), ocols = static_cast<size_t>(c.out_m1 - c.out_m0);
      CUDA_TRY(copy2d({(char*)hp->matrix_EB + o0 * tile_b, (const char*)dp.matrix_EB + o0 * tile_b, X * tile_b, X * tile_b,
                       ocols * tile_b, L * Y}, D2H, s_out));
    } else if (k == K - 1 && bins_to_host) {
      CUDA_TRY(cudaMemcpyAsync(hp->matrix_EB, dp.matrix_EB, eb_b, D2H, s_out));
    }
  }
  if (ev && cells) {
    // the D2H stream has waited for every chunk's walk: reduce the finished bins where they are
    CUDA_TRY(launch_pupil_sums(dp.matrix_EB, hp->L, hp->Y, hp->X, hp->EBy, hp->EBx, ev->mask_size, ev->step_y, ev->step_x,
                               d_perceive, d_cells, s_out));
    if (ev->params && metrics_b) {
      CUDA_TRY(launch_eval_metrics(d_perceive, hp->Y, hp->X, static_cast<int>(n_epy), static_cast<int>(n_epx), *ev->params,
                                   d_metrics, image_b ? d_image : nullptr, s_out));
      CUDA_TRY(cudaMemcpyAsync(ev->metrics, d_metrics, metrics_b, D2H, s_out));
      if (image_b) CUDA_TRY(cudaMemcpyAsync(ev->image, d_image, image_b, D2H, s_out));
    }
    if (ev->perceive && perceive_b) CUDA_TRY(cudaMemcpyAsync(ev->perceive, d_perceive, perceive_b, D2H, s_out));
    if (ev->cell_sums) CUDA_TRY(cudaMemcpyAsync(ev->cell_sums, d_cells, cells * 4, D2H, s_out));
  }
  CUDA_TRY(cudaEventRecord(span[5], s_out));
  // end of the walk stage: the later of the two walk streams (the D2H stream waited for both)
  CUDA_TRY(cudaStreamWaitEvent(s_run2[0], ev_run[K - 1], 0));
  if (K > 1) CUDA_TRY(cudaStreamWaitEvent(s_run2[0], ev_run[K - 2], 0));
  CUDA_TRY(cudaEventRecord(span[3], s_run2[0]));
  CUDA_TRY(cudaStreamSynchronize(s_in));
  CUDA_TRY(cudaStreamSynchronize(s_run2[0]));
  CUDA_TRY(cudaStreamSynchronize(s_run2[1]));
  CUDA_TRY(cudaStreamSynchronize(s_out));
  if (timings_ms)
    for (int k = 0; k < 3; ++k) CUDA_TRY(cudaEventElapsedTime(&timings_ms[k], span[2 * k], span[2 * k + 1]));
  return WGRT_OK;
}
}  // namespace

extern "C" {

int wgrt_seed_rng(uint32_t* dev_states, int64_t n, int64_t first_index, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (n < 0 || (n > 0 && !dev_states)) return fail(WGRT_ERR_INVALID, "bad arguments");
  CUDA_TRY(launch_seed_rng(dev_states, n, first_index, static_cast<cudaStream_t>(stream)));
  return WGRT_OK;
}

int wgrt_counters_read(uint64_t* out, int n) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!out || n < 0) return fail(WGRT_ERR_INVALID, "bad arguments");
  Workspace* w = nullptr;
  int rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  if (n > WGRT_NUM_COUNTERS) n = WGRT_NUM_COUNTERS;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out, w->counters(), sizeof(uint64_t) * n, cudaMemcpyDeviceToHost));
  return WGRT_OK;
}

int wgrt_counters_reset(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  Workspace* w = nullptr;
  int rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemset(w->counters(), 0, sizeof(uint64_t) * WGRT_NUM_COUNTERS));
  return WGRT_OK;
}

int wgrt_debug_locate(const double* verts, int64_t n_verts, const int64_t* offsets, int64_t n_polys,
                      const double* px, const double* py, int64_t n_points, int32_t* out, int mode) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!verts || !offsets || !px || !py || !out || n_verts < 0 || n_polys < 0 || n_polys > 250 || n_points < 0)
    return fail(WGRT_ERR_INVALID, "bad arguments");
  Workspace* w = nullptr;
  int rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  const size_t vb = (size_t)n_verts * 16, ob = (size_t)(n_polys + 1) * 8, pb = (size_t)n_points * 8;
  CUDA_TRY(w->arena.reserve(padded(vb) + padded(ob) + 2 * padded(pb) + padded((size_t)n_points * 4)));
  Arena ar{static_cast<char*>(w->arena.ptr), w->arena.bytes};
  double* d_v = static_cast<double*>(ar.take(vb));
  int64_t* d_o = static_cast<int64_t*>(ar.take(ob));
  double* d_x = static_cast<double*>(ar.take(pb));
  double* d_y = static_cast<double*>(ar.take(pb));
  int32_t* d_out = static_cast<int32_t*>(ar.take((size_t)n_points * 4));
  if (vb) CUDA_TRY(cudaMemcpy(d_v, verts, vb, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d_o, offsets, ob, cudaMemcpyHostToDevice));
  if (pb) {
    CUDA_TRY(cudaMemcpy(d_x, px, pb, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_y, py, pb, cudaMemcpyHostToDevice));
  }
  if (mode == 0) {
    CUDA_TRY(launch_debug_locate_literal(d_v, d_o, n_polys, d_x, d_y, n_points, d_out, nullptr));
  } else {
    RegionSet rs;
    const double* vs[NUM_REGIONS] = {d_v, d_v, d_v, d_v, d_v};
    const int64_t* os[NUM_REGIONS] = {d_o, d_o, d_o, d_o, d_o};
    const int64_t nv[NUM_REGIONS] = {0, 0, 0, n_verts, 0};
    const int64_t np[NUM_REGIONS] = {0, 0, 0, n_polys, 0};
    rc = setup_regions(*w, rs, vs, os, nv, np);
    if (rc != WGRT_OK) return rc;
    CUDA_TRY(launch_region_build(rs, true, nullptr));
    w->index_stale = true;  // the walk's index was overwritten
    CUDA_TRY(launch_debug_locate_grid(rs, REG_FC, d_x, d_y, n_points, d_out, w->counters(), mode == 2 ? 1 : (mode == 3 ? 2 : 0), nullptr));
  }
  CUDA_TRY(cudaDeviceSynchronize());
  if (n_points) CUDA_TRY(cudaMemcpy(out, d_out, (size_t)n_points * 4, cudaMemcpyDeviceToHost));
  return WGRT_OK;
}

int wgrt_debug_efield(const double* ete, const double* etm, const double* delta, const double* jones,
                      int64_t n, double* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!ete || !etm || !delta || !jones || !out || n < 0) return fail(WGRT_ERR_INVALID, "bad arguments");
  Workspace* w = nullptr;
  int rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  const size_t b = (size_t)n * 8;
  CUDA_TRY(w->arena.reserve(3 * padded(b) + padded(8 * b) + padded(3 * b)));
  Arena ar{static_cast<char*>(w->arena.ptr), w->arena.bytes};
  double* d_te = static_cast<double*>(ar.take(b));
  double* d_tm = static_cast<double*>(ar.take(b));
  double* d_dl = static_cast<double*>(ar.take(b));
  double* d_j = static_cast<double*>(ar.take(8 * b));
  double* d_o = static_cast<double*>(ar.take(3 * b));
  if (n) {
    CUDA_TRY(cudaMemcpy(d_te, ete, b, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_tm, etm, b, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_dl, delta, b, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_j, jones, 8 * b, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(launch_debug_efield(d_te, d_tm, d_dl, d_j, n, d_o, nullptr));
  CUDA_TRY(cudaDeviceSynchronize());
  if (n) CUDA_TRY(cudaMemcpy(out, d_o, 3 * b, cudaMemcpyDeviceToHost));
  return WGRT_OK;
}

int wgrt_debug_xorshift(uint32_t* states, int64_t n, int draws, double* out_last) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!states || n < 0 || draws < 0) return fail(WGRT_ERR_INVALID, "bad arguments");
  Workspace* w = nullptr;
  int rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  CUDA_TRY(w->arena.reserve(padded((size_t)n * 4) + padded((size_t)n * 8)));
  Arena ar{static_cast<char*>(w->arena.ptr), w->arena.bytes};
  uint32_t* d_s = static_cast<uint32_t*>(ar.take((size_t)n * 4));
  double* d_u = static_cast<double*>(ar.take((size_t)n * 8));
  if (n) CUDA_TRY(cudaMemcpy(d_s, states, (size_t)n * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(launch_debug_xorshift(d_s, n, draws, d_u, nullptr));
  CUDA_TRY(cudaDeviceSynchronize());
  if (n) {
    CUDA_TRY(cudaMemcpy(states, d_s, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (out_last) CUDA_TRY(cudaMemcpy(out_last, d_u, (size_t)n * 8, cudaMemcpyDeviceToHost));
  }
  return WGRT_OK;
}

int wgrt_debug_check_failures(uint64_t* out, int reset) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!out) return fail(WGRT_ERR_INVALID, "bad arguments");
  CUDA_TRY(cudaDeviceSynchronize());
  unsigned long long v = 0;
  cudaError_t e = walk_check_failures(&v, reset != 0);
  if (e == cudaErrorNotSupported) return fail(WGRT_ERR_UNSUPPORTED, "not a checked build (libwgrt_checked.so)");
  CUDA_TRY(e);
  *out = v;
  return WGRT_OK;
}

int wgrt_debug_set_tie_tolerance(double tol) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (tol != tol) return fail(WGRT_ERR_INVALID, "tolerance is NaN");
  set_tie_tolerance(tol);
  return WGRT_OK;
}

int wgrt_debug_deposit_inside(const double* rect, const double* px, const double* py, int64_t n, int32_t* out, int mode) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!rect || !px || !py || !out || n < 0 || (mode != 0 && mode != 1)) return fail(WGRT_ERR_INVALID, "bad arguments");
  Workspace* w = nullptr;
  int rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  const size_t pb = (size_t)n * 8;
  CUDA_TRY(cudaDeviceSynchronize());   // the arena may still be in use by an asynchronous launch
  CUDA_TRY(w->arena.reserve(padded(64) + 2 * padded(pb) + padded((size_t)n * 4)));
  Arena ar{static_cast<char*>(w->arena.ptr), w->arena.bytes};
  double* d_r = static_cast<double*>(ar.take(64));
  double* d_x = static_cast<double*>(ar.take(pb));
  double* d_y = static_cast<double*>(ar.take(pb));
  int32_t* d_o = static_cast<int32_t*>(ar.take((size_t)n * 4));
  CUDA_TRY(cudaMemcpy(d_r, rect, 64, cudaMemcpyHostToDevice));
  if (n) {
    CUDA_TRY(cudaMemcpy(d_x, px, pb, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_y, py, pb, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(launch_debug_deposit_inside(d_r, d_x, d_y, n, d_o, mode == 0 ? 1 : 0, nullptr));
  CUDA_TRY(cudaDeviceSynchronize());
  if (n) CUDA_TRY(cudaMemcpy(out, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return WGRT_OK;
}

int wgrt_debug_fma_peak(double* fp64_tflops, double* fp32_tflops) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!fp64_tflops || !fp32_tflops) return fail(WGRT_ERR_INVALID, "bad arguments");
  Workspace* w = nullptr;
  int rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(launch_fma_peak(w->num_sms, fp64_tflops, fp32_tflops));
  return WGRT_OK;
}

int wgrt_eval_pupil_sums(const float* dev_EB, int64_t L, int64_t Yf, int64_t Xf, int64_t EBy, int64_t EBx,
                         int mask_size, int step_y, int step_x, float* dev_out, float* dev_cell_sums,
                         void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!dev_EB || L < 0 || Yf < 0 || Xf < 0 || EBy <= 0 || EBx <= 0 || mask_size <= 0 || step_y <= 0 || step_x <= 0)
    return fail(WGRT_ERR_INVALID, "bad arguments");
  if (mask_size > 220) return fail(WGRT_ERR_UNSUPPORTED, "pupil mask wider than 220 bins");
  CUDA_TRY(launch_pupil_sums(dev_EB, L, Yf, Xf, EBy, EBx, mask_size, step_y, step_x, dev_out, dev_cell_sums,
                             static_cast<cudaStream_t>(stream)));
  return WGRT_OK;
}

int wgrt_eval_metrics(const float* dev_perceive, int64_t Yf, int64_t Xf, int n_epy, int n_epx,
                      const wgrt_eval_params_t* params, double* dev_metrics, float* dev_image, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!dev_perceive || !params || !dev_metrics || Yf < 0 || Xf < 0 || n_epy < 0 || n_epx < 0)
    return fail(WGRT_ERR_INVALID, "bad arguments");
  CUDA_TRY(launch_eval_metrics(dev_perceive, Yf, Xf, n_epy, n_epx, *params, dev_metrics, dev_image,
                               static_cast<cudaStream_t>(stream)));
  return WGRT_OK;
}

int wgrt_eval_metrics_host(const float* perceive, int64_t Yf, int64_t Xf, int n_epy, int n_epx,
                           const wgrt_eval_params_t* params, double* metrics, float* image) {
  if (!perceive || !params || !metrics || Yf < 0 || Xf < 0 || n_epy < 0 || n_epx < 0) return fail(WGRT_ERR_INVALID, "bad arguments");
  const size_t n_ep = static_cast<size_t>(n_epy) * n_epx, pix = static_cast<size_t>(Yf * Xf);
  const size_t pb = 3 * pix * n_ep * 4, mb = n_ep * WGRT_EVAL_NUM * sizeof(double), ib = image ? pix * 3 * n_ep * 4 : 0;
  float *d_p, *d_i;
  double* d_m;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    Workspace* w = nullptr;
    int rc = get_workspace(&w);
    if (rc != WGRT_OK) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(w->arena.reserve(padded(pb) + padded(mb) + padded(ib)));
    Arena ar{static_cast<char*>(w->arena.ptr), w->arena.bytes};
    d_p = static_cast<float*>(ar.take(pb));
    d_m = static_cast<double*>(ar.take(mb));
    d_i = ib ? static_cast<float*>(ar.take(ib)) : nullptr;
    if (pb) CUDA_TRY(cudaMemcpy(d_p, perceive, pb, cudaMemcpyHostToDevice));
  }
  int rc = wgrt_eval_metrics(d_p, Yf, Xf, n_epy, n_epx, params, d_m, d_i, nullptr);
  if (rc != WGRT_OK) return rc;
  CUDA_TRY(cudaDeviceSynchronize());
  if (mb) CUDA_TRY(cudaMemcpy(metrics, d_m, mb, cudaMemcpyDeviceToHost));
  if (ib) CUDA_TRY(cudaMemcpy(image, d_i, ib, cudaMemcpyDeviceToHost));
  return WGRT_OK;
}

int wgrt_bins_pack_u8(const float* dev_bins, int64_t n, uint8_t* dev_out, uint32_t* dev_stats, float limit, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!dev_bins || !dev_out || !dev_stats || n < 0 || (n & 3) || (reinterpret_cast<uintptr_t>(dev_bins) & 15) ||
      (reinterpret_cast<uintptr_t>(dev_out) & 3))
    return fail(WGRT_ERR_INVALID, "bins_pack_u8: n must be a multiple of 4, bins 16-byte and out 4-byte aligned");
  Workspace* w = nullptr;
  int rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  if (!(limit >= 0.f && limit <= 255.f)) return fail(WGRT_ERR_INVALID, "bins_pack_u8: limit must be within [0, 255]");
  CUDA_TRY(launch_bins_pack_u8(dev_bins, n, dev_out, dev_stats, limit, w->num_sms, static_cast<cudaStream_t>(stream)));
  return WGRT_OK;
}

int wgrt_bins_unpack_u8(const uint8_t* dev_in, int64_t n, float* dev_bins, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!dev_bins || !dev_in || n < 0 || (n & 3) || (reinterpret_cast<uintptr_t>(dev_bins) & 15) ||
      (reinterpret_cast<uintptr_t>(dev_in) & 3))
    return fail(WGRT_ERR_INVALID, "bins_unpack_u8: n must be a multiple of 4, bins 16-byte and in 4-byte aligned");
  Workspace* w = nullptr;
  int rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  CUDA_TRY(launch_bins_unpack_u8(dev_in, n, dev_bins, w->num_sms, static_cast<cudaStream_t>(stream)));
  return WGRT_OK;
}

int wgrt_eval_pupil_sums_host(const float* EB, int64_t L, int64_t Yf, int64_t Xf, int64_t EBy, int64_t EBx,
                              int mask_size, int step_y, int step_x, float* out, float* cell_sums) {
  if (!EB || EBy <= 0 || EBx <= 0 || mask_size <= 0 || step_y <= 0 || step_x <= 0)
    return fail(WGRT_ERR_INVALID, "bad arguments");
  Workspace* w = nullptr;
  size_t tiles, n_out, eb_b;
  float *d_eb, *d_out, *d_cs;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    int rc = get_workspace(&w);
    if (rc != WGRT_OK) return rc;
    tiles = (size_t)(L * Yf * Xf);
    const size_t n_epy = EBy >= mask_size ? (size_t)((EBy - mask_size) / step_y + 1) : 0;
    const size_t n_epx = EBx >= mask_size ? (size_t)((EBx - mask_size) / step_x + 1) : 0;
    n_out = tiles * n_epy * n_epx;
    eb_b = tiles * (size_t)(EBy * EBx) * 4;
    CUDA_TRY(w->arena.reserve(padded(eb_b) + padded(n_out * 4) + padded(tiles * 4)));
    Arena ar{static_cast<char*>(w->arena.ptr), w->arena.bytes};
    d_eb = static_cast<float*>(ar.take(eb_b));
    d_out = static_cast<float*>(ar.take(n_out * 4));
    d_cs = static_cast<float*>(ar.take(tiles * 4));
    if (eb_b) CUDA_TRY(cudaMemcpy(d_eb, EB, eb_b, cudaMemcpyHostToDevice));
  }
  int rc = wgrt_eval_pupil_sums(d_eb, L, Yf, Xf, EBy, EBx, mask_size, step_y, step_x, d_out, d_cs, nullptr);
  if (rc != WGRT_OK) return rc;
  CUDA_TRY(cudaDeviceSynchronize());
  if (out && n_out) CUDA_TRY(cudaMemcpy(out, d_out, n_out * 4, cudaMemcpyDeviceToHost));
  if (cell_sums && tiles) CUDA_TRY(cudaMemcpy(cell_sums, d_cs, tiles * 4, cudaMemcpyDeviceToHost));
  return WGRT_OK;
}

}  // extern "C"

// ---- legacy deterministic energy-splitting tracer (row f4) ---------------------------------------------
namespace {

int validate_legacy(const wgrt_legacy_problem_t* p) {
  if (!p) return fail(WGRT_ERR_INVALID, "null problem");
  if (p->capacity < 0 || p->useful_count_in < 0 || p->useful_count_in > p->capacity)
    return fail(WGRT_ERR_INVALID, "useful_count_in must be within [0, capacity]");
  if (p->max_steps < 0) return fail(WGRT_ERR_INVALID, "max_steps < 0");
  if (p->X <= 0 || p->Y <= 0 || p->EBx <= 0 || p->EBy <= 0) return fail(WGRT_ERR_INVALID, "X, Y, EBy, EBx must be positive");
  if (p->n_FC < 0 || p->n_OC < 0 || p->n_FC > 250 || p->n_OC > 250) return fail(WGRT_ERR_INVALID, "n_FC / n_OC must be within [0, 250]");
  if (p->C_ic < 24 || p->C_fc < 20 || p->C_oc < 26)
    return fail(WGRT_ERR_INVALID, "LUT channel counts too small (need C_ic>=24, C_fc>=20, C_oc>=26)");
  if (p->IC_n < 0 || p->FC_n < 0 || p->OC_n < 0 || p->eff_reg1_n < 0 || p->eff_reg2_n < 0) return fail(WGRT_ERR_INVALID, "negative vertex count");
#define NEED(f) \
  if (!p->f) return fail(WGRT_ERR_INVALID, "null pointer: " #f)
  NEED(total_ray_counter); NEED(IC); NEED(FC); NEED(FC_offset); NEED(OC); NEED(OC_offset); NEED(eff_reg1); NEED(eff_reg2);
  NEED(eff_reg_FOV); NEED(eff_reg_FOV_range); NEED(lut_ic1); NEED(lut_ic2); NEED(lut_fc1); NEED(lut_fc2); NEED(lut_oc);
  NEED(lut_TIR); NEED(lut_gap); NEED(matrix_EB);
  if (p->capacity > 0) NEED(vectors);
#undef NEED
  return WGRT_OK;
}

// the region index works on the polygons alone: present them as a (ray-less) full-colour problem
wgrt_problem_t polygons_of(const wgrt_legacy_problem_t& lp) {
  wgrt_problem_t p{};
  p.IC = lp.IC; p.IC_n = lp.IC_n;
  p.FC = lp.FC; p.FC_n = lp.FC_n; p.FC_offset = lp.FC_offset; p.n_FC = lp.n_FC;
  p.OC = lp.OC; p.OC_n = lp.OC_n; p.OC_offset = lp.OC_offset; p.n_OC = lp.n_OC;
  p.eff_reg1 = lp.eff_reg1; p.eff_reg1_n = lp.eff_reg1_n;
  p.eff_reg2 = lp.eff_reg2; p.eff_reg2_n = lp.eff_reg2_n;
  return p;
}

int legacy_step_device(Workspace& w, const wgrt_legacy_problem_t& lp, bool soa, cudaStream_t stream, bool build_index) {
  const wgrt_problem_t poly = polygons_of(lp);
  RegionSet rs;
  int rc = region_set_of(w, poly, rs);
  if (rc != WGRT_OK) return rc;
  if (build_index) {
    CUDA_TRY(launch_region_build(rs, w.index_stale, stream));
    w.index_stale = false;
  }
  // dropped-children counter: the spare counter slot behind the event counters
  CUDA_TRY(launch_legacy_step(lp, rs, soa, w.counters() + WGRT_NUM_COUNTERS, stream));
  return WGRT_OK;
}

}  // namespace

extern "C" int wgrt_legacy_problem_size(void) { return static_cast<int>(sizeof(wgrt_legacy_problem_t)); }

extern "C" int wgrt_legacy_step(const wgrt_legacy_problem_t* p, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  int rc = validate_legacy(p);
  if (rc != WGRT_OK) return rc;
  Workspace* w = nullptr;
  rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  if (p->useful_count_in == 0) return WGRT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const wgrt_problem_t poly = polygons_of(*p);
  rc = validate_device_offsets(*w, poly, st);
  if (rc != WGRT_OK) return rc;
  rc = order_after_last_use(*w, st);
  if (rc != WGRT_OK) return rc;
  rc = legacy_step_device(*w, *p, false, st, true);
  if (rc != WGRT_OK) return rc;
  return record_last_use(*w, st);
}

extern "C" int wgrt_legacy_pack_active(const double* dev_src, double* dev_dst, int64_t src_len, int32_t* dev_out_count,
                                       void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (src_len < 0 || !dev_out_count || (src_len > 0 && (!dev_src || !dev_dst))) return fail(WGRT_ERR_INVALID, "bad arguments");
  CUDA_TRY(launch_legacy_pack(dev_src, dev_dst, src_len, 0, 0, false, dev_out_count, static_cast<cudaStream_t>(stream)));
  return WGRT_OK;
}

extern "C" int wgrt_legacy_trace_host(const wgrt_legacy_problem_t* hp, int max_generations, uint64_t* stats) {
  std::lock_guard<std::mutex> lk(g_mu);
  int rc = validate_legacy(hp);
  if (rc != WGRT_OK) return rc;
  if (max_generations < 0) return fail(WGRT_ERR_INVALID, "max_generations < 0");
  rc = check_offsets(hp->FC_offset, hp->n_FC, hp->FC_n, "FC_offset");
  if (rc != WGRT_OK) return rc;
  rc = check_offsets(hp->OC_offset, hp->n_OC, hp->OC_n, "OC_offset");
  if (rc != WGRT_OK) return rc;
  if (hp->capacity >= (int64_t(1) << 31)) return fail(WGRT_ERR_UNSUPPORTED, "capacity must be below 2^31 rows");
  Workspace* w = nullptr;
  rc = get_workspace(&w);
  if (rc != WGRT_OK) return rc;
  CUDA_TRY(cudaDeviceSynchronize());   // the arena may be in use by an asynchronous launch
  const size_t cap = static_cast<size_t>(hp->capacity), fov = static_cast<size_t>(hp->X * hp->Y);
  const size_t rows_b = cap * WGRT_LEGACY_COLS * sizeof(double);
  const size_t eb_b = fov * static_cast<size_t>(hp->EBy * hp->EBx) * 4;
  wgrt_legacy_problem_t dp = *hp;
  double *qa = nullptr, *qb = nullptr, *aos = nullptr;
  int32_t* d_cnt = nullptr;   // [0] child counter, [1] packed count
  struct Item { const void* src; void** dst; size_t bytes; };
  std::vector<Item> items = {
      {nullptr, (void**)&qa, rows_b}, {nullptr, (void**)&qb, rows_b}, {hp->vectors, (void**)&aos, rows_b},
      {nullptr, (void**)&d_cnt, 16},
      {hp->IC, (void**)&dp.IC, (size_t)hp->IC_n * 16}, {hp->FC, (void**)&dp.FC, (size_t)hp->FC_n * 16},
      {hp->FC_offset, (void**)&dp.FC_offset, (size_t)(hp->n_FC + 1) * 8}, {hp->OC, (void**)&dp.OC, (size_t)hp->OC_n * 16},
      {hp->OC_offset, (void**)&dp.OC_offset, (size_t)(hp->n_OC + 1) * 8},
      {hp->eff_reg1, (void**)&dp.eff_reg1, (size_t)hp->eff_reg1_n * 16}, {hp->eff_reg2, (void**)&dp.eff_reg2, (size_t)hp->eff_reg2_n * 16},
      {hp->eff_reg_FOV, (void**)&dp.eff_reg_FOV, fov * 64}, {hp->eff_reg_FOV_range, (void**)&dp.eff_reg_FOV_range, fov * 32},
      {hp->lut_ic1, (void**)&dp.lut_ic1, fov * hp->C_ic * 16}, {hp->lut_ic2, (void**)&dp.lut_ic2, fov * hp->C_ic * 16},
      {hp->lut_fc1, (void**)&dp.lut_fc1, fov * hp->n_FC * hp->C_fc * 16}, {hp->lut_fc2, (void**)&dp.lut_fc2, fov * hp->n_FC * hp->C_fc * 16},
      {hp->lut_oc, (void**)&dp.lut_oc, fov * hp->n_OC * hp->C_oc * 16},
      {hp->lut_TIR, (void**)&dp.lut_TIR, fov * 32}, {hp->lut_gap, (void**)&dp.lut_gap, fov * 64},
      {hp->matrix_EB, (void**)&dp.matrix_EB, eb_b},
  };
  size_t total = 0;
  for (auto& it : items) total += padded(it.bytes);
  CUDA_TRY(w->arena.reserve(total));
  Arena ar{static_cast<char*>(w->arena.ptr), w->arena.bytes};
  for (auto& it : items) {
    *it.dst = ar.take(it.bytes);
    if (!*it.dst) return fail(WGRT_ERR_CUDA, "arena overflow");
  }
  const size_t n0 = static_cast<size_t>(hp->useful_count_in);
  for (auto& it : items) {
    size_t b = it.bytes;
    if (it.src == hp->vectors) b = n0 * WGRT_LEGACY_COLS * sizeof(double);   // only the initial rows carry data
    if (it.src && b) CUDA_TRY(cudaMemcpy(*it.dst, it.src, b, cudaMemcpyHostToDevice));
  }
  unsigned long long* d_dropped = w->counters() + WGRT_NUM_COUNTERS;
  CUDA_TRY(cudaMemset(d_dropped, 0, sizeof(unsigned long long)));
  CUDA_TRY(launch_legacy_transpose(aos, qa, hp->useful_count_in, hp->capacity, true, nullptr));

  uint64_t st[8] = {0, 0, 0, 0, 0, static_cast<uint64_t>(hp->useful_count_in), 0, 0};
  int64_t count = hp->useful_count_in;
  bool first = true;
  while (count > 0 && st[0] < static_cast<uint64_t>(max_generations)) {
    const int32_t start[2] = {static_cast<int32_t>(count), 0};
    CUDA_TRY(cudaMemcpy(d_cnt, start, 8, cudaMemcpyHostToDevice));
    wgrt_legacy_problem_t gp = dp;
    gp.vectors = qa;
    gp.useful_count_in = count;
    gp.total_ray_counter = d_cnt;
    rc = legacy_step_device(*w, gp, true, nullptr, first);
    if (rc != WGRT_OK) return rc;
    first = false;
    int32_t after = 0;
    CUDA_TRY(cudaMemcpy(&after, d_cnt, 4, cudaMemcpyDeviceToHost));
    const int64_t live_end = std::min<int64_t>(after, hp->capacity);
    st[2] += static_cast<uint64_t>(count);
    st[3] += static_cast<uint64_t>(after - count);
    CUDA_TRY(launch_legacy_pack(qa, qb, live_end, hp->capacity, hp->capacity, true, d_cnt + 1, nullptr));
    int32_t packed = 0;
    CUDA_TRY(cudaMemcpy(&packed, d_cnt + 1, 4, cudaMemcpyDeviceToHost));
    std::swap(qa, qb);
    count = packed;
    st[0] += 1;
    st[5] = std::max<uint64_t>(st[5], static_cast<uint64_t>(count));
  }
  st[1] = static_cast<uint64_t>(count);
  unsigned long long dropped = 0;
  CUDA_TRY(cudaMemcpy(&dropped, d_dropped, sizeof dropped, cudaMemcpyDeviceToHost));
  st[4] = dropped;
  if (count > 0) {
    CUDA_TRY(launch_legacy_transpose(qa, aos, count, hp->capacity, false, nullptr));
    CUDA_TRY(cudaMemcpy(hp->vectors, aos, static_cast<size_t>(count) * WGRT_LEGACY_COLS * sizeof(double), cudaMemcpyDeviceToHost));
  }
  CUDA_TRY(cudaMemcpy(hp->matrix_EB, dp.matrix_EB, eb_b, cudaMemcpyDeviceToHost));
  if (stats) memcpy(stats, st, sizeof st);
  return WGRT_OK;
}
