// ctp_api.cu -- C ABI of libctopprm_b200.so (see include/ctopprm_b200.h).
// Host-side plumbing only: handle, device layouts, scratch arenas, kernel dispatch.
#include "../../include/ctopprm_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "ctp_device.cuh"
#include "ctp_graph.cuh"
#include "ctp_edt.cuh"
#include "ctp_paths.cuh"
#include "ctp_prm.cuh"
#include "ctp_cluster.cuh"

// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(e_ == cudaErrorMemoryAllocation ? CTP_ERR_NOMEM : CTP_ERR_CUDA,         \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKR(call)               \
  do {                          \
    int r_ = (call);            \
    if (r_ != CTP_OK) return r_; \
  } while (0)
#define REQUIRE(cond, msg) \
  do {                     \
    if (!(cond)) return fail(CTP_ERR_INVALID, "%s: %s", __func__, msg); \
  } while (0)

struct Arena {
  char *base = nullptr;
  size_t cap = 0, used = 0;
};

struct ctp_map {
  int device = 0;
  int n = 0;
  int layouts = 0, active = CTP_L_PLAIN;
  MapDev dev{};
  float *d_plain = nullptr, *d_cell8 = nullptr, *d_brick = nullptr;
  cudaArray_t arr = nullptr;
  cudaTextureObject_t tex = 0;
  int tex_shift = 0;
  Arena ws;  // scratch of the _dev entry points
  Arena io;  // device staging of the host-pointer entry points
  unsigned long long *d_lookups = nullptr;
  int sm_count = 148;
  size_t map_bytes = 0;
  // stage profiling (ctp_profile): pairs of events recorded on the caller's stream
  double knn_per_cell = 22.0;  // expected points inside a ball of one cell size (see k_knn_setup)
  int sssp_cluster = -1;      // shortest paths on a cluster of 8 CTAs per query: -1 automatic, 0 never, 1 whenever possible
  int sssp_frontier = 1;      // 1 = queue-based edge-parallel shortest-path kernel, 0 = flag-scan kernel
  int edge_order = -1;        // edge kernel walks rows in cell-sorted order (1), node order (0), or decides (-1):
                              // sorted pays off when the gathers leave the L2 (grid > ~100 MB) or are not texture fetches
  int knn_tile = 1;           // 1 = direct search + listed wider passes + warp ring search, 0 = per-thread ring search only
  unsigned long long launches = 0;  // kernels enqueued by this handle (ctp_launch_counter_read)
  int prm_group = 1;  // ctp_build_prm: 1 = flat lookup-parallel kernel (kNN edges are ~3-12 samples long)
  // chunk pipeline of the host-pointer ctp_sample_multiple
  bool pipe_ready = false;
  cudaStream_t pipe_st[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_in[2], ev_comp[2], ev_small[2], ev_out[2], ev_rej[2], ev_rejd[2];
  int64_t *h_off = nullptr;  // pinned staging of per-chunk CSR offsets
  int32_t *h_nf = nullptr;   // pinned staging of per-chunk accept counts
  unsigned long long io_h2d = 0, io_d2h = 0;  // bytes moved by the chunk pipelines (ctp_io_counter_read)
  double pipe_head_factor = 2.3;  // candidates per node uploaded before the rest of a stream is asked for (0 = all)
  int64_t h_off_cap = 0;
  int64_t pipe_chunk = 0;    // queries per chunk (0 = automatic)
  bool reject_layout_auto = true;  // large grids: rejection reads the CELL8 records when they exist
  bool sssp_cluster_big = true;  // clusters of 16 CTAs (opt-in size) are tried for very small batches until one fails
  int sssp_smem_optin = -1;   // opt-in shared memory per CTA of this device, -1 until the graph kernels were configured
  unsigned knn_smem_set = 0;  // bit K: the dynamic shared-memory limit of the k-NN kernels <K> has been raised
  int64_t pipe_nu_seen = 0;  // largest n_used (candidates consumed by a query) seen by the running pipeline call
  void *rm_block = nullptr;   // device block of the last destroyed ctp_roadmap, kept for the next one (cudaMalloc / cudaFree
  size_t rm_block_cap = 0;    // cost milliseconds each on a context that maps tens of GiB)
  double knn_radius = 0;      // > 0: neighbours farther than this are dropped from the k-NN table (radius-neighbour mode)
  int seg_bin = -1;           // ctp_edge_check: bin large batches by voxel block first: -1 automatic, 0 never, 1 always
  int64_t csr_pair_cap = 0;   // overflow pairs of the adjacency stage (0 = one per node)
  int perq_cluster = -1;      // per-query setup / scan kernels on clusters of 8 CTAs: -1 automatic, 0 never, 1 always
  int occ_cache[32] = {0};    // resident CTAs per SM of the edge kernels, per (layout, lanes per edge): see occ_slot
  char *h_stage = nullptr;    // page-locked staging of the small host-pointer calls (StagedIn / StagedOut)
  size_t h_stage_cap = 0;
  int edge_filter = 1;        // roadmap edges: float32 whole-edge filter in front of the exact kernel: 0 never, 1 where it pays
                              // (calibrated on the first launch), 2 always (CTP_TUNE_EDGE_FILTER)
  double filter_cal_clr = -1e300, filter_cal_cdc = -1e300;  // (threshold, spacing) the calibration below was taken for
  bool filter_pays = true;
  float stat_gmax = 0.f, stat_vmax = 0.f;  // grid statistics behind the filter's error bound (k_map_stats)
  bool stat_finite = false;
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;    // pool, two per recorded span
  std::vector<int> prof_stage;         // stage id of span i (events 2i, 2i+1)
  size_t prof_used = 0;                // spans recorded since the last read
};

// RAII span: records an event pair around a stage when profiling is on
struct StageSpan {
  ctp_map *m;
  cudaStream_t st;
  size_t slot = (size_t)-1;
  StageSpan(ctp_map *mm, int stage, cudaStream_t s) : m(mm), st(s) {
    if (!m->prof) return;
    if (m->prof_used * 2 + 2 > m->prof_ev.size()) {
      for (int i = 0; i < 2; i++) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        m->prof_ev.push_back(e);
      }
      m->prof_stage.push_back(stage);
    }
    slot = m->prof_used++;
    m->prof_stage[slot] = stage;
    cudaEventRecord(m->prof_ev[2 * slot], st);
  }
  ~StageSpan() {
    if (slot != (size_t)-1) cudaEventRecord(m->prof_ev[2 * slot + 1], st);
  }
};

static int arena_reserve(Arena &a, size_t bytes) {
  if (bytes <= a.cap) return CTP_OK;
  CK(cudaDeviceSynchronize());  // nothing may still be using the old block
  if (a.base) CK(cudaFree(a.base));
  a.base = nullptr;
  a.cap = 0;
  size_t want = bytes + bytes / 4 + (1 << 20);
  CK(cudaMalloc(&a.base, want));
  a.cap = want;
  return CTP_OK;
}
static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
template <typename T> static T *arena_take(Arena &a, size_t count) {
  T *p = reinterpret_cast<T *>(a.base + a.used);
  a.used += align_up(count * sizeof(T));
  return p;
}

extern "C" const char *ctp_last_error(void) { return g_err; }
extern "C" int ctp_version(void) { return 100; }
extern "C" int ctp_device_count(int *count) {
  REQUIRE(count, "null");
  CK(cudaGetDeviceCount(count));
  return CTP_OK;
}

// ---------------------------------- map handle --------------------------------------------
static int map_build_layouts(ctp_map *m) {
  const int n = m->n;
  if (m->layouts & CTP_L_CELL8) {
    const size_t n1 = (size_t)n - 1, bytes = n1 * n1 * n1 * 32;
    CK(cudaMalloc(&m->d_cell8, bytes));
    m->map_bytes += bytes;
    k_build_cell8<<<m->sm_count * 8, 256>>>(m->d_plain, n, m->d_cell8);
    m->launches++;
    CK(cudaGetLastError());
  }
  if (m->layouts & CTP_L_BRICK) {
    const int nb = (n + 3) / 4;
    const size_t bytes = (size_t)nb * nb * nb * 256;
    CK(cudaMalloc(&m->d_brick, bytes));
    m->map_bytes += bytes;
    k_build_brick<<<m->sm_count * 8, 256>>>(m->d_plain, n, nb, m->d_brick);
    m->launches++;
    CK(cudaGetLastError());
  }
  if (m->layouts & CTP_L_TEX) {
    // 2-D gather texture, x-slabs laid out as a mosaic (see Corners<CTP_L_TEX>)
    int shift = 0;
    while ((size_t)n * (((size_t)n + ((size_t)1 << shift) - 1) >> shift) > 32768 && shift < 16) shift++;
    const size_t tiles_u = (size_t)1 << shift, tiles_v = ((size_t)n + tiles_u - 1) >> shift;
    if (tiles_u * n > 32768 || tiles_v * n > 32768)
      return fail(CTP_ERR_INVALID, "CTP_LAYOUT_TEX supports at most 1024 voxels per axis (2-D gather texture limit)");
    m->tex_shift = shift;
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
    CK(cudaMallocArray(&m->arr, &fd, tiles_u * n, tiles_v * n, cudaArrayTextureGather));
    m->map_bytes += tiles_u * tiles_v * (size_t)n * n * 4;
    for (int x = 0; x < n; x++) {
      const size_t tu = (size_t)x & (tiles_u - 1), tv = (size_t)x >> shift;
      CK(cudaMemcpy2DToArrayAsync(m->arr, tu * n * sizeof(float), tv * n, m->d_plain + (size_t)x * n * n,
                                  (size_t)n * sizeof(float), (size_t)n * sizeof(float), (size_t)n,
                                  cudaMemcpyDeviceToDevice, 0));
    }
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof rd);
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = m->arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof td);
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&m->tex, &rd, &td, nullptr));
  }
  {
    // largest adjacent-voxel difference, largest |voxel|, finiteness: the error bound of the float32 verdict filter
    unsigned *d_st = nullptr, h_st[3] = {0, 0, 1};
    CK(cudaMalloc(&d_st, 3 * sizeof(unsigned)));
    cudaError_t e1 = cudaMemsetAsync(d_st, 0, 3 * sizeof(unsigned), 0);
    if (e1 == cudaSuccess) {
      k_map_stats<<<m->sm_count * 8, 256>>>(m->d_plain, n, d_st);
      m->launches++;
      e1 = cudaMemcpy(h_st, d_st, sizeof h_st, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_st);
    CK(e1);
    memcpy(&m->stat_gmax, &h_st[0], 4);
    memcpy(&m->stat_vmax, &h_st[1], 4);
    m->stat_finite = h_st[2] == 0;
  }
  CK(cudaDeviceSynchronize());
  return CTP_OK;
}

static void map_fill_params(ctp_map *m, const double c[3], const double e[3]) {
  MapDev &d = m->dev;
  const double N = (double)m->n;                       // src/esdf_map.cpp:24
  const double max_e = fmax(fmax(e[0], e[1]), e[2]);   // src/esdf_map.cpp:27
  d.cx = c[0]; d.cy = c[1]; d.cz = c[2];
  d.lox = c[0] - e[0] / 2; d.loy = c[1] - e[1] / 2; d.loz = c[2] - e[2] / 2;  // :75
  d.hix = c[0] + e[0] / 2; d.hiy = c[1] + e[1] / 2; d.hiz = c[2] + e[2] / 2;  // :76
  d.s1 = 2.0 / max_e;
  d.s2 = N / 2.0;
  d.nm1 = N - 1.0;
  d.nd = N;
  d.dp = max_e / (N - 1.0);                            // :124
  d.mex = c[0] - 1.0 * max_e / 2.0;                    // :125
  d.mey = c[1] - 1.0 * max_e / 2.0;
  d.mez = c[2] - 1.0 * max_e / 2.0;
  d.n = m->n;
  // Is getClearence's position test implied by the index tests of the interpolation?  Every step of the index map
  // q = ((p - c) * s1 + 1) * s2 is monotone in p, so it is enough to look at the two positions next to the box: the
  // first one above it must already give floor(q) >= N-1 (rejected as "ceil >= N"), the first one below it q < 1
  // (floor <= 0: negative is rejected, 0 makes the kernel do the comparisons).  Same operations, one rounding each,
  // as ctp_w2i on the device.
  {
    const double cc[3] = {c[0], c[1], c[2]};
    const double lo[3] = {d.lox, d.loy, d.loz}, hi[3] = {d.hix, d.hiy, d.hiz};
    bool ok = d.s1 > 0 && d.s2 > 0;
    for (int a = 0; a < 3 && ok; a++) {
      volatile double t = nextafter(hi[a], INFINITY) - cc[a];
      t = t * d.s1;
      t = t + 1.0;
      t = t * d.s2;
      const double q_hi = t;
      t = nextafter(lo[a], -INFINITY) - cc[a];
      t = t * d.s1;
      t = t + 1.0;
      t = t * d.s2;
      const double q_lo = t;
      ok = (floor(q_hi) >= N - 1.0) && (q_lo < 1.0);
    }
    d.box_implied = ok ? 1 : 0;
  }
  {
    // float32 verdict filter (ctp_device.cuh, EdgeFilt): eta bounds the FP64 rounding of the reference's own index
    // coordinate against the affine form qa + t * qv (sample position: 2u * |p|, index transform: 4u * |q|, the same
    // for qa, 2u relative on qv; |p| <= |c| + E * 2^15 / N and |q| <= 2^15 on every edge the filter accepts, in voxels
    // via the factor N / E); eps is the clearance bound for delta_cap, with a safety factor of 1.5.
    const double u = 1.1102230246251565e-16;
    const double cmax = fmax(fmax(fabs(c[0]), fabs(c[1])), fabs(c[2]));
    const double vox_per_m = fabs(d.s1 * d.s2);
    const double pmax = cmax + max_e * 32768.0 / N;
    const double eta = 8.0 * u * pmax * vox_per_m + 16.0 * u * (32768.0 + CTP_FILT_QV_MAX);
    const double delta_cap = ((4.5 * CTP_FILT_QV_MAX + 3.0) * 5.9604644775390625e-08 + eta) * 1.0001 + 1.2e-7;
    const double G = (double)m->stat_gmax, V = (double)m->stat_vmax;
    const double eps = 1.5 * (3.0 * G * (delta_cap + 5.9604644775390625e-08) + 8.0 * 5.9604644775390625e-08 * (G + V)) + 1e-300;
    d.filt_eta = eta;
    d.filt_g = m->stat_gmax;
    d.filt_eps = eps;
    // worth it only while the undecided band stays narrow (a grid of garbage would send every sample to the exact code);
    // maps whose position test is not implied by the index tests (non-cubic extents) keep the exact kernel
    d.filt_ok = (m->stat_finite && d.box_implied && eta < 1e-6 && eps == eps && eps < 1e-3 && d.s1 > 0 && d.s2 > 0 && m->n >= 4) ? 1 : 0;
  }
  d.nb = (m->n + 3) / 4;
  d.plain = m->d_plain;
  d.cell8 = m->d_cell8;
  d.brick = m->d_brick;
  d.tex = m->tex;
  d.tex_cshift = m->tex_shift;
  d.tex_cmask = (1 << m->tex_shift) - 1;
}

static int pick_active(int layouts) {
  if (layouts & CTP_L_TEX) return CTP_L_TEX;
  if (layouts & CTP_L_CELL8) return CTP_L_CELL8;
  if (layouts & CTP_L_BRICK) return CTP_L_BRICK;
  return CTP_L_PLAIN;
}

static int map_create_common(const void *vox, bool is_f64, int n, const double c[3], const double e[3],
                             int device, int layouts, ctp_map **out) {
  REQUIRE(vox && c && e && out, "null argument");
  REQUIRE(n >= 2 && n <= 2048, "voxels per axis must be in [2, 2048]");
  REQUIRE((layouts & ~15) == 0, "unknown layout bits");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
  CK(cudaSetDevice(device));
  ctp_map *m = new (std::nothrow) ctp_map();
  if (!m) return fail(CTP_ERR_NOMEM, "host allocation failed");
  m->device = device;
  m->n = n;
  m->layouts = layouts | CTP_L_PLAIN;
  cudaDeviceProp prop;
  cudaError_t pe = cudaGetDeviceProperties(&prop, device);
  if (pe == cudaSuccess) m->sm_count = prop.multiProcessorCount;
  const size_t cnt = (size_t)n * n * n;
  int rc = CTP_OK;
  auto body = [&]() -> int {
    CK(cudaMalloc(&m->d_plain, cnt * sizeof(float)));
    m->map_bytes += cnt * sizeof(float);
    if (!is_f64) {
      CK(cudaMemcpy(m->d_plain, vox, cnt * sizeof(float), cudaMemcpyHostToDevice));
    } else {
      // chunked narrow f64 -> f32 on the device
      const size_t chunk = (size_t)1 << 24;
      double *tmp = nullptr;
      CK(cudaMalloc(&tmp, std::min(chunk, cnt) * sizeof(double)));
      for (size_t o = 0; o < cnt; o += chunk) {
        const size_t c2 = std::min(chunk, cnt - o);
        cudaError_t e1 = cudaMemcpy(tmp, (const double *)vox + o, c2 * sizeof(double), cudaMemcpyHostToDevice);
        if (e1 != cudaSuccess) { cudaFree(tmp); CK(e1); }
        k_narrow_f64<<<m->sm_count * 4, 256>>>(tmp, c2, m->d_plain + o);
        m->launches++;
      }
      cudaError_t e2 = cudaDeviceSynchronize();
      cudaFree(tmp);
      CK(e2);
    }
    CK(cudaMalloc(&m->d_lookups, 4 * sizeof(unsigned long long)));  // [0] lookups, [1] / [2] edges decided / seen by the whole-edge filter
    CK(cudaMemset(m->d_lookups, 0, 4 * sizeof(unsigned long long)));
    CKR(map_build_layouts(m));
    return CTP_OK;
  };
  rc = body();
  if (rc != CTP_OK) {
    ctp_map_destroy(m);
    return rc;
  }
  m->active = pick_active(m->layouts);
  map_fill_params(m, c, e);
  *out = m;
  return CTP_OK;
}

extern "C" int ctp_map_create(const float *vox, int n, const double c[3], const double e[3], int device,
                              int layouts, ctp_map **out) {
  return map_create_common(vox, false, n, c, e, device, layouts, out);
}
extern "C" int ctp_map_create_f64(const double *vox, int n, const double c[3], const double e[3], int device,
                                  int layouts, ctp_map **out) {
  return map_create_common(vox, true, n, c, e, device, layouts, out);
}

// common part of the generators: a handle whose plain grid is produced on the device by `fill`
template <typename Fill>
static int map_create_generated(int n, const double c[3], const double e[3], int device, int layouts, ctp_map **out,
                                Fill fill) {
  REQUIRE(c && e && out, "null argument");
  REQUIRE(n >= 2 && n <= 2048, "voxels per axis must be in [2, 2048]");
  REQUIRE((layouts & ~15) == 0, "unknown layout bits");
  REQUIRE(e[0] == e[1] && e[1] == e[2], "the generator places voxels on a cubic lattice: extents must be equal");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
  CK(cudaSetDevice(device));
  ctp_map *m = new (std::nothrow) ctp_map();
  if (!m) return fail(CTP_ERR_NOMEM, "host allocation failed");
  m->device = device;
  m->n = n;
  m->layouts = layouts | CTP_L_PLAIN;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) m->sm_count = prop.multiProcessorCount;
  const size_t cnt = (size_t)n * n * n;
  auto body = [&]() -> int {
    CK(cudaMalloc(&m->d_plain, cnt * sizeof(float)));
    m->map_bytes += cnt * sizeof(float);
    CKR(fill(m));
    CK(cudaMalloc(&m->d_lookups, 4 * sizeof(unsigned long long)));  // [0] lookups, [1] / [2] edges decided / seen by the whole-edge filter
    CK(cudaMemset(m->d_lookups, 0, 4 * sizeof(unsigned long long)));
    CKR(map_build_layouts(m));
    return CTP_OK;
  };
  const int rc = body();
  if (rc != CTP_OK) {
    ctp_map_destroy(m);
    return rc;
  }
  m->active = pick_active(m->layouts);
  map_fill_params(m, c, e);
  *out = m;
  return CTP_OK;
}

// frees a list of device pointers on scope exit
struct DevScratch {
  std::vector<void *> p;
  ~DevScratch() { for (void *x : p) cudaFree(x); }
  template <typename T> cudaError_t get(T **out, size_t bytes) {
    void *v = nullptr;
    cudaError_t e = cudaMalloc(&v, bytes ? bytes : 1);
    if (e == cudaSuccess) p.push_back(v);
    *out = (T *)v;
    return e;
  }
};

// ESDF of vertical cylinders generated on the device (no host grid, no host->device copy of N^3 floats)
extern "C" int ctp_map_create_columns(int n, const double c[3], const double e[3], const double *cx, const double *cy,
                                      const double *r, int n_col, int device, int layouts, ctp_map **out) {
  REQUIRE(n_col >= 0 && (n_col == 0 || (cx && cy && r)), "null argument");
  return map_create_generated(n, c, e, device, layouts, out, [&](ctp_map *m) -> int {
    DevScratch sc;
    double *d_par = nullptr;
    float *d_slab = nullptr;
    CK(sc.get(&d_par, (size_t)std::max(1, n_col) * 3 * sizeof(double)));
    CK(sc.get(&d_slab, (size_t)n * n * sizeof(float)));
    if (n_col) {
      CK(cudaMemcpy(d_par, cx, n_col * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_par + n_col, cy, n_col * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_par + 2 * n_col, r, n_col * 8, cudaMemcpyHostToDevice));
    }
    // voxel i sits at (c - E/2) + i * (E/N): same expression and operation order as synth.axis_coords
    const double step = e[0] / (double)n;
    k_gen_columns<<<(n * n + 255) / 256, 256>>>(n, c[0] - e[0] / 2, c[1] - e[1] / 2, step, d_par, d_par + n_col,
                                                d_par + 2 * n_col, n_col, d_slab);
    k_broadcast_z<<<m->sm_count * 8, 256>>>(d_slab, n, m->d_plain);
    CK(cudaDeviceSynchronize());
    return CTP_OK;
  });
}

// exact signed Euclidean distance field of a voxel occupancy grid (ctp_edt.cuh); occ in .npy order, non-zero =
// occupied.  d2_signed (optional, host, n^3 int32) receives the exact squared voxel distances (negative inside).
extern "C" int ctp_map_create_occupancy(const uint8_t *occ, int n, const double c[3], const double e[3], int device,
                                        int layouts, int32_t *d2_signed, ctp_map **out) {
  REQUIRE(occ, "null argument");
  return map_create_generated(n, c, e, device, layouts, out, [&](ctp_map *m) -> int {
    DevScratch sc;
    const size_t cnt = (size_t)n * n * n;
    const int64_t lines = (int64_t)n * n;
    uint8_t *d_occ = nullptr;
    int32_t *d_a = nullptr, *d_b = nullptr, *d_d2 = nullptr;
    uint16_t *d_s = nullptr, *d_t = nullptr;
    CK(sc.get(&d_occ, cnt));
    CK(sc.get(&d_a, cnt * 4));
    CK(sc.get(&d_b, cnt * 4));
    CK(sc.get(&d_s, cnt * 2));
    CK(sc.get(&d_t, cnt * 2));
    if (d2_signed) CK(sc.get(&d_d2, cnt * 4));
    CK(cudaMemcpy(d_occ, occ, cnt, cudaMemcpyHostToDevice));
    const double pitch = e[0] / (double)n;
    for (int feat = 1; feat >= 0; feat--) {
      k_edt_scan_z<<<(unsigned)((lines + 7) / 8), 256>>>(d_occ, n, feat, d_a);
      // along y: lines (x, z), element stride n, line base x*n*n + z
      k_edt_envelope<<<(unsigned)((lines + 255) / 256), 256>>>(d_a, n, lines, n, (int64_t)n * n, n, d_s, d_t, d_b);
      // along x: lines (y, z), element stride n*n, line base y*n + z
      k_edt_envelope<<<(unsigned)((lines + 255) / 256), 256>>>(d_b, n, lines, lines, 0, (int64_t)n * n, d_s, d_t, d_a);
      k_edt_emit<<<m->sm_count * 8, 256>>>(d_a, d_occ, n, pitch, feat, m->d_plain, d_d2);
      m->launches += 4;
    }
    CK(cudaDeviceSynchronize());
    if (d2_signed) CK(cudaMemcpy(d2_signed, d_d2, cnt * 4, cudaMemcpyDeviceToHost));
    return CTP_OK;
  });
}

// the grid back on the host, in .npy order (x*n*n + y*n + z)
extern "C" int ctp_map_download(ctp_map *m, float *vox) {
  REQUIRE(m && vox, "null");
  CK(cudaSetDevice(m->device));
  CK(cudaMemcpy(vox, m->d_plain, (size_t)m->n * m->n * m->n * sizeof(float), cudaMemcpyDeviceToHost));
  return CTP_OK;
}

extern "C" int ctp_map_destroy(ctp_map *m) {
  if (!m) return CTP_OK;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  if (m->tex) cudaDestroyTextureObject(m->tex);
  if (m->arr) cudaFreeArray(m->arr);
  cudaFree(m->d_plain);
  cudaFree(m->d_cell8);
  cudaFree(m->d_brick);
  cudaFree(m->d_lookups);
  if (m->h_stage) cudaFreeHost(m->h_stage);
  cudaFree(m->ws.base);
  cudaFree(m->io.base);
  if (m->rm_block) cudaFree(m->rm_block);
  for (cudaEvent_t e : m->prof_ev) cudaEventDestroy(e);
  if (m->pipe_ready) {
    for (int i = 0; i < 4; i++) cudaStreamDestroy(m->pipe_st[i]);
    for (int i = 0; i < 2; i++) {
      cudaEventDestroy(m->ev_in[i]);
      cudaEventDestroy(m->ev_comp[i]);
      cudaEventDestroy(m->ev_small[i]);
      cudaEventDestroy(m->ev_out[i]);
      cudaEventDestroy(m->ev_rej[i]);
      cudaEventDestroy(m->ev_rejd[i]);
    }
  }
  if (m->h_off) cudaFreeHost(m->h_off);
  if (m->h_nf) cudaFreeHost(m->h_nf);
  delete m;
  return CTP_OK;
}

extern "C" int ctp_map_set_layout(ctp_map *m, int layout) {
  REQUIRE(m, "null map");
  REQUIRE(layout == CTP_L_PLAIN || layout == CTP_L_CELL8 || layout == CTP_L_TEX || layout == CTP_L_BRICK,
          "layout must be a single CTP_LAYOUT_* value");
  REQUIRE(m->layouts & layout, "layout was not materialised at ctp_map_create");
  m->active = layout;
  return CTP_OK;
}

extern "C" int ctp_set_prm_group(ctp_map *m, int group) {
  REQUIRE(m, "null map");
  REQUIRE(group == 1 || group == 4 || group == 8 || group == 16 || group == 32, "group must be 1, 4, 8, 16 or 32");
  m->prm_group = group;
  return CTP_OK;
}

extern "C" int ctp_set_tunable(ctp_map *m, int key, double value) {
  REQUIRE(m, "null map");
  switch (key) {
    case CTP_TUNE_KNN_PER_CELL:
      REQUIRE(value >= 1.0 && value <= 256.0, "expected points per ball must be in [1, 256]");
      m->knn_per_cell = value;
      return CTP_OK;
    case CTP_TUNE_KNN_TILE:
      m->knn_tile = value != 0.0;
      return CTP_OK;
    case CTP_TUNE_KNN_RADIUS:
      REQUIRE(value >= 0, "radius must be >= 0 (0 = off)");
      m->knn_radius = value;
      return CTP_OK;
    case CTP_TUNE_SEG_BIN:
      m->seg_bin = value < 0 ? -1 : (value != 0.0);
      return CTP_OK;
    case CTP_TUNE_CSR_PAIR_CAP:
      REQUIRE(value >= 0 && value <= 1e9, "pair capacity must be >= 0");
      m->csr_pair_cap = (int64_t)value;
      return CTP_OK;
    case CTP_TUNE_PERQ_CLUSTER:
      m->perq_cluster = value < 0 ? -1 : (value != 0.0);
      return CTP_OK;
    case CTP_TUNE_PIPE_HEAD:
      REQUIRE(value >= 0 && value <= 1e6, "head factor must be >= 0");
      m->pipe_head_factor = value;
      return CTP_OK;
    case CTP_TUNE_SSSP_CLUSTER:
      m->sssp_cluster = value < 0 ? -1 : (value != 0.0);
      return CTP_OK;
    case CTP_TUNE_SSSP_FRONTIER:
      m->sssp_frontier = value != 0.0;
      return CTP_OK;
    case CTP_TUNE_EDGE_ORDER:
      m->edge_order = value < 0 ? -1 : (value != 0.0);
      return CTP_OK;
    case CTP_TUNE_EDGE_FILTER:
      m->edge_filter = value == 2.0 ? 2 : (value != 0.0);
      return CTP_OK;
    case CTP_TUNE_PIPE_CHUNK:
      REQUIRE(value >= 0 && value <= 65535, "queries per chunk must be in [0, 65535]");
      m->pipe_chunk = (int64_t)value;
      return CTP_OK;
    default:
      return fail(CTP_ERR_INVALID, "ctp_set_tunable: unknown key %d", key);
  }
}

extern "C" int ctp_map_info(const ctp_map *m, double info[12]) {
  REQUIRE(m && info, "null");
  info[0] = m->dev.lox; info[1] = m->dev.loy; info[2] = m->dev.loz;
  info[3] = m->dev.hix; info[4] = m->dev.hiy; info[5] = m->dev.hiz;
  info[6] = m->n;
  info[7] = m->active;
  info[8] = m->device;
  info[9] = (double)(m->map_bytes + m->ws.cap + m->io.cap);
  info[10] = m->prm_group;
  info[11] = m->sm_count;
  return CTP_OK;
}

extern "C" int ctp_lookup_counter_read(ctp_map *m, void *stream, uint64_t *lookups, int reset) {
  REQUIRE(m && lookups, "null");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long v = 0;
  CK(cudaMemcpyAsync(&v, m->d_lookups, sizeof v, cudaMemcpyDeviceToHost, st));
  if (reset) CK(cudaMemsetAsync(m->d_lookups, 0, sizeof v, st));
  CK(cudaStreamSynchronize(st));
  *lookups = v;
  return CTP_OK;
}

// edges the whole-edge float32 filter decided / was shown, and the reference-order lookups of the decided ones, since
// the last reset (counted by k_edge_mid)
extern "C" int ctp_filter_counter_read(ctp_map *m, void *stream, uint64_t out[3], int reset) {
  REQUIRE(m && out, "null");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long v[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(v, m->d_lookups + 1, sizeof v, cudaMemcpyDeviceToHost, st));
  if (reset) CK(cudaMemsetAsync(m->d_lookups + 1, 0, sizeof v, st));
  CK(cudaStreamSynchronize(st));
  out[0] = v[0];
  out[1] = v[1];
  out[2] = v[2];
  return CTP_OK;
}

extern "C" int ctp_launch_counter_read(ctp_map *m, uint64_t *launches, int reset) {
  REQUIRE(m && launches, "null");
  *launches = m->launches;
  if (reset) m->launches = 0;
  return CTP_OK;
}

extern "C" int ctp_io_counter_read(ctp_map *m, uint64_t *h2d_bytes, uint64_t *d2h_bytes, int reset) {
  REQUIRE(m && h2d_bytes && d2h_bytes, "null");
  *h2d_bytes = m->io_h2d;
  *d2h_bytes = m->io_d2h;
  if (reset) m->io_h2d = m->io_d2h = 0;
  return CTP_OK;
}

extern "C" int ctp_profile(ctp_map *m, int enable) {
  REQUIRE(m, "null map");
  m->prof = enable != 0;
  return CTP_OK;
}

extern "C" int ctp_profile_read(ctp_map *m, double ms[CTP_STAGE_COUNT], int64_t calls[CTP_STAGE_COUNT], int reset) {
  REQUIRE(m && ms, "null");
  CK(cudaSetDevice(m->device));
  for (int i = 0; i < CTP_STAGE_COUNT; i++) {
    ms[i] = 0;
    if (calls) calls[i] = 0;
  }
  for (size_t i = 0; i < m->prof_used; i++) {
    CK(cudaEventSynchronize(m->prof_ev[2 * i + 1]));
    float t = 0;
    CK(cudaEventElapsedTime(&t, m->prof_ev[2 * i], m->prof_ev[2 * i + 1]));
    const int sg = m->prof_stage[i];
    if (sg >= 0 && sg < CTP_STAGE_COUNT) {
      ms[sg] += t;
      if (calls) calls[sg]++;
    }
  }
  if (reset) m->prof_used = 0;
  return CTP_OK;
}

extern "C" int ctp_host_alloc(void **ptr, size_t bytes) {
  REQUIRE(ptr, "null");
  CK(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return CTP_OK;
}
extern "C" int ctp_host_free(void *ptr) {
  if (ptr) CK(cudaFreeHost(ptr));
  return CTP_OK;
}

#define DISPATCH_L(m, ...)                                                     \
  switch ((m)->active) {                                                       \
    case CTP_L_PLAIN: { constexpr int L = CTP_L_PLAIN; __VA_ARGS__; } break;   \
    case CTP_L_CELL8: { constexpr int L = CTP_L_CELL8; __VA_ARGS__; } break;   \
    case CTP_L_TEX:   { constexpr int L = CTP_L_TEX;   __VA_ARGS__; } break;   \
    default:          { constexpr int L = CTP_L_BRICK; __VA_ARGS__; } break;   \
  }

// Per-query kernels that one CTA per query would leave on a handful of SMs (few queries of very many points) run on a
// thread-block cluster of 8 CTAs per query instead (ctp_prm.cuh: k_knn_setup, k_knn_scan, k_csr_rowptr).
#define CTP_PERQ_CL 8
static inline bool perq_clustered(const ctp_map *m, int64_t q, int64_t len) {
  if (m->perq_cluster >= 0) return m->perq_cluster != 0 && q * CTP_PERQ_CL < ((int64_t)1 << 30);
  return q * CTP_PERQ_CL <= 2 * (int64_t)m->sm_count && len >= 16384;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_perq_cluster(void (*kern)(KArgs...), int64_t q, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(q * CTP_PERQ_CL));
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CTP_PERQ_CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline int grid_for(int64_t items, int block, int cap_blocks) {
  int64_t g = (items + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap_blocks) g = cap_blocks;
  return (int)g;
}

// ------------------------------ point-wise entry points ------------------------------------
extern "C" int ctp_clearance_dev(ctp_map *m, const double *xyz, int64_t cnt, double *out, void *stream) {
  REQUIRE(m && cnt >= 0, "bad argument");
  if (cnt == 0) return CTP_OK;
  REQUIRE(xyz && out, "null buffer");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_L(m, k_clearance<L><<<grid_for(cnt, 256, m->sm_count * 8), 256, 0, st>>>(m->dev, xyz, cnt, out));
  m->launches++;
  CK(cudaGetLastError());
  return CTP_OK;
}

extern "C" int ctp_in_collision_dev(ctp_map *m, const double *xyz, int64_t cnt, double clr, uint8_t *out,
                                    void *stream) {
  REQUIRE(m && cnt >= 0, "bad argument");
  if (cnt == 0) return CTP_OK;
  REQUIRE(xyz && out, "null buffer");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_L(m, k_in_collision<L><<<grid_for(cnt, 256, m->sm_count * 8), 256, 0, st>>>(m->dev, xyz, cnt, clr, out));
  m->launches++;
  CK(cudaGetLastError());
  return CTP_OK;
}

extern "C" int ctp_gradient_dev(ctp_map *m, const double *xyz, int64_t cnt, double *grad, double *centre,
                                void *stream) {
  REQUIRE(m && cnt >= 0, "bad argument");
  if (cnt == 0) return CTP_OK;
  REQUIRE(xyz && grad && centre, "null buffer");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_L(m, k_gradient<L><<<grid_for(cnt, 256, m->sm_count * 8), 256, 0, st>>>(m->dev, xyz, cnt, grad, centre));
  m->launches++;
  CK(cudaGetLastError());
  return CTP_OK;
}

// host-pointer wrappers: stage through the io arena
// The small host-pointer calls (polylines: tens of kilobytes in a handful of arrays) used to issue one cudaMemcpyAsync
// per array from / to pageable memory -- ten staged copies of ~10 us each around a kernel of ~100 us.  StagedIn packs the
// input arrays into one page-locked block laid out like the device arena and uploads it with ONE copy; StagedOut brings
// the (contiguous) output takes back with ONE copy and hands the pieces out.
static int stage_reserve(ctp_map *m, size_t bytes) {
  if (bytes <= m->h_stage_cap) return CTP_OK;
  if (m->h_stage) {
    CK(cudaDeviceSynchronize());
    cudaFreeHost(m->h_stage);
  }
  m->h_stage = nullptr;
  m->h_stage_cap = 0;
  const size_t want = bytes + bytes / 2 + (1 << 16);
  CK(cudaHostAlloc((void **)&m->h_stage, want, cudaHostAllocDefault));
  m->h_stage_cap = want;
  return CTP_OK;
}
struct StagedIn {
  ctp_map *m;
  char *dev0 = nullptr;
  size_t host_off;  // where this block starts in the staging buffer
  size_t used = 0;
  StagedIn(ctp_map *mm, size_t at) : m(mm), host_off(at) {}
  template <typename T> T *put(const T *src, size_t count) {
    T *d = arena_take<T>(m->io, count);
    if (!dev0) dev0 = (char *)d;
    const size_t off = (size_t)((char *)d - dev0);
    memcpy(m->h_stage + host_off + off, src, count * sizeof(T));
    used = off + align_up(count * sizeof(T));
    return d;
  }
  cudaError_t upload(cudaStream_t st) const {
    return used ? cudaMemcpyAsync(dev0, m->h_stage + host_off, used, cudaMemcpyHostToDevice, st) : cudaSuccess;
  }
};
struct StagedOut {
  ctp_map *m;
  char *dev0 = nullptr;
  size_t host_off;
  size_t used = 0;
  StagedOut(ctp_map *mm, size_t at) : m(mm), host_off(at) {}
  template <typename T> T *take(size_t count) {
    T *d = arena_take<T>(m->io, count);
    if (!dev0) dev0 = (char *)d;
    used = (size_t)((char *)d - dev0) + align_up(count * sizeof(T));
    return d;
  }
  cudaError_t download(cudaStream_t st) const {
    return used ? cudaMemcpyAsync(m->h_stage + host_off, dev0, used, cudaMemcpyDeviceToHost, st) : cudaSuccess;
  }
  template <typename T> void get(const T *d, T *dst, size_t count) const {
    memcpy(dst, m->h_stage + host_off + (size_t)((const char *)d - dev0), count * sizeof(T));
  }
};

struct IoScope {
  ctp_map *m;
  explicit IoScope(ctp_map *mm) : m(mm) { m->io.used = 0; }
};

extern "C" int ctp_clearance(ctp_map *m, const double *xyz, int64_t cnt, double *out) {
  REQUIRE(m && cnt >= 0, "bad argument");
  if (cnt == 0) return CTP_OK;
  REQUIRE(xyz && out, "null buffer");
  CK(cudaSetDevice(m->device));
  CKR(arena_reserve(m->io, align_up(cnt * 24) + align_up(cnt * 8)));
  IoScope sc(m);
  double *d_x = arena_take<double>(m->io, cnt * 3), *d_o = arena_take<double>(m->io, cnt);
  CK(cudaMemcpyAsync(d_x, xyz, cnt * 24, cudaMemcpyHostToDevice, 0));
  CKR(ctp_clearance_dev(m, d_x, cnt, d_o, 0));
  CK(cudaMemcpyAsync(out, d_o, cnt * 8, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return CTP_OK;
}

extern "C" int ctp_in_collision(ctp_map *m, const double *xyz, int64_t cnt, double clr, uint8_t *out) {
  REQUIRE(m && cnt >= 0, "bad argument");
  if (cnt == 0) return CTP_OK;
  REQUIRE(xyz && out, "null buffer");
  CK(cudaSetDevice(m->device));
  CKR(arena_reserve(m->io, align_up(cnt * 24) + align_up(cnt)));
  IoScope sc(m);
  double *d_x = arena_take<double>(m->io, cnt * 3);
  uint8_t *d_o = arena_take<uint8_t>(m->io, cnt);
  CK(cudaMemcpyAsync(d_x, xyz, cnt * 24, cudaMemcpyHostToDevice, 0));
  CKR(ctp_in_collision_dev(m, d_x, cnt, clr, d_o, 0));
  CK(cudaMemcpyAsync(out, d_o, cnt, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return CTP_OK;
}

extern "C" int ctp_gradient(ctp_map *m, const double *xyz, int64_t cnt, double *grad, double *centre) {
  REQUIRE(m && cnt >= 0, "bad argument");
  if (cnt == 0) return CTP_OK;
  REQUIRE(xyz && grad && centre, "null buffer");
  CK(cudaSetDevice(m->device));
  CKR(arena_reserve(m->io, 3 * align_up(cnt * 24)));
  IoScope sc(m);
  double *d_x = arena_take<double>(m->io, cnt * 3), *d_g = arena_take<double>(m->io, cnt * 3),
         *d_c = arena_take<double>(m->io, cnt * 3);
  CK(cudaMemcpyAsync(d_x, xyz, cnt * 24, cudaMemcpyHostToDevice, 0));
  CKR(ctp_gradient_dev(m, d_x, cnt, d_g, d_c, 0));
  CK(cudaMemcpyAsync(grad, d_g, cnt * 24, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(centre, d_c, cnt * 24, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return CTP_OK;
}

// ----------------------------------- segment checks ----------------------------------------
// slot of (layout, lanes per edge) in ctp_map::occ_cache: layouts 1, 2, 4, 8 x groups 1, 4, 8, 16, 32 (+ one)
static inline int occ_slot(int layout, int group) {
  const int li = layout == CTP_L_PLAIN ? 0 : layout == CTP_L_CELL8 ? 1 : layout == CTP_L_TEX ? 2 : 3;
  const int gi = group == 1 ? 0 : group == 4 ? 1 : group == 8 ? 2 : group == 16 ? 3 : group == 2 ? 5 : 4;  // 2 = flat kernel, growing rounds
  return li * 6 + gi;
}
template <int L, int G>
static int launch_edges_LG(ctp_map *m, const EdgeSrc &S, int64_t cnt, double clr, double cdc, int guards,
                           const EdgeOut &O, cudaStream_t st) {
  int &occ = m->occ_cache[occ_slot(L, G)];  // resident CTAs per SM for this instantiation, on this handle's device
  if (occ == 0) {
    int o = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_edge_march<L, G>, 256, 0));
    occ = o > 0 ? o : 1;
  }
  const int64_t groups_per_block = 256 / G;
  int64_t blocks = (cnt + groups_per_block - 1) / groups_per_block;
  const int64_t cap = (int64_t)m->sm_count * occ;  // persistent: one resident wave
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_edge_march<L, G><<<(int)blocks, 256, 0, st>>>(m->dev, S, cnt, clr, cdc, guards, O);
  m->launches++;
  CK(cudaGetLastError());
  return CTP_OK;
}

template <int L>
static int launch_edges_flat(ctp_map *m, const EdgeSrc &S, int64_t cnt, double clr, double cdc, int guards,
                             const EdgeOut &O, cudaStream_t st) {
  const bool grow = !S.indexed;  // arbitrary segments: short first rounds; k-NN edges: one round of 16
  int &occ = m->occ_cache[occ_slot(L, grow ? 2 : 1)];
  if (occ == 0) {
    int o = 0;
    if (grow) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_edge_flat<L, true>, 32 * CTP_FLAT_WARPS, 0));
    else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_edge_flat<L, false>, 32 * CTP_FLAT_WARPS, 0));
    occ = o > 0 ? o : 1;
  }
  const int64_t per_block = CTP_FLAT_EPW * CTP_FLAT_WARPS;
  int64_t blocks = (cnt + per_block - 1) / per_block;
  const int64_t cap = (int64_t)m->sm_count * occ;  // persistent: one resident wave
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (grow) k_edge_flat<L, true><<<(int)blocks, 32 * CTP_FLAT_WARPS, 0, st>>>(m->dev, S, cnt, clr, cdc, guards, O);
  else k_edge_flat<L, false><<<(int)blocks, 32 * CTP_FLAT_WARPS, 0, st>>>(m->dev, S, cnt, clr, cdc, guards, O);
  m->launches++;
  CK(cudaGetLastError());
  return CTP_OK;
}

template <int L>
static int launch_edges_L(ctp_map *m, int group, const EdgeSrc &S, int64_t cnt, double clr, double cdc,
                          int guards, const EdgeOut &O, cudaStream_t st) {
  switch (group) {
    case 1: return launch_edges_flat<L>(m, S, cnt, clr, cdc, guards, O, st);
    case 4: return launch_edges_LG<L, 4>(m, S, cnt, clr, cdc, guards, O, st);
    case 8: return launch_edges_LG<L, 8>(m, S, cnt, clr, cdc, guards, O, st);
    case 16: return launch_edges_LG<L, 16>(m, S, cnt, clr, cdc, guards, O, st);
    default: return launch_edges_LG<L, 32>(m, S, cnt, clr, cdc, guards, O, st);
  }
}

static int launch_edges(ctp_map *m, int group, const EdgeSrc &S, int64_t cnt, double clr, double cdc, int guards,
                        const EdgeOut &O, cudaStream_t st) {
  DISPATCH_L(m, return launch_edges_L<L>(m, group, S, cnt, clr, cdc, guards, O, st));
  return CTP_OK;
}

// Scratch of the whole-edge filter: the node table (cell + fraction of every node's index coordinate), the list of
// undecided edges and its counter
struct EdgeScratch {
  float4 *qt = nullptr;
  int32_t *list = nullptr;   // edges the middle probe left undecided
  int32_t *count = nullptr;  // its length
};
static size_t ws_bytes_edge_filter(int64_t rows, int64_t edges) {
  return align_up((size_t)rows * 16) + align_up((size_t)edges * 4) + 256;
}
static EdgeScratch edge_scratch_take(ctp_map *m, int64_t rows, int64_t edges) {
  EdgeScratch s;
  s.qt = arena_take<float4>(m->ws, rows);
  s.list = arena_take<int32_t>(m->ws, edges);
  s.count = arena_take<int32_t>(m->ws, 64);
  return s;
}
// may the roadmap edges of this launch go through k_edge_mid first?
static bool edge_filter_applies(const ctp_map *m, const EdgeSrc &S, int64_t cnt, double clr, double cdc, int group) {
  return m->edge_filter && group == 1 && m->dev.filt_ok && m->n <= 1024 && S.indexed && !S.from && S.small && !S.perm &&
         cnt > 0 && cnt < ((int64_t)1 << 31) && clr == clr && fabs(clr) < 1e30 && cdc > 0 && cdc < 1e30;
}

// The n*k directed checks of a roadmap: the whole-edge float32 filter decides what it can with certainty from one
// lookup per edge (k_edge_mid), the exact flat kernel marches the rest from the list.  `rows` = nodes the table covers
// (all queries of the launch).
template <int L>
static int launch_edges_prm_L(ctp_map *m, EdgeSrc S, int64_t rows, int64_t cnt, double clr, double cdc, const EdgeOut &O,
                              const EdgeScratch &es, cudaStream_t st) {
  k_node_qtable<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(m->dev, S.a, rows, es.qt);
  CK(cudaMemsetAsync(es.count, 0, 64 * sizeof(int32_t), st));
  MidPar P;
  const double vox = 1.0 / (m->dev.s1 * m->dev.s2);
  P.hc = (float)(vox / cdc);
  P.f_need = (float)(clr + 2.0 * m->dev.filt_eps);
  if ((double)P.f_need < clr + 2.0 * m->dev.filt_eps) P.f_need = nextafterf(P.f_need, INFINITY);
  P.lip_g = nextafterf((float)((double)m->dev.filt_g * 1.001), INFINITY);
  P.eta = nextafterf((float)m->dev.filt_eta, INFINITY);
  const int64_t per_cta = CTP_MID_THREADS * CTP_MID_IPT;
  const unsigned ctas = (unsigned)((cnt + per_cta - 1) / per_cta);
  // middle probe for every edge; the exact kernel for the edges it leaves on the list.  (Measured and dropped: further
  // probes stepping along the undecided edges -- they decide 99 % of the C2 edges from a fifth of the lookups, but the
  // dependent, divergent probes cost as much as marching those edges exactly.)
  k_edge_mid<L><<<ctas, CTP_MID_THREADS, 0, st>>>(m->dev, S, (uint32_t)cnt, P, es.qt, O, es.list, es.count, m->d_lookups);
  m->launches += 2;
  CK(cudaGetLastError());
  // the first filtered launch for this (threshold, spacing) on this map: see what the middle probe decides; where it
  // is little (edges long against the clearance around them, or a grid with jumps) the filter costs more than it saves
  // and the following launches go straight to the exact kernel.  One synchronisation, once; verdicts do not depend on it.
  if (m->filter_cal_clr != clr || m->filter_cal_cdc != cdc) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap == cudaStreamCaptureStatusNone && cnt >= 4096) {
      int32_t left = 0;
      CK(cudaMemcpyAsync(&left, es.count, sizeof left, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      m->filter_cal_clr = clr;
      m->filter_cal_cdc = cdc;
      m->filter_pays = (double)left <= 0.65 * (double)cnt;
    }
  }
  S.perm = es.list;
  S.m_dev = es.count;
  S.order = nullptr;
  return launch_edges_flat<L>(m, S, cnt, clr, cdc, 0, O, st);
}
static int launch_edges_prm(ctp_map *m, int group, const EdgeSrc &S, int64_t rows, int64_t cnt, double clr, double cdc,
                            const EdgeOut &O, const EdgeScratch &es, cudaStream_t st) {
  const bool calibrated = m->filter_cal_clr == clr && m->filter_cal_cdc == cdc;
  if (!es.qt || !edge_filter_applies(m, S, cnt, clr, cdc, group) || (calibrated && !m->filter_pays && m->edge_filter != 2))
    return launch_edges(m, group, S, cnt, clr, cdc, 0, O, st);
  DISPATCH_L(m, return launch_edges_prm_L<L>(m, S, rows, cnt, clr, cdc, O, es, st));
  return CTP_OK;
}

static int check_edge_args(ctp_map *m, int64_t cnt, double cdc, int group) {
  REQUIRE(m && cnt >= 0, "bad argument");
  REQUIRE(cdc > 0, "collision_distance_check must be > 0 (topological_prm_clustering.hpp:270-275)");
  REQUIRE(group == 0 || group == 1 || group == 4 || group == 8 || group == 16 || group == 32,
          "group must be 0, 1, 4, 8, 16 or 32");
  return CTP_OK;
}

extern "C" int ctp_edge_check_dev(ctp_map *m, const double *a, const double *b, int64_t cnt, double clr,
                                  double cdc, int mode, int group, uint8_t *free_out, double *hit,
                                  int32_t *hit_idx, int32_t *lookups, void *stream) {
  CKR(check_edge_args(m, cnt, cdc, group));
  REQUIRE(mode == CTP_EDGE_NODES || mode == CTP_EDGE_GUARDS, "mode");
  if (cnt == 0) return CTP_OK;
  REQUIRE(a && b && free_out, "null buffer");
  CK(cudaSetDevice(m->device));
  EdgeSrc S{a, b, nullptr, nullptr, 0, 0, 0, 0, nullptr};
  EdgeOut O{free_out, hit, hit_idx, lookups, m->d_lookups};
  cudaStream_t st = (cudaStream_t)stream;
  // A large unordered batch on a grid beyond the L2 reads the one-sector CELL8 records where they exist (isolated
  // lookups pay for every sector a layout touches: 2^24 random segments on the 1024^3 grid 5.8 ms on the gather texture,
  // 4.6 ms on CELL8).  Binning the segments by the voxel block of their midpoint first (CTP_TUNE_SEG_BIN) is available
  // but off by default: measured on the same batch it does not pay (5.6 ms texture, 5.6 ms CELL8, 4.3 ms BRICK) -- the
  // kernel is bound by per-segment setup and wasted speculative samples, not by DRAM or TLB misses, and the
  // indirection un-coalesces the endpoint loads and the verdict stores.
  const bool big_grid = (size_t)m->n * m->n * m->n * 4 > ((size_t)100 << 20);
  const bool bin = m->seg_bin > 0;
  const int active_keep = m->active;
  const size_t keep = m->ws.used;
  if (bin && cnt < ((int64_t)1 << 31)) {
    const unsigned nb = (unsigned)((m->n + (1 << CTP_BIN_SHIFT) - 1) >> CTP_BIN_SHIFT);
    int bits = 0;
    while ((1u << bits) < nb) bits++;
    if (bits <= 8) {
      const int64_t nkeys = (int64_t)1 << (3 * bits);
      CKR(arena_reserve(m->ws, keep + 2 * align_up(cnt * 4) + 2 * align_up((nkeys + 1) * 4) + 4096));
      unsigned *d_key = arena_take<unsigned>(m->ws, cnt);
      int32_t *d_perm = arena_take<int32_t>(m->ws, cnt);
      int32_t *d_cnt = arena_take<int32_t>(m->ws, nkeys + 1), *d_start = arena_take<int32_t>(m->ws, nkeys + 1);
      m->ws.used = keep;  // dead once the kernels below have run (stream order)
      CK(cudaMemsetAsync(d_cnt, 0, (size_t)nkeys * 4, st));
      const int blocks = (int)((cnt + 255) / 256);
      k_seg_bin_count<<<blocks, 256, 0, st>>>(m->dev, a, b, cnt, nb, d_key, d_cnt);
      if (nkeys >= 16384) CK(launch_perq_cluster(k_excl_scan_i32<CTP_PERQ_CL>, 1, st, d_cnt, nkeys, d_start));
      else k_excl_scan_i32<1><<<1, 1024, 0, st>>>(d_cnt, nkeys, d_start);
      k_seg_bin_scatter<<<blocks, 256, 0, st>>>(d_key, cnt, d_start, d_cnt, d_perm);
      m->launches += 3;
      CK(cudaGetLastError());
      S.perm = d_perm;
    }
  }
  if (m->seg_bin < 0 && big_grid && cnt >= 65536 && m->reject_layout_auto && (m->layouts & CTP_L_CELL8)) m->active = CTP_L_CELL8;
  const int rc = launch_edges(m, group ? group : 16, S, cnt, clr, cdc, mode == CTP_EDGE_GUARDS, O, st);
  m->active = active_keep;
  return rc;
}

extern "C" int ctp_edge_check_indexed_dev(ctp_map *m, const double *nodes, int64_t n_nodes, const int32_t *from,
                                          const int32_t *to, int64_t cnt, double clr, double cdc, int group,
                                          uint8_t *free_out, int32_t *hit_idx, int32_t *lookups, void *stream) {
  CKR(check_edge_args(m, cnt, cdc, group));
  if (cnt == 0) return CTP_OK;
  REQUIRE(nodes && from && to && free_out && n_nodes > 0, "null buffer");
  CK(cudaSetDevice(m->device));
  EdgeSrc S{nodes, nullptr, from, to, n_nodes, 1, 1, 1, nullptr};
  EdgeOut O{free_out, nullptr, hit_idx, lookups, m->d_lookups};
  return launch_edges(m, group ? group : 1, S, cnt, clr, cdc, 0, O, (cudaStream_t)stream);
}

static int pipe_init(ctp_map *m);  // streams and events of the chunk pipelines (defined with them below)

static int auto_group(const double *a, const double *b, int64_t cnt, double cdc) {
  const int64_t probe = std::min<int64_t>(cnt, 4096);
  double sum = 0;
  for (int64_t i = 0; i < probe; i++) {
    const int64_t e = (cnt / probe) * i;
    const double dx = b[3 * e] - a[3 * e], dy = b[3 * e + 1] - a[3 * e + 1], dz = b[3 * e + 2] - a[3 * e + 2];
    const double d = sqrt(dx * dx + dy * dy + dz * dz);
    if (d == d && d < 1e300) sum += d;
  }
  const double samples = sum / (double)probe / cdc + 1.0;
  return samples <= 16 ? 1 : samples <= 24 ? 16 : 32;
}

extern "C" int ctp_edge_check(ctp_map *m, const double *a, const double *b, int64_t cnt, double clr, double cdc,
                              int mode, int group, uint8_t *free_out, double *hit, int32_t *hit_idx,
                              int32_t *lookups) {
  CKR(check_edge_args(m, cnt, cdc, group));
  if (cnt == 0) return CTP_OK;
  REQUIRE(a && b && free_out, "null buffer");
  CK(cudaSetDevice(m->device));
  if (group == 0) group = auto_group(a, b, cnt, cdc);
  const int64_t CH = (int64_t)1 << 20;  // segments per chunk of the copy / compute / copy pipeline
  if (cnt > 2 * CH) {
    // large batches: endpoints go up, verdicts come back and the kernel runs on three streams over double-buffered
    // chunks, so that with page-locked host buffers the call costs the slower of the two copy directions
    CKR(pipe_init(m));
    const size_t per = 2 * align_up(CH * 24) + align_up(CH) + align_up(CH * 24) + 2 * align_up(CH * 4);
    CKR(arena_reserve(m->io, 2 * per + 1024));
    IoScope sc(m);
    struct { double *a, *b, *h; uint8_t *f; int32_t *hi, *l; } B[2];
    for (int i = 0; i < 2; i++) {
      B[i].a = arena_take<double>(m->io, CH * 3);
      B[i].b = arena_take<double>(m->io, CH * 3);
      B[i].f = arena_take<uint8_t>(m->io, CH);
      B[i].h = hit ? arena_take<double>(m->io, CH * 3) : nullptr;
      B[i].hi = hit_idx ? arena_take<int32_t>(m->io, CH) : nullptr;
      B[i].l = lookups ? arena_take<int32_t>(m->io, CH) : nullptr;
    }
    cudaStream_t s_in = m->pipe_st[0], s_comp = m->pipe_st[1], s_out = m->pipe_st[2];
    const int64_t n_chunks = (cnt + CH - 1) / CH;
    for (int64_t c = 0; c < n_chunks; c++) {
      const int bi = (int)(c & 1);
      const int64_t e0 = c * CH, len = std::min(CH, cnt - e0);
      if (c >= 2) CK(cudaStreamWaitEvent(s_in, m->ev_comp[bi], 0));  // the kernel that read this buffer last
      CK(cudaMemcpyAsync(B[bi].a, a + 3 * e0, len * 24, cudaMemcpyHostToDevice, s_in));
      CK(cudaMemcpyAsync(B[bi].b, b + 3 * e0, len * 24, cudaMemcpyHostToDevice, s_in));
      CK(cudaEventRecord(m->ev_in[bi], s_in));
      CK(cudaStreamWaitEvent(s_comp, m->ev_in[bi], 0));
      if (c >= 2) CK(cudaStreamWaitEvent(s_comp, m->ev_out[bi], 0));  // its previous outputs have left
      CKR(ctp_edge_check_dev(m, B[bi].a, B[bi].b, len, clr, cdc, mode, group, B[bi].f, B[bi].h, B[bi].hi, B[bi].l, s_comp));
      CK(cudaEventRecord(m->ev_comp[bi], s_comp));
      CK(cudaStreamWaitEvent(s_out, m->ev_comp[bi], 0));
      CK(cudaMemcpyAsync(free_out + e0, B[bi].f, len, cudaMemcpyDeviceToHost, s_out));
      if (hit) CK(cudaMemcpyAsync(hit + 3 * e0, B[bi].h, len * 24, cudaMemcpyDeviceToHost, s_out));
      if (hit_idx) CK(cudaMemcpyAsync(hit_idx + e0, B[bi].hi, len * 4, cudaMemcpyDeviceToHost, s_out));
      if (lookups) CK(cudaMemcpyAsync(lookups + e0, B[bi].l, len * 4, cudaMemcpyDeviceToHost, s_out));
      CK(cudaEventRecord(m->ev_out[bi], s_out));
    }
    m->io_h2d += (unsigned long long)cnt * 48;
    m->io_d2h += (unsigned long long)cnt * (1 + (hit ? 24 : 0) + (hit_idx ? 4 : 0) + (lookups ? 4 : 0));
    CK(cudaStreamSynchronize(s_out));
    CK(cudaStreamSynchronize(s_comp));
    CK(cudaStreamSynchronize(s_in));
    return CTP_OK;
  }
  CKR(arena_reserve(m->io, 2 * align_up(cnt * 24) + align_up(cnt) + align_up(cnt * 24) + 2 * align_up(cnt * 4)));
  IoScope sc(m);
  double *d_a = arena_take<double>(m->io, cnt * 3), *d_b = arena_take<double>(m->io, cnt * 3);
  uint8_t *d_f = arena_take<uint8_t>(m->io, cnt);
  double *d_h = hit ? arena_take<double>(m->io, cnt * 3) : nullptr;
  int32_t *d_hi = hit_idx ? arena_take<int32_t>(m->io, cnt) : nullptr;
  int32_t *d_l = lookups ? arena_take<int32_t>(m->io, cnt) : nullptr;
  CK(cudaMemcpyAsync(d_a, a, cnt * 24, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_b, b, cnt * 24, cudaMemcpyHostToDevice, 0));
  CKR(ctp_edge_check_dev(m, d_a, d_b, cnt, clr, cdc, mode, group, d_f, d_h, d_hi, d_l, 0));
  CK(cudaMemcpyAsync(free_out, d_f, cnt, cudaMemcpyDeviceToHost, 0));
  if (hit) CK(cudaMemcpyAsync(hit, d_h, cnt * 24, cudaMemcpyDeviceToHost, 0));
  if (hit_idx) CK(cudaMemcpyAsync(hit_idx, d_hi, cnt * 4, cudaMemcpyDeviceToHost, 0));
  if (lookups) CK(cudaMemcpyAsync(lookups, d_l, cnt * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return CTP_OK;
}

extern "C" int ctp_edge_check_indexed(ctp_map *m, const double *nodes, int64_t n_nodes, const int32_t *from,
                                      const int32_t *to, int64_t cnt, double clr, double cdc, int group,
                                      uint8_t *free_out, int32_t *hit_idx, int32_t *lookups) {
  CKR(check_edge_args(m, cnt, cdc, group));
  if (cnt == 0) return CTP_OK;
  REQUIRE(nodes && from && to && free_out && n_nodes > 0, "null buffer");
  for (int64_t i = 0; i < cnt; i++)
    REQUIRE(from[i] >= 0 && from[i] < n_nodes && to[i] < n_nodes, "edge index out of range");
  CK(cudaSetDevice(m->device));
  CKR(arena_reserve(m->io, align_up(n_nodes * 24) + 4 * align_up(cnt * 4) + align_up(cnt)));
  IoScope sc(m);
  double *d_n = arena_take<double>(m->io, n_nodes * 3);
  int32_t *d_from = arena_take<int32_t>(m->io, cnt), *d_to = arena_take<int32_t>(m->io, cnt);
  uint8_t *d_f = arena_take<uint8_t>(m->io, cnt);
  int32_t *d_hi = hit_idx ? arena_take<int32_t>(m->io, cnt) : nullptr;
  int32_t *d_l = lookups ? arena_take<int32_t>(m->io, cnt) : nullptr;
  CK(cudaMemcpyAsync(d_n, nodes, n_nodes * 24, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_from, from, cnt * 4, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_to, to, cnt * 4, cudaMemcpyHostToDevice, 0));
  CKR(ctp_edge_check_indexed_dev(m, d_n, n_nodes, d_from, d_to, cnt, clr, cdc, group, d_f, d_hi, d_l, 0));
  CK(cudaMemcpyAsync(free_out, d_f, cnt, cudaMemcpyDeviceToHost, 0));
  if (hit_idx) CK(cudaMemcpyAsync(hit_idx, d_hi, cnt * 4, cudaMemcpyDeviceToHost, 0));
  if (lookups) CK(cudaMemcpyAsync(lookups, d_l, cnt * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return CTP_OK;
}

// ------------------------------- workspace sizing -------------------------------------------
static int knn_max_cells(int64_t n) {
  int64_t c = 2 * n + 64;
  if (c > (1 << 22)) c = 1 << 22;
  return (int)c;
}
static size_t ws_bytes_reject(int64_t q, int64_t m) {
  const int64_t nblk = (m + CTP_REJ_BLOCK - 1) / CTP_REJ_BLOCK;
  return align_up((size_t)q * m) + align_up((size_t)q * nblk * 4);
}
static size_t ws_bytes_knn(int64_t q, int64_t n) {
  const size_t mc = (size_t)knn_max_cells(n);
  return align_up((size_t)q * sizeof(KnnGrid)) + 2 * align_up((size_t)q * n * 16) + 3 * align_up((size_t)q * n * 4) + align_up((size_t)q * n) +
         align_up((size_t)q * mc * 4) + align_up((size_t)q * (mc + 1) * 4) + 512;
}
static size_t ws_bytes_csr(int64_t q, int64_t n);
static size_t ws_bytes_prm(int64_t q, int64_t n, int k) {
  return ws_bytes_knn(q, n) + align_up((size_t)q * n * csr_ns(k) * 4) + align_up((size_t)q * n * k) + ws_bytes_csr(q, n) +
         ws_bytes_edge_filter(q * n, q * n * k);
}

extern "C" int ctp_reserve(ctp_map *m, int64_t q, int64_t n, int64_t cand, int k) {
  REQUIRE(m && q >= 1 && n >= 2 && cand >= 0 && k >= 1 && k <= CTP_KNN_MAXK, "bad argument");
  CK(cudaSetDevice(m->device));
  const size_t need = ws_bytes_reject(q, cand) + ws_bytes_prm(q, n, k) + align_up((size_t)q * n * 24) + (1 << 16);
  return arena_reserve(m->ws, need);
}

// ------------------------------- rejection sampling -----------------------------------------
// rejection over the first cnt candidates of each stream; streams are `stride` candidates apart
static int sample_reject_launch(ctp_map *m, const double *start, const double *goal, const double *cand, int64_t q,
                                int64_t cnt, int64_t stride, double clr, int64_t n, double *nodes, int32_t *n_filled,
                                int32_t *n_used, uint8_t *accept, cudaStream_t st, int32_t *src_idx = nullptr) {
  const int nblk = (int)std::max<int64_t>(1, (cnt + CTP_REJ_BLOCK - 1) / CTP_REJ_BLOCK);
  const size_t keep = m->ws.used;
  CKR(arena_reserve(m->ws, keep + ws_bytes_reject(q, cnt) + 512));
  uint8_t *d_acc = accept ? accept : arena_take<uint8_t>(m->ws, (size_t)q * cnt + 1);
  int32_t *d_bc = arena_take<int32_t>(m->ws, (size_t)q * nblk);
  m->ws.used = keep;  // scratch is dead once the three kernels below have run (stream order)
  dim3 grid((unsigned)nblk, (unsigned)q);
  StageSpan span(m, CTP_STAGE_REJECT, st);
  // Candidates are scattered over the whole ellipsoid: one isolated lookup each.  When the grid does not fit the L2
  // and the one-sector-per-lookup records (CELL8, 8x the memory) were materialised, the rejection reads those -- one
  // DRAM sector per candidate instead of the four of two texture gathers (C3: 0.38 -> 0.22 ms) -- while the edge
  // marches, whose consecutive samples share sectors, stay on the active layout.
  const bool big = (size_t)m->n * m->n * m->n * 4 > ((size_t)100 << 20);
  if (big && (m->layouts & CTP_L_CELL8) && m->reject_layout_auto) {
    k_reject_flags<CTP_L_CELL8><<<grid, CTP_REJ_BLOCK, 0, st>>>(m->dev, cand, cnt, stride, clr, d_acc, d_bc, nblk);
  } else {
    DISPATCH_L(m, k_reject_flags<L><<<grid, CTP_REJ_BLOCK, 0, st>>>(m->dev, cand, cnt, stride, clr, d_acc, d_bc, nblk));
  }
  m->launches++;
  CK(cudaGetLastError());
  k_reject_scan<<<(unsigned)q, 1024, 0, st>>>(d_bc, nblk, n, n_filled);
  m->launches++;
  CK(cudaGetLastError());
  // n_used defaults to "whole stream consumed"; the scatter pass overwrites it when n-2 were found
  k_fill_i32<<<grid_for(q, 256, 1024), 256, 0, st>>>(n_used, q, n == 2 ? 0 : (int32_t)cnt);  // n == 2: the loop of :309 never runs
  m->launches++;
  k_reject_scatter<<<grid, CTP_REJ_BLOCK, 0, st>>>(cand, start, goal, d_acc, d_bc, nblk, cnt, stride, n, nodes, n_used, src_idx);
  m->launches++;
  CK(cudaGetLastError());
  return CTP_OK;
}

extern "C" int ctp_sample_reject_dev(ctp_map *m, const double *start, const double *goal, const double *cand,
                                     int64_t q, int64_t cnt, double clr, int64_t n, double *nodes,
                                     int32_t *n_filled, int32_t *n_used, uint8_t *accept, void *stream) {
  REQUIRE(m && q >= 1 && cnt >= 0 && n >= 2, "bad argument");
  REQUIRE(start && goal && nodes && n_filled && n_used && (cand || cnt == 0), "null buffer");
  REQUIRE(q <= 65535, "at most 65535 queries per call");
  CK(cudaSetDevice(m->device));
  return sample_reject_launch(m, start, goal, cand, q, cnt, cnt, clr, n, nodes, n_filled, n_used, accept,
                              (cudaStream_t)stream);
}

extern "C" int ctp_sample_reject(ctp_map *m, const double *start, const double *goal, const double *cand,
                                 int64_t q, int64_t cnt, double clr, int64_t n, double *nodes, int32_t *n_filled,
                                 int32_t *n_used, uint8_t *accept) {
  REQUIRE(m && q >= 1 && cnt >= 0 && n >= 2, "bad argument");
  REQUIRE(start && goal && nodes && n_filled && n_used && (cand || cnt == 0), "null buffer");
  CK(cudaSetDevice(m->device));
  CKR(arena_reserve(m->io, 2 * align_up(q * 24) + align_up(q * cnt * 24) + align_up(q * n * 24) +
                               2 * align_up(q * 4) + align_up(q * cnt + 1)));
  IoScope sc(m);
  double *d_s = arena_take<double>(m->io, q * 3), *d_g = arena_take<double>(m->io, q * 3);
  double *d_c = arena_take<double>(m->io, q * cnt * 3), *d_n = arena_take<double>(m->io, q * n * 3);
  int32_t *d_nf = arena_take<int32_t>(m->io, q), *d_nu = arena_take<int32_t>(m->io, q);
  uint8_t *d_acc = arena_take<uint8_t>(m->io, q * cnt + 1);
  CK(cudaMemcpyAsync(d_s, start, q * 24, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_g, goal, q * 24, cudaMemcpyHostToDevice, 0));
  if (cnt) CK(cudaMemcpyAsync(d_c, cand, q * cnt * 24, cudaMemcpyHostToDevice, 0));
  CK(cudaMemsetAsync(d_n, 0, q * n * 24, 0));
  m->ws.used = 0;
  CKR(ctp_sample_reject_dev(m, d_s, d_g, d_c, q, cnt, clr, n, d_n, d_nf, d_nu, d_acc, 0));
  CK(cudaMemcpyAsync(nodes, d_n, q * n * 24, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(n_filled, d_nf, q * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(n_used, d_nu, q * 4, cudaMemcpyDeviceToHost, 0));
  if (accept && cnt) CK(cudaMemcpyAsync(accept, d_acc, q * cnt, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  for (int64_t i = 0; i < q; i++)
    if (n_filled[i] < n) return fail(CTP_ERR_NEED_CANDIDATES, "query %lld: only %d of %lld nodes accepted", (long long)i, n_filled[i], (long long)n);
  return CTP_OK;
}

// ------------------------------------ candidate sampling -----------------------------------
extern "C" int ctp_sample_ellipsoid_dev(ctp_map *m, const double *start, const double *goal, int64_t q,
                                        int64_t cnt, double ratio, int planar, uint64_t seed, double *cand,
                                        void *stream) {
  REQUIRE(m && q >= 1 && cnt >= 0 && ratio > 1.0, "bad argument (ratio must be > 1)");
  if (cnt == 0) return CTP_OK;
  REQUIRE(start && goal && cand, "null buffer");
  CK(cudaSetDevice(m->device));
  k_sample_ellipsoid<<<grid_for(q * cnt, 256, m->sm_count * 16), 256, 0, (cudaStream_t)stream>>>(
      start, goal, q, cnt, ratio, planar, seed, cand);
  m->launches++;
  CK(cudaGetLastError());
  return CTP_OK;
}

extern "C" int ctp_sample_ellipsoid(ctp_map *m, const double *start, const double *goal, int64_t q, int64_t cnt,
                                    double ratio, int planar, uint64_t seed, double *cand) {
  REQUIRE(m && q >= 1 && cnt >= 0, "bad argument");
  if (cnt == 0) return CTP_OK;
  REQUIRE(start && goal && cand, "null buffer");
  CK(cudaSetDevice(m->device));
  CKR(arena_reserve(m->io, 2 * align_up(q * 24) + align_up(q * cnt * 24)));
  IoScope sc(m);
  double *d_s = arena_take<double>(m->io, q * 3), *d_g = arena_take<double>(m->io, q * 3);
  double *d_c = arena_take<double>(m->io, q * cnt * 3);
  CK(cudaMemcpyAsync(d_s, start, q * 24, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_g, goal, q * 24, cudaMemcpyHostToDevice, 0));
  CKR(ctp_sample_ellipsoid_dev(m, d_s, d_g, q, cnt, ratio, planar, seed, d_c, 0));
  CK(cudaMemcpyAsync(cand, d_c, q * cnt * 24, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return CTP_OK;
}

// the same candidates together with the draws they were made from (draws: q x cnt x 4 = u, g0, g1, g2), so a test can
// push the very same draws through the CPU restatement of random_ellipsoid_point
extern "C" int ctp_sample_ellipsoid_draws(ctp_map *m, const double *start, const double *goal, int64_t q, int64_t cnt,
                                          double ratio, int planar, uint64_t seed, double *cand, double *draws) {
  REQUIRE(m && q >= 1 && cnt >= 0 && ratio > 1.0, "bad argument (ratio must be > 1)");
  if (cnt == 0) return CTP_OK;
  REQUIRE(start && goal && cand && draws, "null buffer");
  CK(cudaSetDevice(m->device));
  CKR(arena_reserve(m->io, 2 * align_up(q * 24) + align_up(q * cnt * 24) + align_up(q * cnt * 32)));
  IoScope sc(m);
  double *d_s = arena_take<double>(m->io, q * 3), *d_g = arena_take<double>(m->io, q * 3);
  double *d_c = arena_take<double>(m->io, q * cnt * 3), *d_d = arena_take<double>(m->io, q * cnt * 4);
  CK(cudaMemcpyAsync(d_s, start, q * 24, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_g, goal, q * 24, cudaMemcpyHostToDevice, 0));
  k_sample_ellipsoid<<<grid_for(q * cnt, 256, m->sm_count * 16), 256>>>(d_s, d_g, q, cnt, ratio, planar, seed, d_c, d_d);
  m->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(cand, d_c, q * cnt * 24, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(draws, d_d, q * cnt * 32, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return CTP_OK;
}

// --------------------------------------- k-NN ----------------------------------------------
// slice (optional, q == 1): {first, last} positions of the cell-sorted order on entry, moved to cell boundaries on
// return (one 16-byte read back); only the rows of the points at those positions are searched and written
static int knn_launch(ctp_map *m, const double *nodes, int64_t q, int64_t n, int k, int32_t *idx, int idx_stride,
                      float *d2, cudaStream_t st, const float4 **sorted_out = nullptr, int64_t *slice = nullptr) {
  const int mc = knn_max_cells(n);
  const int64_t total = q * n;
  KnnGrid *d_grid = arena_take<KnnGrid>(m->ws, q);
  float4 *d_pts = arena_take<float4>(m->ws, total), *d_sorted = arena_take<float4>(m->ws, total);
  int32_t *d_cell_of = arena_take<int32_t>(m->ws, total);
  int32_t *d_cnt = arena_take<int32_t>(m->ws, (size_t)q * mc);
  int32_t *d_start = arena_take<int32_t>(m->ws, (size_t)q * (mc + 1));
  if (sorted_out) *sorted_out = d_sorted;
  int32_t *d_fb = arena_take<int32_t>(m->ws, total), *d_fb2 = arena_take<int32_t>(m->ws, total);
  int32_t *d_fbn = arena_take<int32_t>(m->ws, 2);  // [0] ring-search list, [1] direct-search list
  uint8_t *d_fbk = arena_take<uint8_t>(m->ws, total);
  CK(cudaMemsetAsync(d_cnt, 0, (size_t)q * mc * 4, st));
  const bool clustered = perq_clustered(m, q, n);
  if (clustered) CK(launch_perq_cluster(k_knn_setup<CTP_PERQ_CL>, q, st, nodes, n, mc, m->knn_per_cell, d_grid, d_pts));
  else k_knn_setup<1><<<(unsigned)q, 1024, 0, st>>>(nodes, n, mc, m->knn_per_cell, d_grid, d_pts);
  m->launches++;
  CK(cudaGetLastError());
  const int blocks = (int)((total + 255) / 256);
  k_knn_count<<<blocks, 256, 0, st>>>(d_pts, n, total, d_grid, mc, d_cell_of, d_cnt);
  m->launches++;
  if (clustered) CK(launch_perq_cluster(k_knn_scan<CTP_PERQ_CL>, q, st, d_grid, mc, d_cnt, d_start));
  else k_knn_scan<1><<<(unsigned)q, 1024, 0, st>>>(d_grid, mc, d_cnt, d_start);
  m->launches++;
  k_knn_scatter<<<blocks, 256, 0, st>>>(d_pts, n, total, mc, d_cell_of, d_start, d_cnt, d_sorted);
  m->launches++;
  CK(cudaGetLastError());
  int64_t first = 0, last = total;
  if (slice) {
    int64_t *d_bounds = reinterpret_cast<int64_t *>(d_fb);  // free until the search kernels run
    k_knn_slice_bounds<<<1, 32, 0, st>>>(d_grid, d_start, slice[0], slice[1], d_bounds);
    m->launches++;
    CK(cudaMemcpyAsync(slice, d_bounds, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    first = slice[0];
    last = slice[1];
    if (last <= first) return CTP_OK;
  }
  if (m->knn_tile && n > 4 * k) {
    CK(cudaMemsetAsync(d_fbn, 0, 8, st));
    const int lblocks = m->sm_count * 8;
#define KNN_CASE(K)                                                                                               \
  case K:                                                                                                         \
    if (!(m->knn_smem_set & (1u << K))) {                                                                         \
      CK(cudaFuncSetAttribute(k_knn_direct<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, CTP_KNN_DIRECT_SMEM)); \
      CK(cudaFuncSetAttribute(k_knn_direct_list<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,                  \
                              CTP_KNN_DIRECT_SMEM));                                                              \
      m->knn_smem_set |= 1u << K;                                                                                 \
    }                                                                                                             \
    k_knn_direct<K><<<(int)((last - first + CTP_KNN_DIRECT_THREADS - 1) / CTP_KNN_DIRECT_THREADS),                \
                      CTP_KNN_DIRECT_THREADS, CTP_KNN_DIRECT_SMEM, st>>>(d_sorted, n, last, d_grid, mc, d_start,   \
                                                                         idx, idx_stride, d2, d_fb2, d_fbk,       \
                                                                         d_fbn + 1, first);                       \
    k_knn_direct_list<K><<<lblocks, CTP_KNN_DIRECT_THREADS, CTP_KNN_DIRECT_SMEM, st>>>(                           \
        d_sorted, n, d_fb2, d_fbk, d_fbn + 1, 1, d_grid, mc, d_start, idx, idx_stride, d2, d_fb, d_fbn);          \
    k_knn_ring_warp<K><<<lblocks, 32 * CTP_KNN_RW_WARPS, 0, st>>>(d_sorted, n, d_fb, d_fbn, d_grid, mc, d_start,  \
                                                                  idx, idx_stride, d2);                          \
    m->launches += 3;                                                                                             \
    break;
    switch (k) {
      KNN_CASE(1) KNN_CASE(2) KNN_CASE(3) KNN_CASE(4) KNN_CASE(5) KNN_CASE(6) KNN_CASE(7) KNN_CASE(8)
      KNN_CASE(9) KNN_CASE(10) KNN_CASE(11) KNN_CASE(12) KNN_CASE(13) KNN_CASE(14) KNN_CASE(15) KNN_CASE(16)
    }
#undef KNN_CASE
    if (getenv("CTP_DEBUG_KNN")) {
      int fb[2] = {0, 0};
      cudaMemcpyAsync(fb, d_fbn, 8, cudaMemcpyDeviceToHost, st);
      cudaStreamSynchronize(st);
      fprintf(stderr, "[ctp] knn: %lld points, %d redone by the direct search (%.2f%%), %d by the ring search (%.2f%%)\n",
              (long long)total, fb[1], 100.0 * fb[1] / (double)total, fb[0], 100.0 * fb[0] / (double)total);
    }
  } else {
    const int sblocks = (int)((last - first + 127) / 128);
#define KNN_CASE(K)                                                                                                \
  case K:                                                                                                          \
    k_knn_search<K><<<sblocks, 128, 0, st>>>(d_sorted, n, last, d_grid, mc, d_start, idx, idx_stride, d2, first); \
    m->launches++;                                                                                             \
    break;
    switch (k) {
      KNN_CASE(1) KNN_CASE(2) KNN_CASE(3) KNN_CASE(4) KNN_CASE(5) KNN_CASE(6) KNN_CASE(7) KNN_CASE(8)
      KNN_CASE(9) KNN_CASE(10) KNN_CASE(11) KNN_CASE(12) KNN_CASE(13) KNN_CASE(14) KNN_CASE(15) KNN_CASE(16)
    }
#undef KNN_CASE
  }
  if (m->knn_radius > 0) {
    const float r2 = (float)(m->knn_radius * m->knn_radius);
    k_knn_radius_mask<<<(unsigned)((total * k + 255) / 256), 256, 0, st>>>(nodes, n, total, k, r2, idx, idx_stride, d2);
    m->launches++;
  }
  CK(cudaGetLastError());
  return CTP_OK;
}

extern "C" int ctp_knn_dev(ctp_map *m, const double *nodes, int64_t q, int64_t n, int k, int32_t *idx, float *d2,
                           void *stream) {
  REQUIRE(m && q >= 1 && n >= 1 && k >= 1 && k <= CTP_KNN_MAXK, "bad argument (k must be in [1,16])");
  REQUIRE(nodes && idx, "null buffer");
  REQUIRE(q * n < ((int64_t)1 << 31), "q*n must fit in int32");
  CK(cudaSetDevice(m->device));
  const size_t keep = m->ws.used;
  CKR(arena_reserve(m->ws, keep + ws_bytes_knn(q, n) + 4096));
  int rc = knn_launch(m, nodes, q, n, k, idx, k, d2, (cudaStream_t)stream);
  m->ws.used = keep;
  return rc;
}

extern "C" int ctp_knn(ctp_map *m, const double *nodes, int64_t q, int64_t n, int k, int32_t *idx, float *d2) {
  REQUIRE(m && q >= 1 && n >= 1 && k >= 1 && k <= CTP_KNN_MAXK, "bad argument (k must be in [1,16])");
  REQUIRE(nodes && idx, "null buffer");
  CK(cudaSetDevice(m->device));
  CKR(arena_reserve(m->io, align_up(q * n * 24) + 2 * align_up(q * n * k * 4)));
  IoScope sc(m);
  double *d_n = arena_take<double>(m->io, q * n * 3);
  int32_t *d_i = arena_take<int32_t>(m->io, q * n * k);
  float *d_d = d2 ? arena_take<float>(m->io, q * n * k) : nullptr;
  CK(cudaMemcpyAsync(d_n, nodes, q * n * 24, cudaMemcpyHostToDevice, 0));
  m->ws.used = 0;
  CKR(ctp_knn_dev(m, d_n, q, n, k, d_i, d_d, 0));
  CK(cudaMemcpyAsync(idx, d_i, q * n * k * 4, cudaMemcpyDeviceToHost, 0));
  if (d2) CK(cudaMemcpyAsync(d2, d_d, q * n * k * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return CTP_OK;
}

// ------------------------------------- PRM build -------------------------------------------
// symmetric union of the free directed edges -> CSR (:374-385); rec holds the neighbour ids
struct CsrScratch {
  int32_t *own;     // entries a row contributes to itself
  int32_t *xcnt;    // entries other rows push into it
  int64_t *nnz;
  uint16_t *rec16;  // null when the ids do not fit 16 bits
  void *xtra;       // rows x CTP_CSR_XCAP parked ids (16-bit with rec16, else 32-bit)
  int32_t *big, *big_cnt;
  int2 *big_pairs;  // (row, id) of the extra entries that did not fit a row's staging line
  int pair_cap;
  int32_t *big_tmp; // one line of n + 32 ids per CTA of k_csr_emit_big
  int big_ctas;
};
static size_t ws_bytes_csr(int64_t q, int64_t n) {
  const size_t rows = (size_t)q * n;
  return 3 * align_up(rows * 4) + align_up((size_t)q * 8) + align_up(rows * 32) + align_up(rows * CTP_CSR_XCAP * 4) +
         align_up(rows * 8) + align_up((size_t)256 * (size_t)(n + 32) * 4) + 2048;
}
static CsrScratch csr_scratch_take(ctp_map *m, int64_t q, int64_t n, int k) {
  Arena &a = m->ws;
  const size_t rows = (size_t)q * n;
  CsrScratch s;
  s.own = arena_take<int32_t>(a, rows);
  s.xcnt = arena_take<int32_t>(a, rows);
  s.nnz = arena_take<int64_t>(a, q);
  const bool small = n <= 65535 && k <= 14;
  s.rec16 = small ? arena_take<uint16_t>(a, rows * 16) : nullptr;
  s.xtra = small ? (void *)arena_take<uint16_t>(a, rows * CTP_CSR_XCAP) : (void *)arena_take<int32_t>(a, rows * CTP_CSR_XCAP);
  s.big = arena_take<int32_t>(a, rows);
  s.big_cnt = arena_take<int32_t>(a, 2);  // [0] rows on the big list, [1] overflow pairs
  s.pair_cap = (int)std::min<size_t>(rows, (size_t)1 << 26);
  if (m->csr_pair_cap > 0) s.pair_cap = (int)std::min<int64_t>(s.pair_cap, m->csr_pair_cap);
  s.big_pairs = arena_take<int2>(a, (size_t)s.pair_cap);
  s.big_ctas = std::max(1, std::min(m->sm_count, 256));
  s.big_tmp = arena_take<int32_t>(a, (size_t)s.big_ctas * (size_t)(n + 32));
  return s;
}

// csr_flags: CTP_CSR_UPPER = only entries with col > row; CTP_CSR_NO_WEIGHTS = w is not written (may be null)
static int csr_emit(ctp_map *m, const double *nodes, int64_t q, int64_t n, int k, int32_t *d_rec, const uint8_t *d_free,
                    const CsrScratch &sc, int csr_flags, int32_t *row_ptr, int32_t *col, double *w, int64_t cap,
                    int64_t *nnz_off, cudaStream_t st) {
  const int64_t rows = q * n, edges = rows * k;
  const int ns = csr_ns(k);
  const int upper = (csr_flags & CTP_CSR_UPPER) ? 1 : 0;
  if (csr_flags & CTP_CSR_NO_WEIGHTS) w = nullptr;
  StageSpan span(m, CTP_STAGE_CSR, st);
  const int rb = (int)((rows + 255) / 256);
  CK(cudaMemsetAsync(sc.big_cnt, 0, 8, st));
  CK(cudaMemsetAsync(sc.xcnt, 0, (size_t)rows * 4, st));
  CsrSplit sp;
  sp.small = edges < ((int64_t)1 << 31) && n < ((int64_t)1 << 31);
  sp.dk = ctp_div_make((uint32_t)k);
  sp.dn = ctp_div_make(sp.small ? (uint32_t)n : 1u);
  k_csr_rowmask<<<rb, 256, 0, st>>>(d_free, n, k, ns, rows, sp, upper, d_rec, sc.own, sc.rec16);
  m->launches++;
  if (sc.rec16)
    k_csr_count<16, uint16_t, true><<<rb, 256, 0, st>>>(d_rec, sc.rec16, n, k, rows, sp, upper, sc.xcnt, (uint16_t *)sc.xtra, sc.big, sc.big_cnt, sc.big_pairs, sc.pair_cap);
  else if (ns == 16)
    k_csr_count<16, int32_t, false><<<rb, 256, 0, st>>>(d_rec, nullptr, n, k, rows, sp, upper, sc.xcnt, (int32_t *)sc.xtra, sc.big, sc.big_cnt, sc.big_pairs, sc.pair_cap);
  else
    k_csr_count<32, int32_t, false><<<rb, 256, 0, st>>>(d_rec, nullptr, n, k, rows, sp, upper, sc.xcnt, (int32_t *)sc.xtra, sc.big, sc.big_cnt, sc.big_pairs, sc.pair_cap);
  m->launches++;
  if (perq_clustered(m, q, n)) CK(launch_perq_cluster(k_csr_rowptr<CTP_PERQ_CL>, q, st, sc.own, sc.xcnt, n, row_ptr, sc.nnz));
  else k_csr_rowptr<1><<<(unsigned)q, 1024, 0, st>>>(sc.own, sc.xcnt, n, row_ptr, sc.nnz);
  m->launches++;
  k_csr_query_offsets<<<1, 1024, 0, st>>>(sc.nnz, q, nnz_off);
  m->launches++;
  const int64_t tcap = std::min<int64_t>(cap, 2 * edges);
  const int eb = (int)((rows + CTP_CSR_EMIT_ROWS - 1) / CTP_CSR_EMIT_ROWS);
  if (sc.rec16)
    k_csr_emit<16, uint16_t><<<eb, CTP_CSR_EMIT_ROWS, 0, st>>>(nodes, d_rec, sc.rec16, (const uint16_t *)sc.xtra, n, k, rows, sp, upper, row_ptr, nnz_off, tcap, col, w);
  else if (ns == 16)
    k_csr_emit<16, int32_t><<<eb, CTP_CSR_EMIT_ROWS, 0, st>>>(nodes, d_rec, nullptr, (const int32_t *)sc.xtra, n, k, rows, sp, upper, row_ptr, nnz_off, tcap, col, w);
  else
    k_csr_emit<32, int32_t><<<eb, CTP_CSR_EMIT_ROWS, 0, st>>>(nodes, d_rec, nullptr, (const int32_t *)sc.xtra, n, k, rows, sp, upper, row_ptr, nnz_off, tcap, col, w);
  m->launches++;
  if (sc.rec16)
    k_csr_emit_big<16, uint16_t><<<sc.big_ctas, 256, 0, st>>>(nodes, d_rec, n, k, sp, upper, sc.big, sc.big_cnt, (const uint16_t *)sc.xtra, sc.big_pairs, sc.pair_cap, row_ptr, nnz_off, tcap, sc.big_tmp, col, w);
  else if (ns == 16)
    k_csr_emit_big<16, int32_t><<<sc.big_ctas, 256, 0, st>>>(nodes, d_rec, n, k, sp, upper, sc.big, sc.big_cnt, (const int32_t *)sc.xtra, sc.big_pairs, sc.pair_cap, row_ptr, nnz_off, tcap, sc.big_tmp, col, w);
  else
    k_csr_emit_big<32, int32_t><<<sc.big_ctas, 256, 0, st>>>(nodes, d_rec, n, k, sp, upper, sc.big, sc.big_cnt, (const int32_t *)sc.xtra, sc.big_pairs, sc.pair_cap, row_ptr, nnz_off, tcap, sc.big_tmp, col, w);
  m->launches++;
  CK(cudaGetLastError());
  return CTP_OK;
}

extern "C" int ctp_build_prm_ex_dev(ctp_map *m, const double *nodes, int64_t q, int64_t n, int k, double clr,
                                    double cdc, int csr_flags, int32_t *nbr, uint8_t *directed_free, int32_t *row_ptr,
                                    int32_t *col, double *w, int64_t cap, int64_t *nnz_off, void *stream) {
  REQUIRE(m && q >= 1 && n >= 2 && k >= 1 && k <= CTP_KNN_MAXK, "bad argument (k must be in [1,16])");
  REQUIRE((csr_flags & ~(CTP_CSR_UPPER | CTP_CSR_NO_WEIGHTS)) == 0, "unknown csr_flags bits");
  REQUIRE(nodes && row_ptr && col && (w || (csr_flags & CTP_CSR_NO_WEIGHTS)) && nnz_off && cap >= 0, "null buffer");
  REQUIRE(cdc > 0, "collision_distance_check must be > 0");
  REQUIRE(q * n * k < ((int64_t)1 << 31), "q*n*k must fit in int32");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t keep = m->ws.used;
  CKR(arena_reserve(m->ws, keep + ws_bytes_prm(q, n, k) + 8192));
  const int64_t rows = q * n, edges = rows * k;
  const int ns = csr_ns(k);
  int32_t *d_rec = arena_take<int32_t>(m->ws, rows * ns);  // per node: k neighbour ids + free-edge mask
  uint8_t *d_free = directed_free ? directed_free : arena_take<uint8_t>(m->ws, edges);
  const CsrScratch csc = csr_scratch_take(m, q, n, k);
  const EdgeScratch esc = edge_scratch_take(m, rows, edges);
  const float4 *d_sorted_keep = nullptr;
  int rc;
  {
    StageSpan span(m, CTP_STAGE_KNN, st);
    rc = knn_launch(m, nodes, q, n, k, d_rec, ns, nullptr, st, &d_sorted_keep);
  }
  // the k-NN scratch is dead except for the cell-sorted points, which the edge kernel reads; nothing
  // below takes from the arena again, so the block stays intact until this call returns
  m->ws.used = keep;
  CKR(rc);
  // the n*k directed checks (:355-397)
  // the flat kernel can walk the rows in the k-NN stage's cell-sorted order (spatial locality of the gathers)
  const bool use_order = m->edge_order >= 0 ? m->edge_order != 0
                                            : ((size_t)m->n * m->n * m->n * 4 > ((size_t)100 << 20) || m->active != CTP_L_TEX);
  EdgeSrc S{nodes, nullptr, nullptr, d_rec, n, k, 1, ns, (m->prm_group == 1 && use_order) ? d_sorted_keep : nullptr};
  S.small = edges < ((int64_t)1 << 31) && n < ((int64_t)1 << 31);
  S.dk = ctp_div_make((uint32_t)k);
  S.dn = ctp_div_make(S.small ? (uint32_t)n : 1u);
  EdgeOut O{d_free, nullptr, nullptr, nullptr, m->d_lookups};
  {
    StageSpan span(m, CTP_STAGE_EDGES, st);
    CKR(launch_edges_prm(m, m->prm_group, S, rows, edges, clr, cdc, O, esc, st));
  }
  CKR(csr_emit(m, nodes, q, n, k, d_rec, d_free, csc, csr_flags, row_ptr, col, w, cap, nnz_off, st));
  if (nbr) {
    k_rec_to_nbr<<<(int)((edges + 255) / 256), 256, 0, st>>>(d_rec, ns, k, edges, nbr);
    m->launches++;
  }
  CK(cudaGetLastError());
  return CTP_OK;
}

extern "C" int ctp_build_prm_dev(ctp_map *m, const double *nodes, int64_t q, int64_t n, int k, double clr,
                                 double cdc, int32_t *nbr, uint8_t *directed_free, int32_t *row_ptr, int32_t *col,
                                 double *w, int64_t cap, int64_t *nnz_off, void *stream) {
  return ctp_build_prm_ex_dev(m, nodes, q, n, k, clr, cdc, 0, nbr, directed_free, row_ptr, col, w, cap, nnz_off, stream);
}

static int build_prm_host_tail(ctp_map *m, int64_t q, int64_t n, int k, int32_t *nbr, uint8_t *directed_free,
                               int32_t *row_ptr, int32_t *col, double *w, int64_t cap, int64_t *nnz_off,
                               const int32_t *d_nbr, const uint8_t *d_free, const int32_t *d_rp, const int32_t *d_col,
                               const double *d_w, const int64_t *d_off) {
  CK(cudaMemcpyAsync(nnz_off, d_off, (q + 1) * 8, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(row_ptr, d_rp, q * (n + 1) * 4, cudaMemcpyDeviceToHost, 0));
  if (nbr) CK(cudaMemcpyAsync(nbr, d_nbr, q * n * k * 4, cudaMemcpyDeviceToHost, 0));
  if (directed_free) CK(cudaMemcpyAsync(directed_free, d_free, q * n * k, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  const int64_t nnz = nnz_off[q];
  if (nnz > cap) return fail(CTP_ERR_CAPACITY, "CSR needs %lld entries, capacity is %lld", (long long)nnz, (long long)cap);
  if (nnz) {
    CK(cudaMemcpyAsync(col, d_col, nnz * 4, cudaMemcpyDeviceToHost, 0));
    CK(cudaMemcpyAsync(w, d_w, nnz * 8, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
  }
  return CTP_OK;
}

extern "C" int ctp_build_prm(ctp_map *m, const double *nodes, int64_t q, int64_t n, int k, double clr, double cdc,
                             int32_t *nbr, uint8_t *directed_free, int32_t *row_ptr, int32_t *col, double *w,
                             int64_t cap, int64_t *nnz_off) {
  REQUIRE(m && q >= 1 && n >= 2 && k >= 1 && k <= CTP_KNN_MAXK, "bad argument (k must be in [1,16])");
  REQUIRE(nodes && row_ptr && col && w && nnz_off, "null buffer");
  CK(cudaSetDevice(m->device));
  const int64_t edges = q * n * k, dcap = 2 * edges;
  CKR(arena_reserve(m->io, align_up(q * n * 24) + align_up(edges * 4) + align_up(edges) + align_up(q * (n + 1) * 4) +
                               align_up(dcap * 4) + align_up(dcap * 8) + align_up((q + 1) * 8)));
  IoScope sc(m);
  double *d_n = arena_take<double>(m->io, q * n * 3);
  int32_t *d_nbr = arena_take<int32_t>(m->io, edges);
  uint8_t *d_free = arena_take<uint8_t>(m->io, edges);
  int32_t *d_rp = arena_take<int32_t>(m->io, q * (n + 1));
  int32_t *d_col = arena_take<int32_t>(m->io, dcap);
  double *d_w = arena_take<double>(m->io, dcap);
  int64_t *d_off = arena_take<int64_t>(m->io, q + 1);
  CK(cudaMemcpyAsync(d_n, nodes, q * n * 24, cudaMemcpyHostToDevice, 0));
  m->ws.used = 0;
  CKR(ctp_build_prm_dev(m, d_n, q, n, k, clr, cdc, d_nbr, d_free, d_rp, d_col, d_w, dcap, d_off, 0));
  return build_prm_host_tail(m, q, n, k, nbr, directed_free, row_ptr, col, w, cap, nnz_off, d_nbr, d_free, d_rp,
                             d_col, d_w, d_off);
}

// CSR from a given k-NN table and its directed verdicts (the merge step when the n*k checks were
// sharded over several GPUs): same symmetric-union rule as ctp_build_prm
// the middle stage of the PRM build on its own: the q*n*k directed checks of a k-NN table (stride k, -1 = no
// neighbour) against THIS map.  The neighbour and the adjacency stages do not look at the map, so a sweep over many
// maps can run them once for all queries and only this stage (and the rejection) map by map.
extern "C" int ctp_edge_check_knn_dev(ctp_map *m, const double *nodes, int64_t q, int64_t n, int k, const int32_t *nbr,
                                      double clr, double cdc, uint8_t *directed_free, void *stream) {
  REQUIRE(m && q >= 1 && n >= 2 && k >= 1 && k <= CTP_KNN_MAXK, "bad argument (k must be in [1,16])");
  REQUIRE(nodes && nbr && directed_free, "null buffer");
  REQUIRE(cdc > 0, "collision_distance_check must be > 0");
  REQUIRE(q * n * k < ((int64_t)1 << 31), "q*n*k must fit in int32");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t edges = q * n * k;
  EdgeSrc S{nodes, nullptr, nullptr, nbr, n, k, 1, k, nullptr};
  S.small = 1;
  S.dk = ctp_div_make((uint32_t)k);
  S.dn = ctp_div_make((uint32_t)n);
  EdgeOut O{directed_free, nullptr, nullptr, nullptr, m->d_lookups};
  const size_t keep = m->ws.used;
  CKR(arena_reserve(m->ws, keep + ws_bytes_edge_filter(q * n, edges) + 1024));
  const EdgeScratch esc = edge_scratch_take(m, q * n, edges);
  m->ws.used = keep;  // dead once the kernels below have run (stream order)
  StageSpan span(m, CTP_STAGE_EDGES, st);
  return launch_edges_prm(m, m->prm_group, S, q * n, edges, clr, cdc, O, esc, st);
}

// One GPU's share of a single large query split over several: the neighbour search and the
// directed checks for the nodes at positions [first, last) of the k-NN grid's cell order (moved to cell boundaries, so
// that the GPUs agree on who owns which node; a compact part of space, so different GPUs march through different parts
// of the map).  Rows of other nodes come back as nbr = 0x80808080 (below every valid entry, -1 included) and
// directed_free = 0: an element-wise maximum over the GPUs (one all-reduce each) assembles the full tables.
extern "C" int ctp_prm_rows_dev(ctp_map *m, const double *nodes, int64_t n, int k, double clr, double cdc, int64_t first,
                                int64_t last, int32_t *nbr, uint8_t *directed_free, int64_t *rows_done, void *stream) {
  REQUIRE(m && n >= 2 && k >= 1 && k <= CTP_KNN_MAXK, "bad argument (k must be in [1,16])");
  REQUIRE(nodes && nbr && directed_free, "null buffer");
  REQUIRE(first >= 0 && first <= last && last <= n, "slice must satisfy 0 <= first <= last <= n");
  REQUIRE(cdc > 0, "collision_distance_check must be > 0");
  REQUIRE(n * k < ((int64_t)1 << 31), "n*k must fit in int32");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t keep = m->ws.used;
  CKR(arena_reserve(m->ws, keep + ws_bytes_prm(1, n, k) + 8192));
  const int64_t edges = n * k;
  const EdgeScratch esc = edge_scratch_take(m, n, edges);
  CK(cudaMemsetAsync(nbr, 0x80, edges * 4, st));
  CK(cudaMemsetAsync(directed_free, 0, edges, st));
  const float4 *d_sorted = nullptr;
  int64_t slice[2] = {first, last};
  int rc;
  {
    StageSpan span(m, CTP_STAGE_KNN, st);
    rc = knn_launch(m, nodes, 1, n, k, nbr, k, nullptr, st, &d_sorted, slice);
  }
  m->ws.used = keep;
  CKR(rc);
  if (rows_done) *rows_done = slice[1] - slice[0];
  if (slice[1] <= slice[0]) return CTP_OK;
  // work item r*k + slot = row order[r].w of the slice (the flat kernel walks the slice in cell order)
  EdgeSrc S{nodes, nullptr, nullptr, nbr, n, k, 1, k, d_sorted + slice[0]};
  S.small = 1;
  S.dk = ctp_div_make((uint32_t)k);
  S.dn = ctp_div_make((uint32_t)n);
  EdgeOut O{directed_free, nullptr, nullptr, nullptr, m->d_lookups};
  StageSpan span(m, CTP_STAGE_EDGES, st);
  return launch_edges_prm(m, 1, S, n, (slice[1] - slice[0]) * k, clr, cdc, O, esc, st);
}

extern "C" int ctp_csr_from_directed_dev(ctp_map *m, const double *nodes, int64_t q, int64_t n, int k, const int32_t *nbr,
                                         const uint8_t *directed_free, int32_t *row_ptr, int32_t *col, double *w,
                                         int64_t cap, int64_t *nnz_off, void *stream) {
  REQUIRE(m && q >= 1 && n >= 2 && k >= 1 && k <= CTP_KNN_MAXK, "bad argument (k must be in [1,16])");
  REQUIRE(nodes && nbr && directed_free && row_ptr && col && w && nnz_off && cap >= 0, "null buffer");
  REQUIRE(q * n * k < ((int64_t)1 << 31), "q*n*k must fit in int32");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t keep = m->ws.used;
  CKR(arena_reserve(m->ws, keep + ws_bytes_prm(q, n, k) + 8192));
  const int64_t rows = q * n, edges = rows * k;
  const int ns = csr_ns(k);
  int32_t *d_rec = arena_take<int32_t>(m->ws, rows * ns);
  const CsrScratch csc = csr_scratch_take(m, q, n, k);
  m->ws.used = keep;
  k_nbr_to_rec<<<(int)((edges + 255) / 256), 256, 0, st>>>(nbr, ns, k, edges, d_rec);
  m->launches++;
  return csr_emit(m, nodes, q, n, k, d_rec, directed_free, csc, 0, row_ptr, col, w, cap, nnz_off, st);
}

extern "C" int ctp_csr_from_directed(ctp_map *m, const double *nodes, int64_t q, int64_t n, int k, const int32_t *nbr,
                                     const uint8_t *directed_free, int32_t *row_ptr, int32_t *col, double *w, int64_t cap,
                                     int64_t *nnz_off) {
  REQUIRE(m && q >= 1 && n >= 2 && k >= 1 && k <= CTP_KNN_MAXK, "bad argument (k must be in [1,16])");
  REQUIRE(nodes && nbr && directed_free && row_ptr && col && w && nnz_off, "null buffer");
  CK(cudaSetDevice(m->device));
  const int64_t edges = q * n * k, dcap = 2 * edges;
  for (int64_t i = 0; i < edges; i++) REQUIRE(nbr[i] >= -1 && nbr[i] < n, "neighbour index out of range (-1 = no neighbour)");
  CKR(arena_reserve(m->io, align_up(q * n * 24) + align_up(edges * 4) + align_up(edges) + align_up(q * (n + 1) * 4) +
                               align_up(dcap * 4) + align_up(dcap * 8) + align_up((q + 1) * 8)));
  IoScope sc(m);
  double *d_n = arena_take<double>(m->io, q * n * 3);
  int32_t *d_nbr = arena_take<int32_t>(m->io, edges);
  uint8_t *d_free = arena_take<uint8_t>(m->io, edges);
  int32_t *d_rp = arena_take<int32_t>(m->io, q * (n + 1));
  int32_t *d_col = arena_take<int32_t>(m->io, dcap);
  double *d_w = arena_take<double>(m->io, dcap);
  int64_t *d_off = arena_take<int64_t>(m->io, q + 1);
  CK(cudaMemcpyAsync(d_n, nodes, q * n * 24, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_nbr, nbr, edges * 4, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_free, directed_free, edges, cudaMemcpyHostToDevice, 0));
  m->ws.used = 0;
  CKR(ctp_csr_from_directed_dev(m, d_n, q, n, k, d_nbr, d_free, d_rp, d_col, d_w, dcap, d_off, 0));
  return build_prm_host_tail(m, q, n, k, nullptr, nullptr, row_ptr, col, w, cap, nnz_off, nullptr, nullptr, d_rp, d_col,
                             d_w, d_off);
}

// ----------------------------- shortest paths over the CSR roadmaps ---------------------------
static int sssp_launch(ctp_map *m, int64_t q, int64_t n, const int32_t *row_ptr, const int32_t *col, const double *w,
                       const int64_t *nnz_off, const int32_t *src, int ns, double *dist, int32_t *pred, int32_t *label,
                       const double *dist_init, void *stream);
extern "C" int ctp_sssp_dev(ctp_map *m, int64_t q, int64_t n, const int32_t *row_ptr, const int32_t *col, const double *w,
                            const int64_t *nnz_off, const int32_t *src, int ns, double *dist, int32_t *pred,
                            int32_t *label, void *stream) {
  return sssp_launch(m, q, n, row_ptr, col, w, nnz_off, src, ns, dist, pred, label, nullptr, stream);
}
// dist_init (optional, q x n): upper bounds the search starts from instead of +inf
static int sssp_launch(ctp_map *m, int64_t q, int64_t n, const int32_t *row_ptr, const int32_t *col, const double *w,
                       const int64_t *nnz_off, const int32_t *src, int ns, double *dist, int32_t *pred, int32_t *label,
                       const double *dist_init, void *stream) {
  REQUIRE(m && q >= 1 && n >= 1 && ns >= 1 && ns <= CTP_SSSP_THREADS, "bad argument (1 <= sources per query <= 1024)");
  REQUIRE(row_ptr && col && w && nnz_off && src && dist && pred, "null buffer");
  REQUIRE(n < ((int64_t)1 << 31), "n must fit in int32");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)n * 9 + 16;
  // frontier variant: distances + bit mask + u16 queue (node ids must fit 16 bits)
  const size_t smem_f = (size_t)n * 8 + (((size_t)n + 31) / 32) * 4 + (size_t)n * 2 + 16;
  // per handle, not per process: function attributes belong to the device the handle lives on
  int &smem_optin = m->sssp_smem_optin;
  if (smem_optin < 0) {
    int v = 0;
    CK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device));
    smem_optin = v;
    CK(cudaFuncSetAttribute(k_sssp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
    CK(cudaFuncSetAttribute(k_sssp_frontier, cudaFuncAttributeMaxDynamicSharedMemorySize, v - 8192));
    CK(cudaFuncSetAttribute(k_sssp_cluster<CTP_SSSP_CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, v - 8192));
    CK(cudaFuncSetAttribute(k_sssp_cluster<CTP_SSSP_CL_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, v - 8192));
    if (cudaFuncSetAttribute(k_sssp_cluster<CTP_SSSP_CL_BIG>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      m->sssp_cluster_big = false;
    }
  }
  // one query per cluster of CTAs (distances sliced over distributed shared memory): when few queries are in
  // flight, or when a query does not fit one CTA's shared memory.  16 CTAs (opt-in cluster size) when the GPU would
  // otherwise idle: twice the shared memory, so the rows of a 10^4-node roadmap are staged on chip as well.
  const bool fits_one = n <= 65535 && smem_f <= (size_t)smem_optin - 8192;
  auto launch_cluster = [&](int CL) -> bool {
    const int slice = (int)(((n + CL - 1) / CL + 31) / 32 * 32);
    const size_t base = sssp_cluster_smem(slice, 0), limit = (size_t)smem_optin - 8192;
    if (base > limit) return false;
    int cap_e = (int)std::min<size_t>((limit - base) / 12, (size_t)1 << 24);
    if (cap_e < 4 * slice) cap_e = 0;  // not worth staging
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(q * CL));
    cfg.blockDim = dim3(CTP_SSSP_CL_THREADS);
    cfg.dynamicSmemBytes = sssp_cluster_smem(slice, cap_e);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = CL == CTP_SSSP_CL_BIG
                        ? cudaLaunchKernelEx(&cfg, k_sssp_cluster<CTP_SSSP_CL_BIG>, n, slice, cap_e, row_ptr, col, w, nnz_off, src,
                                             ns, dist, pred, dist_init)
                        : cudaLaunchKernelEx(&cfg, k_sssp_cluster<CTP_SSSP_CL>, n, slice, cap_e, row_ptr, col, w, nnz_off, src, ns,
                                             dist, pred, dist_init);
    if (e != cudaSuccess) cudaGetLastError();  // forget it: the caller falls back
    return e == cudaSuccess;
  };
  const bool cluster_ok = m->sssp_frontier && !label && n >= 64;
  const bool want_cluster = m->sssp_cluster < 0 ? (q * CTP_SSSP_CL <= 2 * (int64_t)m->sm_count || !fits_one) : m->sssp_cluster != 0;
  if (cluster_ok && want_cluster) {
    if (m->sssp_cluster_big && q * CTP_SSSP_CL_BIG <= 2 * (int64_t)m->sm_count && n >= 1024) {
      if (launch_cluster(CTP_SSSP_CL_BIG)) {
        m->launches++;
        return CTP_OK;
      }
      m->sssp_cluster_big = false;  // this device or partition cannot co-schedule 16 CTAs
    }
    if (launch_cluster(CTP_SSSP_CL)) {
      m->launches++;
      return CTP_OK;
    }
    if (sssp_cluster_smem((int)(((n + CTP_SSSP_CL - 1) / CTP_SSSP_CL + 31) / 32 * 32), 0) <= (size_t)smem_optin - 8192)
      m->sssp_cluster = 0;  // the launch itself failed: use the one-CTA kernels from now on
  }
  if (m->sssp_frontier && fits_one) {
    k_sssp_frontier<<<(unsigned)q, CTP_SSSP_THREADS, smem_f, st>>>(n, row_ptr, col, w, nnz_off, src, ns, dist, pred, label, dist_init);
  } else if (smem <= (size_t)smem_optin) {
    k_sssp<true><<<(unsigned)q, CTP_SSSP_THREADS, smem, st>>>(n, row_ptr, col, w, nnz_off, src, ns, dist, pred, label,
                                                              nullptr, nullptr, dist_init);
  } else {  // distances do not fit in shared memory: same algorithm on a global (L2-resident) array
    const size_t keep = m->ws.used;
    CKR(arena_reserve(m->ws, keep + align_up((size_t)q * n * 8) + align_up((size_t)q * n) + 1024));
    unsigned long long *g_dist = arena_take<unsigned long long>(m->ws, (size_t)q * n);
    uint8_t *g_chg = arena_take<uint8_t>(m->ws, (size_t)q * n);
    m->ws.used = keep;
    k_sssp<false><<<(unsigned)q, CTP_SSSP_THREADS, 0, st>>>(n, row_ptr, col, w, nnz_off, src, ns, dist, pred, label, g_dist,
                                                            g_chg, dist_init);
  }
  m->launches++;
  CK(cudaGetLastError());
  return CTP_OK;
}

extern "C" int ctp_sssp(ctp_map *m, int64_t q, int64_t n, const int32_t *row_ptr, const int32_t *col, const double *w,
                        const int64_t *nnz_off, const int32_t *src, int ns, double *dist, int32_t *pred, int32_t *label) {
  REQUIRE(m && q >= 1 && n >= 1 && ns >= 1, "bad argument");
  REQUIRE(row_ptr && col && w && nnz_off && src && dist && pred, "null buffer");
  CK(cudaSetDevice(m->device));
  const int64_t nnz = nnz_off[q];
  REQUIRE(nnz >= 0 && nnz_off[0] == 0, "nnz_off must start at 0");
  for (int64_t i = 0; i < nnz; i++) REQUIRE(col[i] >= 0 && col[i] < n, "column index out of range");
  CKR(arena_reserve(m->io, align_up(q * (n + 1) * 4) + align_up(nnz * 4 + 4) + align_up(nnz * 8 + 8) + align_up((q + 1) * 8) +
                               align_up(q * ns * 4) + align_up(q * n * 8) + 2 * align_up(q * n * 4)));
  IoScope sc(m);
  int32_t *d_rp = arena_take<int32_t>(m->io, q * (n + 1));
  int32_t *d_col = arena_take<int32_t>(m->io, nnz + 1);
  double *d_w = arena_take<double>(m->io, nnz + 1);
  int64_t *d_off = arena_take<int64_t>(m->io, q + 1);
  int32_t *d_src = arena_take<int32_t>(m->io, q * ns);
  double *d_dist = arena_take<double>(m->io, q * n);
  int32_t *d_pred = arena_take<int32_t>(m->io, q * n), *d_label = arena_take<int32_t>(m->io, q * n);
  CK(cudaMemcpyAsync(d_rp, row_ptr, q * (n + 1) * 4, cudaMemcpyHostToDevice, 0));
  if (nnz) {
    CK(cudaMemcpyAsync(d_col, col, nnz * 4, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(d_w, w, nnz * 8, cudaMemcpyHostToDevice, 0));
  }
  CK(cudaMemcpyAsync(d_off, nnz_off, (q + 1) * 8, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_src, src, q * ns * 4, cudaMemcpyHostToDevice, 0));
  m->ws.used = 0;
  CKR(ctp_sssp_dev(m, q, n, d_rp, d_col, d_w, d_off, d_src, ns, d_dist, d_pred, label ? d_label : nullptr, 0));
  CK(cudaMemcpyAsync(dist, d_dist, q * n * 8, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(pred, d_pred, q * n * 4, cudaMemcpyDeviceToHost, 0));
  if (label) CK(cudaMemcpyAsync(label, d_label, q * n * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return CTP_OK;
}

// ----------------------------- the floods of the cluster stage --------------------------------
// One flood for q roadmaps: shortest-path launch (all seeds, or the newest seed on top of the distances of the
// previous floods), then predecessors / labels / update counts and the inter-cluster connection records that the
// sequential loop of the reference would have produced (ctp_cluster.cuh).
extern "C" int ctp_cluster_flood_dev(ctp_map *m, int64_t q, int64_t n, const int32_t *row_ptr, const int32_t *col,
                                     const double *w, const int64_t *nnz_off, const int32_t *seeds, int max_seeds,
                                     const int32_t *n_seeds, int mode, double *dist, int32_t *prev, int32_t *label,
                                     int64_t rec_cap, int32_t *rec_nodes, int32_t *rec_label, double *rec_dist,
                                     int32_t *rec_count, uint64_t *n_updates, void *stream) {
  REQUIRE(m && q >= 1 && n >= 1 && max_seeds >= 1 && max_seeds <= CTP_SSSP_THREADS, "bad argument (1 <= max_seeds <= 1024)");
  REQUIRE(mode == 0 || mode == 1, "mode must be 0 (flood from all seeds) or 1 (re-flood from the newest seed)");
  REQUIRE(row_ptr && col && w && nnz_off && seeds && n_seeds && dist && prev && label && rec_count && n_updates, "null buffer");
  REQUIRE(rec_cap >= 0 && rec_cap < ((int64_t)1 << 31) && (rec_cap == 0 || (rec_nodes && rec_label && rec_dist)), "bad record capacity");
  REQUIRE(n < ((int64_t)1 << 31), "n must fit in int32");
  CK(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = q * n;
  const size_t keep = m->ws.used;
  // (the shortest-path launch may take total * 9 bytes more for its global-array variant: reserved here, so that the
  // arena cannot move under the buffers taken below)
  CKR(arena_reserve(m->ws, keep + 2 * align_up(total * 8) + 3 * align_up(total * 4) + align_up(q * max_seeds * 4) + 8192));
  double *d_new = arena_take<double>(m->ws, total);
  int32_t *d_pred = arena_take<int32_t>(m->ws, total);
  int32_t *d_lnew = arena_take<int32_t>(m->ws, total);
  int32_t *d_src = arena_take<int32_t>(m->ws, q * max_seeds);
  const int tb = 256;
  const unsigned gb = (unsigned)((total + tb - 1) / tb);
  CK(cudaMemsetAsync(rec_count, 0, q * sizeof(int32_t), st));
  CK(cudaMemsetAsync(n_updates, 0, q * sizeof(uint64_t), st));
  k_cluster_sources<<<(unsigned)((q * max_seeds + tb - 1) / tb), tb, 0, st>>>(q, seeds, max_seeds, n_seeds, mode, d_src);
  m->launches++;
  int rc = sssp_launch(m, q, n, row_ptr, col, w, nnz_off, d_src, max_seeds, d_new, d_pred, nullptr, mode == 1 ? dist : nullptr, st);
  if (rc != CTP_OK) {
    m->ws.used = keep;
    return rc;
  }
  if (mode == 1) CK(cudaMemcpyAsync(d_lnew, label, total * 4, cudaMemcpyDeviceToDevice, st));
  int32_t *lab_w = mode == 1 ? d_lnew : label;
  // a warp per node while that still is a modest launch (one query, a few queries), a thread per node for big batches
  const bool wide = total <= ((int64_t)1 << 20);
  const unsigned gw = (unsigned)((total * 32 + tb - 1) / tb);
  if (wide)
    k_cluster_prev<32><<<gw, tb, 0, st>>>(q, n, row_ptr, col, w, nnz_off, seeds, max_seeds, n_seeds, mode, dist, d_new, prev, lab_w,
                                          (unsigned long long *)n_updates);
  else
    k_cluster_prev<1><<<gb, tb, 0, st>>>(q, n, row_ptr, col, w, nnz_off, seeds, max_seeds, n_seeds, mode, dist, d_new, prev, lab_w,
                                         (unsigned long long *)n_updates);
  if (mode == 0) k_cluster_label<<<gb, tb, 0, st>>>(q, n, n_seeds, prev, d_new, label);
  if (wide)
    k_cluster_records<32><<<gw, tb, 0, st>>>(q, n, row_ptr, col, w, nnz_off, seeds, max_seeds, n_seeds, mode, dist, label, d_new, lab_w,
                                             prev, (int)rec_cap, rec_nodes, rec_label, rec_dist, rec_count);
  else
    k_cluster_records<1><<<gb, tb, 0, st>>>(q, n, row_ptr, col, w, nnz_off, seeds, max_seeds, n_seeds, mode, dist, label, d_new, lab_w,
                                            prev, (int)rec_cap, rec_nodes, rec_label, rec_dist, rec_count);
  k_cluster_commit<<<gb, tb, 0, st>>>(total, n, n_seeds, d_new, mode == 1 ? d_lnew : nullptr, dist, label);
  m->launches += mode == 0 ? 4 : 3;
  m->ws.used = keep;
  CK(cudaGetLastError());
  return CTP_OK;
}

// q roadmaps resident on the device together with the state of their floods (distance to the nearest seed, predecessor,
// cluster label): what clusterGraph keeps in its HeapNodes between wavefrontFill and the addCentroid calls
struct ctp_roadmap {
  ctp_map *m = nullptr;
  int64_t q = 0, n = 0, nnz = 0, rec_cap = 0;
  int max_seeds = 0;
  int32_t *d_rp = nullptr, *d_col = nullptr, *d_prev = nullptr, *d_label = nullptr, *d_seeds = nullptr, *d_ns = nullptr;
  int32_t *d_rec_nodes = nullptr, *d_rec_label = nullptr, *d_rec_count = nullptr;
  double *d_w = nullptr, *d_dist = nullptr, *d_rec_dist = nullptr;
  int64_t *d_off = nullptr;
  uint64_t *d_upd = nullptr;
  void *block = nullptr;  // all of the above live in this one allocation
  size_t block_cap = 0;
  std::vector<int64_t> h_off;
};

extern "C" int ctp_roadmap_destroy(ctp_roadmap *r) {
  if (!r) return CTP_OK;
  ctp_map *m = r->m;
  if (m) cudaSetDevice(m->device);
  if (r->block) {
    cudaDeviceSynchronize();  // nothing may still be reading the block when its next owner starts writing
    if (m && r->block_cap > m->rm_block_cap) {  // keep the larger of the two blocks for the next roadmap of this map
      if (m->rm_block) cudaFree(m->rm_block);
      m->rm_block = r->block;
      m->rm_block_cap = r->block_cap;
    } else {
      cudaFree(r->block);
    }
  }
  delete r;
  return CTP_OK;
}

extern "C" int ctp_roadmap_create(ctp_map *m, int64_t q, int64_t n, const int32_t *row_ptr, const int32_t *col,
                                  const double *w, const int64_t *nnz_off, int max_seeds, ctp_roadmap **out) {
  REQUIRE(m && out && q >= 1 && n >= 1 && max_seeds >= 1 && max_seeds <= CTP_SSSP_THREADS, "bad argument (1 <= max_seeds <= 1024)");
  REQUIRE(row_ptr && col && w && nnz_off, "null buffer");
  REQUIRE(n < ((int64_t)1 << 31), "n must fit in int32");
  const int64_t nnz = nnz_off[q];
  REQUIRE(nnz >= 0 && nnz_off[0] == 0, "nnz_off must start at 0");
  int64_t widest = 0;
  for (int64_t i = 0; i < q; i++) {
    REQUIRE(nnz_off[i + 1] >= nnz_off[i], "nnz_off must not decrease");
    REQUIRE(row_ptr[i * (n + 1)] == 0 && row_ptr[i * (n + 1) + n] == nnz_off[i + 1] - nnz_off[i], "row_ptr does not match nnz_off");
    widest = std::max(widest, nnz_off[i + 1] - nnz_off[i]);
  }
  for (int64_t i = 0; i < nnz; i++) REQUIRE(col[i] >= 0 && col[i] < n, "column index out of range");
  CK(cudaSetDevice(m->device));
  ctp_roadmap *r = new (std::nothrow) ctp_roadmap();
  if (!r) return fail(CTP_ERR_NOMEM, "out of host memory");
  r->m = m;
  r->q = q;
  r->n = n;
  r->nnz = nnz;
  r->max_seeds = max_seeds;
  r->rec_cap = std::max<int64_t>(widest, 1);  // a record per directed edge at most
  r->h_off.assign(nnz_off, nnz_off + q + 1);
  // one block for everything (taken over from the last destroyed roadmap of this map when it is large enough)
  size_t need = 0;
  auto sz = [&](size_t bytes) {
    const size_t at = need;
    need += align_up(bytes ? bytes : 1);
    return at;
  };
  const size_t o_rp = sz(q * (n + 1) * 4), o_col = sz(nnz * 4 + 4), o_w = sz(nnz * 8 + 8), o_off = sz((q + 1) * 8);
  const size_t o_dist = sz(q * n * 8), o_prev = sz(q * n * 4), o_label = sz(q * n * 4), o_seeds = sz(q * max_seeds * 4);
  const size_t o_ns = sz(q * 4), o_rn = sz(q * r->rec_cap * 8), o_rl = sz(q * r->rec_cap * 4), o_rd = sz(q * r->rec_cap * 8);
  const size_t o_rc = sz(q * 4), o_upd = sz(q * 8);
  if (m->rm_block && m->rm_block_cap >= need) {
    r->block = m->rm_block;
    r->block_cap = m->rm_block_cap;
    m->rm_block = nullptr;
    m->rm_block_cap = 0;
  } else {
    const cudaError_t e = cudaMalloc(&r->block, need);
    if (e != cudaSuccess) {
      r->block = nullptr;
      ctp_roadmap_destroy(r);
      return fail(e == cudaErrorMemoryAllocation ? CTP_ERR_NOMEM : CTP_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    r->block_cap = need;
  }
  char *bp = (char *)r->block;
  r->d_rp = (int32_t *)(bp + o_rp);
  r->d_col = (int32_t *)(bp + o_col);
  r->d_w = (double *)(bp + o_w);
  r->d_off = (int64_t *)(bp + o_off);
  r->d_dist = (double *)(bp + o_dist);
  r->d_prev = (int32_t *)(bp + o_prev);
  r->d_label = (int32_t *)(bp + o_label);
  r->d_seeds = (int32_t *)(bp + o_seeds);
  r->d_ns = (int32_t *)(bp + o_ns);
  r->d_rec_nodes = (int32_t *)(bp + o_rn);
  r->d_rec_label = (int32_t *)(bp + o_rl);
  r->d_rec_dist = (double *)(bp + o_rd);
  r->d_rec_count = (int32_t *)(bp + o_rc);
  r->d_upd = (uint64_t *)(bp + o_upd);
  int rc = CTP_OK;
  auto up = [&](void *d, const void *h, size_t bytes) {
    if (rc == CTP_OK && bytes && cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, 0) != cudaSuccess)
      rc = fail(CTP_ERR_CUDA, "host -> device copy of the roadmap failed");
  };
  up(r->d_rp, row_ptr, q * (n + 1) * 4);
  up(r->d_col, col, nnz * 4);
  up(r->d_w, w, nnz * 8);
  up(r->d_off, nnz_off, (q + 1) * 8);
  if (rc == CTP_OK && cudaStreamSynchronize(0) != cudaSuccess) rc = fail(CTP_ERR_CUDA, "roadmap upload failed");
  if (rc != CTP_OK) {
    ctp_roadmap_destroy(r);
    return rc;
  }
  *out = r;
  return CTP_OK;
}

// Host form: seeds in, the whole state and the records out.  The records of each query come back in the order the
// sequential loop appends them: ascending (distance of the expanding node, its id), neighbours ascending.
extern "C" int ctp_cluster_flood(ctp_roadmap *r, const int32_t *seeds, const int32_t *n_seeds, int mode, double *dist,
                                 int32_t *prev, int32_t *label, int32_t *rec_nodes, int32_t *rec_label, double *rec_dist,
                                 int32_t *rec_count, uint64_t *n_updates) {
  REQUIRE(r && seeds && n_seeds && dist && prev && label && rec_nodes && rec_label && rec_dist && rec_count && n_updates,
          "null argument");
  ctp_map *m = r->m;
  const int64_t q = r->q, n = r->n;
  for (int64_t i = 0; i < q; i++) {
    REQUIRE(n_seeds[i] >= 0 && n_seeds[i] <= r->max_seeds, "n_seeds out of range");
    for (int s = 0; s < n_seeds[i]; s++)
      REQUIRE(seeds[i * r->max_seeds + s] >= 0 && seeds[i * r->max_seeds + s] < n, "seed id out of range");
  }
  CK(cudaSetDevice(m->device));
  CK(cudaMemcpyAsync(r->d_seeds, seeds, q * r->max_seeds * 4, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(r->d_ns, n_seeds, q * 4, cudaMemcpyHostToDevice, 0));
  m->ws.used = 0;
  CKR(ctp_cluster_flood_dev(m, q, n, r->d_rp, r->d_col, r->d_w, r->d_off, r->d_seeds, r->max_seeds, r->d_ns, mode, r->d_dist,
                            r->d_prev, r->d_label, r->rec_cap, r->d_rec_nodes, r->d_rec_label, r->d_rec_dist, r->d_rec_count, r->d_upd, 0));
  // results through the page-locked staging block: asynchronous copies at link speed, one synchronisation
  const size_t o_dist = 0, o_prev = o_dist + align_up(q * n * 8), o_label = o_prev + align_up(q * n * 4),
               o_cnt = o_label + align_up(q * n * 4), o_upd = o_cnt + align_up(q * 4), o_end = o_upd + align_up(q * 8);
  CKR(stage_reserve(m, o_end + (size_t)r->rec_cap * 20 + 1024));
  char *hs = m->h_stage;
  CK(cudaMemcpyAsync(hs + o_dist, r->d_dist, q * n * 8, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(hs + o_prev, r->d_prev, q * n * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(hs + o_label, r->d_label, q * n * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(hs + o_cnt, r->d_rec_count, q * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(hs + o_upd, r->d_upd, q * 8, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  memcpy(dist, hs + o_dist, q * n * 8);
  memcpy(prev, hs + o_prev, q * n * 4);
  memcpy(label, hs + o_label, q * n * 4);
  memcpy(rec_count, hs + o_cnt, q * 4);
  memcpy(n_updates, hs + o_upd, q * 8);
  std::vector<int32_t> order;
  for (int64_t i = 0; i < q; i++) {
    const int64_t c = rec_count[i];
    if (c > r->rec_cap) return fail(CTP_ERR_CAPACITY, "more connection records than directed edges (internal error)");
    if (c == 0) continue;
    int32_t *tmp_label = (int32_t *)(hs + o_end);
    int32_t *tmp_nodes = (int32_t *)(hs + o_end + align_up(c * 4));
    double *tmp_dist = (double *)(hs + o_end + align_up(c * 4) + align_up(c * 8));
    CK(cudaMemcpyAsync(tmp_label, r->d_rec_label + i * r->rec_cap, c * 4, cudaMemcpyDeviceToHost, 0));
    CK(cudaMemcpyAsync(tmp_nodes, r->d_rec_nodes + 2 * i * r->rec_cap, c * 8, cudaMemcpyDeviceToHost, 0));
    CK(cudaMemcpyAsync(tmp_dist, r->d_rec_dist + i * r->rec_cap, c * 8, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    order.resize(c);
    for (int64_t t = 0; t < c; t++) order[t] = (int32_t)t;
    const double *dq = dist + i * n;
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
      const int ea = tmp_nodes[2 * a], eb = tmp_nodes[2 * b];
      if (ea != eb) return dq[ea] < dq[eb] || (dq[ea] == dq[eb] && ea < eb);
      return tmp_nodes[2 * a + 1] < tmp_nodes[2 * b + 1];
    });
    int32_t *on = rec_nodes + 2 * i * r->rec_cap;
    double *od = rec_dist + i * r->rec_cap;
    int32_t *ol = rec_label + i * r->rec_cap;
    for (int64_t t = 0; t < c; t++) {
      ol[t] = tmp_label[order[t]];
      on[2 * t] = tmp_nodes[2 * order[t]];
      on[2 * t + 1] = tmp_nodes[2 * order[t] + 1];
      od[t] = tmp_dist[order[t]];
    }
  }
  return CTP_OK;
}
extern "C" int ctp_roadmap_info(const ctp_roadmap *r, int64_t info[4]) {
  REQUIRE(r && info, "null argument");
  info[0] = r->q;
  info[1] = r->n;
  info[2] = r->nnz;
  info[3] = r->rec_cap;
  return CTP_OK;
}

// Host candidate streams -> host CSRs.  The q queries are cut into chunks that flow through
// streams for the host->device copy, the kernels and the device->host copies with double-buffered device staging, so
// the PCIe transfers in both directions overlap the kernels; with page-locked host buffers
// (ctp_host_alloc) the call runs at the speed of the slowest of the three.
struct PipeBuf {
  double *d_s, *d_g, *d_c, *d_n;
  int32_t *d_nf, *d_nu, *d_rp, *d_col;
  double *d_w;
  int64_t *d_off;
};

static int pipe_init(ctp_map *m) {
  if (m->pipe_ready) return CTP_OK;
  for (int i = 0; i < 4; i++) CK(cudaStreamCreateWithFlags(&m->pipe_st[i], cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    CK(cudaEventCreateWithFlags(&m->ev_in[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_comp[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_small[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_out[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_rej[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&m->ev_rejd[i], cudaEventDisableTiming));
  }
  m->pipe_ready = true;
  return CTP_OK;
}

// pinned host staging of the per-chunk offsets and accept counts (two buffers of qc entries each)
// small results the host waits for (accept counts, CSR offsets) are WRITTEN into mapped page-locked memory by a kernel on
// the compute stream: a cudaMemcpyAsync of a few hundred bytes would queue behind the tens of megabytes of the previous
// chunk's outputs on the device->host copy engine (measured: 0.6 ms per chunk with the compute stream drained)
// int32 column ids -> uint16 (CTP_CSR_COL16), four per thread
__global__ void __launch_bounds__(256) k_narrow_u16(const int32_t *__restrict__ in, int64_t cnt, uint16_t *__restrict__ out) {
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4; i < cnt; i += (int64_t)gridDim.x * blockDim.x * 4) {
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i + u < cnt) out[i + u] = (uint16_t)in[i + u];
  }
}

template <typename T>
__global__ void k_to_host(const T *__restrict__ src, T *__restrict__ dst, int64_t cnt) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < cnt) dst[i] = src[i];
}
template <typename T> static void to_host(const T *src, T *dst_mapped, int64_t cnt, cudaStream_t st) {
  if (cnt > 0) k_to_host<T><<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(src, dst_mapped, cnt);
}

static int pipe_host_staging(ctp_map *m, int64_t qc) {
  if (m->h_off_cap >= 2 * (qc + 1)) return CTP_OK;
  if (m->h_off) cudaFreeHost(m->h_off);
  if (m->h_nf) cudaFreeHost(m->h_nf);
  m->h_off = nullptr;
  m->h_nf = nullptr;
  m->h_off_cap = 0;
  CK(cudaHostAlloc((void **)&m->h_off, 2 * (qc + 1) * sizeof(int64_t), cudaHostAllocMapped));
  // accept counts after the rejection (2 buffers), n_used (2), accept counts after the build (2)
  CK(cudaHostAlloc((void **)&m->h_nf, 6 * (qc + 1) * sizeof(int32_t), cudaHostAllocMapped));
  {  // the kernels write through the host pointers: that needs unified addressing (every 64-bit CUDA platform has it)
    void *d0 = nullptr, *d1 = nullptr;
    CK(cudaHostGetDevicePointer(&d0, m->h_off, 0));
    CK(cudaHostGetDevicePointer(&d1, m->h_nf, 0));
    if (d0 != (void *)m->h_off || d1 != (void *)m->h_nf)
      return fail(CTP_ERR_CUDA, "mapped page-locked memory is not addressable by its host pointer on this device");
  }
  m->h_off_cap = 2 * (qc + 1);
  return CTP_OK;
}

// Shared front half of the chunk pipelines (ctp_sample_multiple, ctp_query_paths): candidates of a
// chunk go up in two phases.  The sequential do/while of the reference stops at the (n-2)-th accepted
// candidate, so the tail of a stream is normally never looked at: the first `head` candidates per query
// are copied and tested; only when some query of the chunk is still short (known on the host from a
// 4-byte-per-query read back) the rest of the chunk's streams follows and the rejection is redone over
// the whole stream.  Outputs are identical either way.
struct PipeFront {
  double *d_s, *d_g, *d_c, *d_n;
  int32_t *d_nf, *d_nu;
  int32_t *d_src = nullptr;  // optional: candidate index of every accepted node (CTP_NODES_AS_INDEX)
  int64_t head = 0;  // candidates per query uploaded up front for the chunk now in this buffer
};

// How much of each stream goes up front: the configured multiple of n for the first chunk, afterwards what the
// chunks seen so far needed (largest n_used + 6 %), so a map with a low acceptance rate stops paying refills
// and one with a high rate stops uploading candidates nobody looks at.
static int64_t pipe_head(const ctp_map *m, int64_t n, int64_t cnt) {
  if (m->pipe_head_factor <= 0) return cnt;
  int64_t head = (int64_t)(m->pipe_head_factor * (double)n) + 256;
  if (m->pipe_nu_seen > 0) head = m->pipe_nu_seen + m->pipe_nu_seen / 16 + 256;
  return head >= cnt ? cnt : head;
}

// Queries per chunk: about `edges` directed edges of work per chunk -- but a handful of very large queries (C3: 4
// queries of 1.3 M edges) would then share one chunk and nothing would overlap, so the batch is cut into at least
// four chunks as long as a chunk keeps a million edges of work.
static int64_t pipe_auto_chunk(int64_t q, int64_t n, int k, int64_t edges) {
  const int64_t per_query = std::max<int64_t>(1, n * k);
  int64_t qc = std::max<int64_t>(1, edges / per_query);
  const int64_t quarter = (q + 3) / 4, floor_q = ((int64_t)1 << 20) / per_query + 1;
  return std::max<int64_t>(1, std::min(qc, std::max(quarter, floor_q)));
}

// stage A: copy the head of the chunk's streams, run the rejection over it, start the n_filled read back
static int pipe_stage_a(ctp_map *m, PipeFront &b, int bi, const double *start, const double *goal, const double *cand,
                        int64_t q0, int64_t qn, int64_t cnt, int64_t n, double clr, bool wait_reuse, int32_t *h_nf,
                        int32_t *h_nu) {
  cudaStream_t s_in = m->pipe_st[0], s_comp = m->pipe_st[1];
  const int64_t head = pipe_head(m, n, cnt);
  b.head = head;
  if (wait_reuse) CK(cudaStreamWaitEvent(s_in, m->ev_comp[bi], 0));  // the staging buffer's last reader
  CK(cudaMemcpyAsync(b.d_s, start + 3 * q0, qn * 24, cudaMemcpyHostToDevice, s_in));
  CK(cudaMemcpyAsync(b.d_g, goal + 3 * q0, qn * 24, cudaMemcpyHostToDevice, s_in));
  if (head)
    CK(cudaMemcpy2DAsync(b.d_c, cnt * 24, cand + 3 * q0 * cnt, cnt * 24, head * 24, qn, cudaMemcpyHostToDevice, s_in));
  m->io_h2d += (unsigned long long)qn * (48 + head * 24);
  m->io_d2h += (unsigned long long)qn * 8;
  CK(cudaEventRecord(m->ev_in[bi], s_in));
  CK(cudaStreamWaitEvent(s_comp, m->ev_in[bi], 0));
  if (wait_reuse) CK(cudaStreamWaitEvent(s_comp, m->ev_out[bi], 0));  // outputs of the buffer's previous chunk have left
  m->ws.used = 0;
  CK(cudaMemsetAsync(b.d_n, 0, qn * n * 24, s_comp));  // rows of a query that runs out of candidates stay defined
  if (b.d_src) CK(cudaMemsetAsync(b.d_src, 0xff, qn * n * 4, s_comp));  // -1 = no node
  CKR(sample_reject_launch(m, b.d_s, b.d_g, b.d_c, qn, head, cnt, clr, n, b.d_n, b.d_nf, b.d_nu, nullptr, s_comp, b.d_src));
  to_host(b.d_nf, h_nf, qn, s_comp);
  to_host(b.d_nu, h_nu, qn, s_comp);
  m->launches += 2;
  CK(cudaGetLastError());
  CK(cudaEventRecord(m->ev_rej[bi], s_comp));
  return CTP_OK;
}

// start of stage B: wait for the head's verdict; the queries that are still short get the rest of their streams
// (runs of consecutive short queries share one strided copy) and the rejection is redone over the chunk -- a query
// that was not short takes its first n accepted candidates from the head either way, so what lies behind its
// head (left over from an earlier chunk) does not matter
static int pipe_stage_b_refill(ctp_map *m, const PipeFront &b, int bi, const double *cand, int64_t q0, int64_t qn,
                               int64_t cnt, int64_t n, double clr, const int32_t *h_nf, const int32_t *h_nu) {
  cudaStream_t s_in = m->pipe_st[0], s_comp = m->pipe_st[1];
  CK(cudaEventSynchronize(m->ev_rej[bi]));
  const int64_t head = b.head;
  if (head >= cnt) return CTP_OK;
  bool is_short = false;
  int64_t nu_max = 0;
  for (int64_t i = 0; i < qn; i++) {
    is_short |= h_nf[i] < n;
    nu_max = std::max<int64_t>(nu_max, h_nu[i]);
  }
  if (is_short) nu_max = std::max(nu_max, std::min(cnt, head + head / 4));  // needed more than the head: guess 25 % more
  m->pipe_nu_seen = std::max(m->pipe_nu_seen, nu_max);
  if (!is_short) return CTP_OK;
  for (int64_t i = 0; i < qn;) {
    if (h_nf[i] >= n) { i++; continue; }
    int64_t j = i + 1;
    while (j < qn && h_nf[j] < n) j++;
    CK(cudaMemcpy2DAsync(b.d_c + 3 * (i * cnt + head), cnt * 24, cand + 3 * ((q0 + i) * cnt + head), cnt * 24,
                         (cnt - head) * 24, j - i, cudaMemcpyHostToDevice, s_in));
    m->io_h2d += (unsigned long long)(j - i) * (cnt - head) * 24;
    i = j;
  }
  CK(cudaEventRecord(m->ev_in[bi], s_in));
  CK(cudaStreamWaitEvent(s_comp, m->ev_in[bi], 0));
  m->ws.used = 0;
  CK(cudaMemsetAsync(b.d_n, 0, qn * n * 24, s_comp));
  if (b.d_src) CK(cudaMemsetAsync(b.d_src, 0xff, qn * n * 4, s_comp));
  CKR(sample_reject_launch(m, b.d_s, b.d_g, b.d_c, qn, cnt, cnt, clr, n, b.d_n, b.d_nf, b.d_nu, nullptr, s_comp, b.d_src));
  return CTP_OK;
}

extern "C" int ctp_sample_multiple(ctp_map *m, const double *start, const double *goal, const double *cand,
                                   int64_t q, int64_t cnt, int64_t n, int k, double clr, double cdc, double *nodes,
                                   int32_t *n_filled, int32_t *row_ptr, int32_t *col, double *w, int64_t cap,
                                   int64_t *nnz_off) {
  return ctp_sample_multiple_ex(m, start, goal, cand, q, cnt, n, k, clr, cdc, 0, nodes, n_filled, row_ptr, col, w, cap, nnz_off);
}

extern "C" int ctp_sample_multiple_ex(ctp_map *m, const double *start, const double *goal, const double *cand,
                                      int64_t q, int64_t cnt, int64_t n, int k, double clr, double cdc, int csr_flags,
                                      double *nodes, int32_t *n_filled, int32_t *row_ptr, int32_t *col, double *w,
                                      int64_t cap, int64_t *nnz_off) {
  REQUIRE(m && q >= 1 && n >= 2 && cnt >= 0 && k >= 1 && k <= CTP_KNN_MAXK, "bad argument");
  REQUIRE((csr_flags & ~(CTP_CSR_UPPER | CTP_CSR_NO_WEIGHTS | CTP_CSR_COL16 | CTP_NODES_AS_INDEX)) == 0, "unknown csr_flags bits");
  const bool want_w = !(csr_flags & CTP_CSR_NO_WEIGHTS);
  const bool col16 = (csr_flags & CTP_CSR_COL16) != 0, node_idx = (csr_flags & CTP_NODES_AS_INDEX) != 0;
  REQUIRE(!col16 || n <= 65536, "CTP_CSR_COL16 needs n <= 65536");
  const int build_flags = csr_flags & (CTP_CSR_UPPER | CTP_CSR_NO_WEIGHTS);
  REQUIRE(start && goal && cand && nodes && n_filled && row_ptr && col && (w || !want_w) && nnz_off, "null buffer");
  REQUIRE(cdc > 0, "collision_distance_check must be > 0");
  CK(cudaSetDevice(m->device));
  CKR(pipe_init(m));
  // chunk size: about 16M directed edges per chunk, at least 1 query, at most 65535 (grid.y limits).  (C2, 512 queries:
  // chunks of 16 / 32 / 64 / 128 queries give 3.4 / 4.6 / 5.6 / 6.0 G checks/s end to end -- every chunk costs the host
  // a synchronisation on its accept counts and the device a drained compute stream.)
  int64_t qc = m->pipe_chunk > 0 ? m->pipe_chunk : pipe_auto_chunk(q, n, k, 16 << 20);
  qc = std::min<int64_t>(std::min<int64_t>(qc, q), 65535);
  REQUIRE(qc * n * k < ((int64_t)1 << 31), "n*k too large for one chunk");
  // chunk boundaries: a short first chunk gets the device->host stream going early, the rest are qc long
  std::vector<int64_t> cs;  // chunk c covers queries [cs[c], cs[c+1])
  cs.push_back(0);
  if (q > qc) cs.push_back(std::max<int64_t>(1, qc / 4));
  while (cs.back() < q) cs.push_back(std::min<int64_t>(q, cs.back() + qc));
  const int64_t n_chunks = (int64_t)cs.size() - 1;
  const int nbuf = n_chunks > 1 ? 2 : 1;
  m->pipe_nu_seen = 0;  // the up-front share of the streams adapts within this call only
  const int64_t ecap = ((csr_flags & CTP_CSR_UPPER) ? 1 : 2) * qc * n * k;  // CSR entries a chunk can produce
  size_t per = 2 * align_up(qc * 24) + align_up(qc * cnt * 24) + align_up(qc * n * 24) + 2 * align_up(qc * 4) +
               align_up(qc * (n + 1) * 4) + align_up(ecap * 4) + (want_w ? align_up(ecap * 8) : 0) + align_up((qc + 1) * 8) +
               (node_idx ? align_up(qc * n * 4) : 0) + (col16 ? align_up(ecap * 2) : 0);
  CKR(arena_reserve(m->io, nbuf * per + 4096));
  CKR(ctp_reserve(m, qc, n, cnt, k));
  CKR(pipe_host_staging(m, qc));
  IoScope sc(m);
  PipeFront F[2];
  struct { int32_t *d_rp, *d_col; double *d_w; int64_t *d_off; uint16_t *d_col16; } B[2];
  for (int i = 0; i < nbuf; i++) {
    F[i].d_s = arena_take<double>(m->io, qc * 3);
    F[i].d_g = arena_take<double>(m->io, qc * 3);
    F[i].d_c = arena_take<double>(m->io, qc * cnt * 3);
    F[i].d_n = arena_take<double>(m->io, qc * n * 3);
    F[i].d_nf = arena_take<int32_t>(m->io, qc);
    F[i].d_nu = arena_take<int32_t>(m->io, qc);
    F[i].d_src = node_idx ? arena_take<int32_t>(m->io, qc * n) : nullptr;
    B[i].d_rp = arena_take<int32_t>(m->io, qc * (n + 1));
    B[i].d_col = arena_take<int32_t>(m->io, ecap);
    B[i].d_w = want_w ? arena_take<double>(m->io, ecap) : nullptr;
    B[i].d_off = arena_take<int64_t>(m->io, qc + 1);
    B[i].d_col16 = col16 ? arena_take<uint16_t>(m->io, ecap) : nullptr;
  }
  cudaStream_t s_comp = m->pipe_st[1], s_out = m->pipe_st[2], s_small = m->pipe_st[3];
  int64_t base = 0;     // CSR entries emitted by the chunks finished so far
  int rc_deferred = CTP_OK;
  nnz_off[0] = 0;
  // CTP_PIPE_TRACE=1: a timeline of the call on stderr (device time stamps of every chunk's phases)
  struct TraceMark { const char *what; int64_t chunk; cudaEvent_t ev; };
  std::vector<TraceMark> trace;
  const bool tracing = getenv("CTP_PIPE_TRACE") != nullptr;
  auto mark = [&](const char *what, int64_t c, cudaStream_t s) {
    if (!tracing) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    trace.push_back({what, c, e});
  };
  mark("call start", -1, m->pipe_st[0]);
  // last part of a chunk: once its offsets are on the host, copy exactly nnz col/w entries out
  auto finish = [&](int64_t c) -> int {
    const int bi = (int)(c % nbuf);
    const int64_t q0 = cs[c], qn = cs[c + 1] - q0;
    CK(cudaEventSynchronize(m->ev_small[bi]));
    const int64_t *ho = m->h_off + (size_t)bi * (qc + 1);
    memcpy(n_filled + q0, m->h_nf + (size_t)(4 + bi) * (qc + 1), (size_t)qn * 4);
    const int64_t nnz = ho[qn];
    for (int64_t i = 1; i <= qn; i++) nnz_off[q0 + i] = base + ho[i];
    if (base + nnz > cap) {
      if (rc_deferred == CTP_OK)
        rc_deferred = fail(CTP_ERR_CAPACITY, "CSR needs more than %lld entries", (long long)cap);
    } else if (nnz) {
      if (col16) CK(cudaMemcpyAsync((uint16_t *)col + base, B[bi].d_col16, nnz * 2, cudaMemcpyDeviceToHost, s_out));
      else CK(cudaMemcpyAsync(col + base, B[bi].d_col, nnz * 4, cudaMemcpyDeviceToHost, s_out));
      if (want_w) CK(cudaMemcpyAsync(w + base, B[bi].d_w, nnz * 8, cudaMemcpyDeviceToHost, s_out));
      m->io_d2h += (unsigned long long)nnz * ((want_w ? 8 : 0) + (col16 ? 2 : 4));
    }
    mark("d2h end", c, s_out);
    CK(cudaEventRecord(m->ev_out[bi], s_out));
    base += nnz;
    return CTP_OK;
  };
  // stage B of chunk c: (refill,) build, fixed-size outputs; the previous chunk's col/w leave first
  auto stage_b = [&](int64_t c) -> int {
    const int bi = (int)(c % nbuf);
    const int64_t q0 = cs[c], qn = cs[c + 1] - q0;
    CKR(pipe_stage_b_refill(m, F[bi], bi, cand, q0, qn, cnt, n, clr, m->h_nf + (size_t)bi * qc,
                            m->h_nf + (size_t)(2 + bi) * (qc + 1)));
    mark("build start", c, s_comp);
    CKR(ctp_build_prm_ex_dev(m, F[bi].d_n, qn, n, k, clr, cdc, build_flags, nullptr, nullptr, B[bi].d_rp, B[bi].d_col,
                             B[bi].d_w, ecap, B[bi].d_off, s_comp));
    if (col16) {  // the chunk's whole column buffer (its nnz is not known on the host yet; what lies beyond is never copied)
      const int64_t ce = ((csr_flags & CTP_CSR_UPPER) ? 1 : 2) * qn * n * k;
      k_narrow_u16<<<(unsigned)std::min<int64_t>((ce + 1023) / 1024, 65535 * 4), 256, 0, s_comp>>>(B[bi].d_col, ce, B[bi].d_col16);
      m->launches++;
    }
    mark("build end", c, s_comp);
    to_host(B[bi].d_off, m->h_off + (size_t)bi * (qc + 1), qn + 1, s_comp);
    to_host(F[bi].d_nf, m->h_nf + (size_t)(4 + bi) * (qc + 1), qn, s_comp);
    m->launches += 2;
    CK(cudaGetLastError());
    CK(cudaEventRecord(m->ev_comp[bi], s_comp));
    CK(cudaEventRecord(m->ev_small[bi], s_comp));
    if (c >= 1) CKR(finish(c - 1));
    CK(cudaStreamWaitEvent(s_out, m->ev_comp[bi], 0));
    mark("d2h start", c, s_out);
    if (node_idx) CK(cudaMemcpyAsync((int32_t *)nodes + q0 * n, F[bi].d_src, qn * n * 4, cudaMemcpyDeviceToHost, s_out));
    else CK(cudaMemcpyAsync(nodes + 3 * q0 * n, F[bi].d_n, qn * n * 24, cudaMemcpyDeviceToHost, s_out));
    CK(cudaMemcpyAsync(row_ptr + q0 * (n + 1), B[bi].d_rp, qn * (n + 1) * 4, cudaMemcpyDeviceToHost, s_out));
    m->io_d2h += (unsigned long long)qn * (n * (node_idx ? 4 : 24) + (n + 1) * 4 + 4) + (unsigned long long)(qn + 1) * 8;
    return CTP_OK;
  };
  for (int64_t c = 0; c <= n_chunks; c++) {
    // B(c-1) before A(c): A(c) reuses the buffers of chunk c-2, whose last events B(c-1) has just recorded
    if (c >= 1) CKR(stage_b(c - 1));
    if (c < n_chunks) {
      const int bi = (int)(c % nbuf);
      const int64_t q0 = cs[c], qn = cs[c + 1] - q0;
      mark("h2d start", c, m->pipe_st[0]);
      CKR(pipe_stage_a(m, F[bi], bi, start, goal, cand, q0, qn, cnt, n, clr, c >= nbuf, m->h_nf + (size_t)bi * qc,
                       m->h_nf + (size_t)(2 + bi) * (qc + 1)));
      mark("h2d end", c, m->pipe_st[0]);
      mark("reject end", c, s_comp);
    }
  }
  CKR(finish(n_chunks - 1));
  CK(cudaStreamSynchronize(s_out));
  CK(cudaStreamSynchronize(s_small));
  CK(cudaStreamSynchronize(s_comp));
  if (tracing) {
    for (const TraceMark &tm : trace) {
      float ms = 0;
      cudaEventElapsedTime(&ms, trace[0].ev, tm.ev);
      fprintf(stderr, "[ctp pipe] %8.3f ms  chunk %2lld  %s\n", ms, (long long)tm.chunk, tm.what);
    }
    for (const TraceMark &tm : trace) cudaEventDestroy(tm.ev);
  }
  if (rc_deferred != CTP_OK) return rc_deferred;
  for (int64_t i = 0; i < q; i++)
    if (n_filled[i] < n)
      return fail(CTP_ERR_NEED_CANDIDATES, "query %lld: only %d of %lld nodes accepted (candidate stream exhausted)",
                  (long long)i, n_filled[i], (long long)n);
  return CTP_OK;
}

// PRM build + shortest start->goal path per query, host candidate streams in, host paths out.  Same
// chunk pipeline as ctp_sample_multiple, but the roadmaps never leave the device: per query only the
// path (ids, coordinates), its length and the accepted-node count come back.
extern "C" int ctp_query_paths(ctp_map *m, const double *start, const double *goal, const double *cand, int64_t q,
                               int64_t cnt, int64_t n, int k, double clr, double cdc, int32_t path_cap, int32_t *path_ids,
                               double *path_xyz, int32_t *path_n, double *length, int32_t *n_filled) {
  REQUIRE(m && q >= 1 && n >= 2 && cnt >= 0 && k >= 1 && k <= CTP_KNN_MAXK && path_cap >= 2, "bad argument");
  REQUIRE(start && goal && cand && path_ids && path_xyz && path_n && length && n_filled, "null buffer");
  REQUIRE(cdc > 0, "collision_distance_check must be > 0");
  CK(cudaSetDevice(m->device));
  CKR(pipe_init(m));
  // larger chunks than ctp_sample_multiple: the per-query CTAs of the graph search need about a machine's worth of
  // queries per launch, and there is no large device->host copy to hide
  int64_t qc = m->pipe_chunk > 0 ? m->pipe_chunk : pipe_auto_chunk(q, n, k, 16 << 20);
  qc = std::min<int64_t>(std::min<int64_t>(qc, q), 65535);
  REQUIRE(qc * n * k < ((int64_t)1 << 31), "n*k too large for one chunk");
  const int64_t n_chunks = (q + qc - 1) / qc;
  const int nbuf = n_chunks > 1 ? 2 : 1;
  m->pipe_nu_seen = 0;  // the up-front share of the streams adapts within this call only
  const int64_t ecap = 2 * qc * n * k;
  // inputs, nodes and the small outputs are double-buffered; CSR and distances are reused chunk after chunk
  const size_t per = 2 * align_up(qc * 24) + align_up(qc * cnt * 24) + align_up(qc * n * 24) + 2 * align_up(qc * 4) +
                     align_up(qc * path_cap * 4) + align_up(qc * path_cap * 24) + align_up(qc * 4) + align_up(qc * 8);
  const size_t shared = align_up(qc * (n + 1) * 4) + align_up(ecap * 4) + align_up(ecap * 8) + align_up((qc + 1) * 8) +
                        align_up(qc * n * 8) + align_up(qc * n * 4) + align_up(qc * 4);
  CKR(arena_reserve(m->io, nbuf * per + shared + 8192));
  CKR(ctp_reserve(m, qc, n, cnt, k));
  CKR(pipe_host_staging(m, qc));
  IoScope sc(m);
  PipeFront F[2];
  struct { int32_t *d_pid; double *d_pxyz; int32_t *d_pn; double *d_len; } B[2];
  for (int i = 0; i < nbuf; i++) {
    F[i].d_s = arena_take<double>(m->io, qc * 3);
    F[i].d_g = arena_take<double>(m->io, qc * 3);
    F[i].d_c = arena_take<double>(m->io, qc * cnt * 3);
    F[i].d_n = arena_take<double>(m->io, qc * n * 3);
    F[i].d_nf = arena_take<int32_t>(m->io, qc);
    F[i].d_nu = arena_take<int32_t>(m->io, qc);
    B[i].d_pid = arena_take<int32_t>(m->io, qc * path_cap);
    B[i].d_pxyz = arena_take<double>(m->io, qc * path_cap * 3);
    B[i].d_pn = arena_take<int32_t>(m->io, qc);
    B[i].d_len = arena_take<double>(m->io, qc);
  }
  int32_t *d_rp = arena_take<int32_t>(m->io, qc * (n + 1));
  int32_t *d_col = arena_take<int32_t>(m->io, ecap);
  double *d_w = arena_take<double>(m->io, ecap);
  int64_t *d_off = arena_take<int64_t>(m->io, qc + 1);
  double *d_dist = arena_take<double>(m->io, qc * n);
  int32_t *d_pred = arena_take<int32_t>(m->io, qc * n);
  int32_t *d_src = arena_take<int32_t>(m->io, qc);
  cudaStream_t s_comp = m->pipe_st[1], s_out = m->pipe_st[2];
  CK(cudaMemsetAsync(d_src, 0, qc * 4, s_comp));  // every search starts at node 0 (= start, :299-306)
  auto stage_b = [&](int64_t c) -> int {
    const int bi = (int)(c % nbuf);
    const int64_t q0 = c * qc, qn = std::min(qc, q - q0);
    CKR(pipe_stage_b_refill(m, F[bi], bi, cand, q0, qn, cnt, n, clr, m->h_nf + (size_t)bi * qc,
                            m->h_nf + (size_t)(2 + bi) * (qc + 1)));
    CKR(ctp_build_prm_dev(m, F[bi].d_n, qn, n, k, clr, cdc, nullptr, nullptr, d_rp, d_col, d_w, ecap, d_off, s_comp));
    CKR(ctp_sssp_dev(m, qn, n, d_rp, d_col, d_w, d_off, d_src, 1, d_dist, d_pred, nullptr, s_comp));
    k_extract_paths<<<(int)((qn + 127) / 128), 128, 0, s_comp>>>(qn, n, F[bi].d_n, d_dist, d_pred, 0, 1, path_cap,
                                                                 B[bi].d_pid, B[bi].d_pxyz, B[bi].d_pn, B[bi].d_len);
    m->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(m->ev_comp[bi], s_comp));
    CK(cudaStreamWaitEvent(s_out, m->ev_comp[bi], 0));
    CK(cudaMemcpyAsync(path_ids + q0 * path_cap, B[bi].d_pid, qn * path_cap * 4, cudaMemcpyDeviceToHost, s_out));
    CK(cudaMemcpyAsync(path_xyz + 3 * q0 * path_cap, B[bi].d_pxyz, qn * path_cap * 24, cudaMemcpyDeviceToHost, s_out));
    CK(cudaMemcpyAsync(path_n + q0, B[bi].d_pn, qn * 4, cudaMemcpyDeviceToHost, s_out));
    CK(cudaMemcpyAsync(length + q0, B[bi].d_len, qn * 8, cudaMemcpyDeviceToHost, s_out));
    CK(cudaMemcpyAsync(n_filled + q0, F[bi].d_nf, qn * 4, cudaMemcpyDeviceToHost, s_out));
    m->io_d2h += (unsigned long long)qn * ((unsigned long long)path_cap * 28 + 16);
    CK(cudaEventRecord(m->ev_out[bi], s_out));
    return CTP_OK;
  };
  for (int64_t c = 0; c <= n_chunks; c++) {
    if (c >= 1) CKR(stage_b(c - 1));
    if (c < n_chunks) {
      const int bi = (int)(c % nbuf);
      const int64_t q0 = c * qc, qn = std::min(qc, q - q0);
      CKR(pipe_stage_a(m, F[bi], bi, start, goal, cand, q0, qn, cnt, n, clr, c >= nbuf, m->h_nf + (size_t)bi * qc,
                       m->h_nf + (size_t)(2 + bi) * (qc + 1)));
    }
  }
  CK(cudaStreamSynchronize(s_out));
  CK(cudaStreamSynchronize(m->pipe_st[3]));
  CK(cudaStreamSynchronize(s_comp));
  for (int64_t i = 0; i < q; i++)
    if (n_filled[i] < n)
      return fail(CTP_ERR_NEED_CANDIDATES, "query %lld: only %d of %lld nodes accepted (candidate stream exhausted)",
                  (long long)i, n_filled[i], (long long)n);
  return CTP_OK;
}

#include "ctp_paths_api.inl"
