// ctp_prm.cuh -- the PRM-construction stages of TopologicalPRMClustering::sampleMultiple
// (include/topological_prm_clustering.hpp:288-401): order-preserving rejection sampling,
// exact float32 k-NN on a uniform grid, symmetric CSR emit.  All kernels are batched over
// q independent queries (blockIdx.y or a flat q*n index).
#pragma once
#include <cooperative_groups.h>
#include "ctp_device.cuh"

// =========================== rejection sampling (:299-327) ================================
// The reference's sequential do/while keeps the first n-2 candidates that are not in
// collision.  Evaluate every candidate, then compact in stream order.

#define CTP_REJ_BLOCK 256

// pass 1: accept bit per candidate + per-block accept count
template <int L>
__global__ void __launch_bounds__(CTP_REJ_BLOCK)
k_reject_flags(MapDev M, const double *__restrict__ cand, int64_t m, int64_t stride, double clr,
               uint8_t *__restrict__ accept, int32_t *__restrict__ block_cnt, int nblk) {
  // candidate i of query q sits at cand[q * stride + i]; only the first m of each stream are looked at
  const int64_t q = blockIdx.y;
  const int64_t i = blockIdx.x * (int64_t)CTP_REJ_BLOCK + threadIdx.x;
  bool ok = false;
  if (i < m) {
    const double *p = cand + 3 * (q * stride + i);
    ok = !ctp_collides(esdf_clearance<L>(M, p[0], p[1], p[2]), clr);
    accept[q * m + i] = ok ? 1 : 0;
  }
  __shared__ int wsum[CTP_REJ_BLOCK / 32];
  const unsigned b = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < CTP_REJ_BLOCK / 32; w++) s += wsum[w];
    block_cnt[q * nblk + blockIdx.x] = s;
  }
}

// pass 2: one CTA per query, exclusive scan of the block counts (in place)
__global__ void __launch_bounds__(1024) k_reject_scan(int32_t *__restrict__ block_cnt, int nblk,
                                                      int64_t n, int32_t *__restrict__ n_filled) {
  __shared__ int wsum[32];
  __shared__ int carry;
  int32_t *bc = block_cnt + (int64_t)blockIdx.x * nblk;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblk ? bc[i] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) >= o) s += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = wsum[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += t;
      }
      wsum[threadIdx.x] = w;
    }
    __syncthreads();
    const int woff = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0;
    const int c = carry;
    if (i < nblk) bc[i] = c + woff + s - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + woff + s;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int64_t tot = carry;
    n_filled[blockIdx.x] = (int32_t)((tot < n - 2 ? tot : n - 2) + 2);
  }
}

// pass 3: scatter accepted candidates to nodes[2 + rank]; rows 0,1 = start, goal (:299-306)
__global__ void __launch_bounds__(CTP_REJ_BLOCK)
k_reject_scatter(const double *__restrict__ cand, const double *__restrict__ start,
                 const double *__restrict__ goal, const uint8_t *__restrict__ accept,
                 const int32_t *__restrict__ block_off, int nblk, int64_t m, int64_t stride, int64_t n,
                 double *__restrict__ nodes, int32_t *__restrict__ n_used, int32_t *__restrict__ src_idx) {
  const int64_t q = blockIdx.y;
  const int64_t i = blockIdx.x * (int64_t)CTP_REJ_BLOCK + threadIdx.x;
  const bool ok = i < m && accept[q * m + i];
  __shared__ int wsum[CTP_REJ_BLOCK / 32];
  const unsigned b = __ballot_sync(0xffffffffu, ok);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) wsum[w] = __popc(b);
  __syncthreads();
  int off = block_off[q * nblk + blockIdx.x];
  for (int j = 0; j < w; j++) off += wsum[j];
  const int64_t rank = off + __popc(b & ((1u << lane) - 1u));
  double *nq = nodes + 3 * q * n;
  if (ok && rank < n - 2) {
    const double *p = cand + 3 * (q * stride + i);
    double *o = nq + 3 * (2 + rank);
    o[0] = p[0];
    o[1] = p[1];
    o[2] = p[2];
    if (src_idx) src_idx[q * n + 2 + rank] = (int32_t)i;  // which candidate of the stream this node is
    if (rank == n - 3) n_used[q] = (int32_t)(i + 1);  // the do/while stops right here
  }
  if (blockIdx.x == 0 && threadIdx.x < 3) {
    nq[threadIdx.x] = start[3 * q + threadIdx.x];
    nq[3 + threadIdx.x] = goal[3 * q + threadIdx.x];
  }
  if (src_idx && blockIdx.x == 0 && threadIdx.x < 2) src_idx[q * n + threadIdx.x] = -2 - (int)threadIdx.x;  // start, goal
}

// =========================== exact float32 k-NN (:329-339) ================================
// FLANN's L2<float> sees the points as float (:300-326) and accumulates diff*diff in float,
// axis order x,y,z.  Restated exactly (no FMA) with (d2, index) tie-break; self excluded.

#define CTP_KNN_MAXK 16

struct KnnGrid {   // per-query uniform grid, filled by k_knn_setup
  float lox, loy, loz;
  float inv_h;
  double h;
  int dx, dy, dz;
  int ncell;
};

__device__ __forceinline__ int knn_cell_1d(float x, float lo, float inv_h, int dim) {
  int c = (int)floorf((x - lo) * inv_h);
  return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}

// one CTA per query: float bounding box, a coarse occupancy bitmap to estimate the volume the
// points actually fill (an ellipsoid clipped by the map and by obstacles fills a third of its
// box or less), then a cell size h such that a ball of radius h is expected to hold `in_ball`
// points -- the quantity the tile search's completeness test depends on.
// CL > 1: one thread-block cluster of CL CTAs per query (few queries of many points would otherwise leave this kernel
// on a handful of SMs): each CTA takes a strided share of the points, boxes and occupancy bitmaps meet through
// distributed shared memory.
template <int CL>
__global__ void __launch_bounds__(1024) k_knn_setup(const double *__restrict__ nodes, int64_t n,
                                                    int max_cells, double in_ball,
                                                    KnnGrid *__restrict__ grids, float4 *__restrict__ pts) {
  namespace cg = cooperative_groups;
  const int64_t q = blockIdx.x / CL;
  const int rank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
  const int64_t first = (int64_t)rank * blockDim.x + threadIdx.x, stride = (int64_t)CL * blockDim.x;
  const double *nq = nodes + 3 * q * n;
  float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int64_t i = first; i < n; i += stride) {
    const float x = (float)nq[3 * i], y = (float)nq[3 * i + 1], z = (float)nq[3 * i + 2];
    pts[q * n + i] = make_float4(x, y, z, __int_as_float((int)i));
    lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
    lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
    lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
  }
  __shared__ float s_lo[3][32], s_hi[3][32];
  __shared__ float s_box[6];  // this CTA's box, read by the other CTAs of the cluster
  __shared__ float s_blo[3], s_binv[3];
  __shared__ int s_bg[3];
  __shared__ unsigned s_occ[128];  // up to 16^3 coarse cells
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[a][threadIdx.x >> 5] = lo[a]; s_hi[a][threadIdx.x >> 5] = hi[a]; }
  }
  if (threadIdx.x < 128) s_occ[threadIdx.x] = 0;
  __syncthreads();
  const int nw = (blockDim.x + 31) / 32;
  for (int a = 0; a < 3; a++)
    for (int w = 0; w < nw; w++) { lo[a] = fminf(lo[a], s_lo[a][w]); hi[a] = fmaxf(hi[a], s_hi[a][w]); }
  if (CL > 1) {
    if (threadIdx.x == 0)
      for (int a = 0; a < 3; a++) { s_box[a] = lo[a]; s_box[3 + a] = hi[a]; }
    cg::this_cluster().sync();
    for (int r = 0; r < CL; r++) {
      const float *ob = cg::this_cluster().map_shared_rank(s_box, r);
      for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], ob[a]); hi[a] = fmaxf(hi[a], ob[3 + a]); }
    }
  }
  double e[3], emax = 0;
  for (int a = 0; a < 3; a++) {
    e[a] = (double)hi[a] - (double)lo[a];
    if (!(e[a] >= 0)) e[a] = 0;  // NaN coordinates
    emax = fmax(emax, e[a]);
  }
  if (!(emax > 0) || !(emax < 1e300)) emax = 1.0;
  const int G = n >= 4000 ? 16 : (n >= 500 ? 8 : (n >= 64 ? 4 : 2));
  if (threadIdx.x == 0) {
    for (int a = 0; a < 3; a++) {
      const bool flat = !(e[a] > emax * 1e-3);  // degenerate axis (planar: true => every z equal)
      s_bg[a] = flat ? 1 : G;
      s_blo[a] = lo[a];
      s_binv[a] = flat ? 0.f : (float)(G / e[a]);
    }
  }
  __syncthreads();
  for (int64_t i = first; i < n; i += stride) {
    const float4 p = pts[q * n + i];
    const int cx = min(max((int)((p.x - s_blo[0]) * s_binv[0]), 0), s_bg[0] - 1);
    const int cy = min(max((int)((p.y - s_blo[1]) * s_binv[1]), 0), s_bg[1] - 1);
    const int cz = min(max((int)((p.z - s_blo[2]) * s_binv[2]), 0), s_bg[2] - 1);
    const int c = (cz * s_bg[1] + cy) * s_bg[0] + cx;
    atomicOr(&s_occ[c >> 5], 1u << (c & 31));
  }
  __syncthreads();
  if (CL > 1) cg::this_cluster().sync();
  if (threadIdx.x == 0 && rank == 0) {
    int occ = 0;
    for (int w = 0; w < 128; w++) {
      unsigned bits = s_occ[w];
      if (CL > 1)
        for (int r = 1; r < CL; r++) bits |= cg::this_cluster().map_shared_rank(s_occ, r)[w];
      occ += __popc(bits);
    }
    if (occ < 1) occ = 1;
    // D-dimensional measure of the occupied coarse cells and the radius whose D-ball holds in_ball points
    int D = 0;
    double vol = (double)occ;
    for (int a = 0; a < 3; a++)
      if (s_bg[a] > 1) { D++; vol *= e[a] / s_bg[a]; }
    double h;
    const double per_point = vol / (double)n;
    if (D == 3) h = cbrt(in_ball * per_point / 4.18879);
    else if (D == 2) h = sqrt(in_ball * per_point / 3.14159);
    else if (D == 1) h = in_ball * per_point / 2.0;
    else h = emax;
    if (!(h > emax * 1e-7)) h = emax * 1e-7;
    int d[3];
    for (int it = 0; it < 256; it++) {
      long long prod = 1;
      for (int a = 0; a < 3; a++) {
        double c = floor(e[a] / h);
        d[a] = c < 1 ? 1 : (c > 1024 ? 1024 : (int)c);
        prod *= d[a];
      }
      if (prod <= max_cells) break;
      h *= 1.1;
    }
    KnnGrid g;
    g.lox = lo[0]; g.loy = lo[1]; g.loz = lo[2];
    g.h = h;
    g.inv_h = (float)(1.0 / h);
    g.dx = d[0]; g.dy = d[1]; g.dz = d[2];
    g.ncell = d[0] * d[1] * d[2];
    grids[q] = g;
  }
  if (CL > 1) cg::this_cluster().sync();  // rank 0 reads the other CTAs' bitmaps: nobody leaves before it is done
}

__global__ void __launch_bounds__(256) k_knn_count(const float4 *__restrict__ pts, int64_t n, int64_t total,
                                                   const KnnGrid *__restrict__ grids, int cell_stride,
                                                   int32_t *__restrict__ cell_of, int32_t *__restrict__ cell_cnt) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t q = t / n;
  const KnnGrid g = grids[q];
  const float4 p = pts[t];
  const int c = (knn_cell_1d(p.z, g.loz, g.inv_h, g.dz) * g.dy + knn_cell_1d(p.y, g.loy, g.inv_h, g.dy)) * g.dx +
                knn_cell_1d(p.x, g.lox, g.inv_h, g.dx);
  cell_of[t] = c;
  atomicAdd(cell_cnt + q * cell_stride + c, 1);
}

// exclusive scan of a 1024-wide tile: returns this thread's inclusive prefix; wsum[31] holds the tile total afterwards
__device__ __forceinline__ int cta_scan_1024(int v, int *wsum) {
  int s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) >= o) s += t;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = wsum[threadIdx.x];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (threadIdx.x >= o) w += t;
    }
    wsum[threadIdx.x] = w;
  }
  __syncthreads();
  return s + ((threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0);
}

// Exclusive scan of `len` values of one query by one CTA (CL == 1) or by a cluster of CL CTAs: each CTA scans a
// contiguous chunk, the chunk totals meet through distributed shared memory and a second sweep adds the offset.
// value(i) reads element i, store(i, prefix) writes it.  Returns the grand total (valid in every thread).
template <int CL, typename Value, typename Store, typename Add>
__device__ __forceinline__ long long cluster_excl_scan(int rank, int64_t len, Value value, Store store, Add add_offset) {
  namespace cg = cooperative_groups;
  __shared__ int wsum[32];
  __shared__ int carry;
  __shared__ int s_total;
  const int64_t chunk = CL > 1 ? ((len + CL - 1) / CL + 1023) / 1024 * 1024 : len;
  const int64_t lo = CL > 1 ? (rank * chunk < len ? rank * chunk : len) : 0;
  const int64_t hi = CL > 1 ? (lo + chunk < len ? lo + chunk : len) : len;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = lo; base < hi; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int v = i < hi ? value(i) : 0;
    const int incl = cta_scan_1024(v, wsum);
    const int c = carry;
    if (i < hi) store(i, c + incl - v);
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + incl;
    __syncthreads();
  }
  if (CL == 1) return carry;
  if (threadIdx.x == 0) s_total = carry;
  cg::this_cluster().sync();
  int off = 0, total = 0;
  for (int r = 0; r < CL; r++) {
    const int t = *cg::this_cluster().map_shared_rank(&s_total, r);
    off += r < rank ? t : 0;
    total += t;
  }
  cg::this_cluster().sync();  // every CTA has read the totals before anyone may exit
  if (off)
    for (int64_t i = lo + threadIdx.x; i < hi; i += 1024) add_offset(i, off);
  return total;
}

// exclusive scan of len counters (one problem): out[i] = sum of in[0..i), out[len] = total; the counters are zeroed
// (they become scatter cursors).  One CTA, or a cluster of CL.
template <int CL>
__global__ void __launch_bounds__(1024) k_excl_scan_i32(int32_t *__restrict__ in, int64_t len, int32_t *__restrict__ out) {
  const int rank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
  const long long total = cluster_excl_scan<CL>(
      rank, len, [&](int64_t i) { return in[i]; },
      [&](int64_t i, int pre) {
        out[i] = pre;
        in[i] = 0;
      },
      [&](int64_t i, int off) { out[i] += off; });
  if (threadIdx.x == 0 && rank == 0) out[len] = (int32_t)total;
}

// per query: exclusive scan of cell counts -> cell_start (ncell+1 entries); CL CTAs (a cluster) per query
template <int CL>
__global__ void __launch_bounds__(1024) k_knn_scan(const KnnGrid *__restrict__ grids, int cell_stride,
                                                   int32_t *__restrict__ cell_cnt, int32_t *__restrict__ cell_start) {
  const int64_t q = blockIdx.x / CL;
  const int rank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
  const int ncell = grids[q].ncell;
  int32_t *cnt = cell_cnt + q * cell_stride;
  int32_t *st = cell_start + q * (cell_stride + 1);
  const long long total = cluster_excl_scan<CL>(
      rank, ncell, [&](int64_t i) { return cnt[i]; },
      [&](int64_t i, int pre) {
        st[i] = pre;
        cnt[i] = 0;  // becomes the scatter cursor
      },
      [&](int64_t i, int off) { st[i] += off; });
  if (threadIdx.x == 0 && rank == 0) st[ncell] = (int32_t)total;
}

__global__ void __launch_bounds__(256) k_knn_scatter(const float4 *__restrict__ pts, int64_t n, int64_t total,
                                                     int cell_stride, const int32_t *__restrict__ cell_of,
                                                     const int32_t *__restrict__ cell_start,
                                                     int32_t *__restrict__ cursor, float4 *__restrict__ sorted) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t q = t / n;
  const int c = cell_of[t];
  const int pos = cell_start[q * (cell_stride + 1) + c] + atomicAdd(cursor + q * cell_stride + c, 1);
  sorted[q * n + pos] = pts[t];
}

// A slice [first, last) of the cell-sorted positions of query 0, moved to cell boundaries: out = {A, B} with
// A / B = the smallest cell start >= first / last.  The order of the points INSIDE a cell depends on the scatter's
// atomics, the cell starts do not: every GPU that built the grid from the same nodes gets the same A and B, so
// slices [shard(r), shard(r + 1)) of different GPUs hold disjoint sets of points that cover the sample.
__global__ void k_knn_slice_bounds(const KnnGrid *__restrict__ grids, const int32_t *__restrict__ cell_start, int64_t first,
                                   int64_t last, int64_t *__restrict__ out) {
  if (threadIdx.x > 1 || blockIdx.x) return;
  const int ncell = grids[0].ncell;
  const int64_t want = threadIdx.x ? last : first;
  int lo = 0, hi = ncell;  // cell_start[0 .. ncell] is non-decreasing, cell_start[ncell] = n
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cell_start[mid] >= want) hi = mid;
    else lo = mid + 1;
  }
  out[threadIdx.x] = cell_start[lo];
}

// k best (d2, index) pairs of a thread, ascending; one 64-bit key per pair (the bit pattern of a
// non-negative float orders like the number, so key order == (distance, index) order)
template <int K> struct TopK {
  unsigned long long key[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < K; s++) key[s] = 0x7f8000007fffffffull;  // (+inf, INT_MAX)
  }
  // ascending order of the K slots by Batcher's odd-even merge network for 16 inputs, with the exchanges that would
  // touch the absent inputs K..15 (+inf: they never move) left out at compile time
  __device__ __forceinline__ void sort() {
#pragma unroll
    for (int p = 1; p < 16; p <<= 1) {
#pragma unroll
      for (int kk = p; kk >= 1; kk >>= 1) {
#pragma unroll
        for (int a = 0; a < 16; a++) {  // exchange (a, a + kk) of the merge step, if the network has it
          const int b = a + kk, r = a - kk % p;
          if (r >= 0 && (r % (2 * kk)) < kk && b <= 15 && a / (2 * p) == b / (2 * p) && b < K) {
            const bool sw = key[b < K ? b : 0] < key[a < K ? a : 0];
            const unsigned long long lo = sw ? key[b < K ? b : 0] : key[a < K ? a : 0];
            const unsigned long long hi = sw ? key[a < K ? a : 0] : key[b < K ? b : 0];
            key[a < K ? a : 0] = lo;
            key[b < K ? b : 0] = hi;
          }
        }
      }
    }
  }
  __device__ __forceinline__ void push(float dd, int j) {
    push_key(((unsigned long long)__float_as_uint(dd) << 32) | (unsigned)j);
  }
  __device__ __forceinline__ void push_key(const unsigned long long k) {
    if (k < key[K - 1]) {
      key[K - 1] = k;
#pragma unroll
      for (int s = K - 1; s > 0; s--) {
        const bool sw = key[s] < key[s - 1];
        const unsigned long long lo = sw ? key[s] : key[s - 1], hi = sw ? key[s - 1] : key[s];
        key[s - 1] = lo;
        key[s] = hi;
      }
    }
  }
  __device__ __forceinline__ float d(int s) const { return __uint_as_float((unsigned)(key[s] >> 32)); }
  __device__ __forceinline__ int id(int s) const { return (int)(unsigned)key[s]; }
  __device__ __forceinline__ bool full() const { return id(K - 1) != 0x7fffffff; }
};

// one thread per (cell-sorted) point: expand Chebyshev rings of cells until the k-th best
// squared distance is strictly below the squared distance to the unvisited region.
template <int K>
__device__ __forceinline__ void knn_ring_search(int64_t t, const float4 *__restrict__ sorted, int64_t n,
                                                const KnnGrid *__restrict__ grids, int cell_stride,
                                                const int32_t *__restrict__ cell_start, int32_t *__restrict__ idx,
                                                int idx_stride, float *__restrict__ d2out) {
  const int64_t q = t / n;
  const KnnGrid g = grids[q];
  const float4 me = sorted[t];
  const int self = __float_as_int(me.w);
  const float4 *sq = sorted + q * n;
  const int32_t *st = cell_start + q * (cell_stride + 1);
  const int cx = knn_cell_1d(me.x, g.lox, g.inv_h, g.dx);
  const int cy = knn_cell_1d(me.y, g.loy, g.inv_h, g.dy);
  const int cz = knn_cell_1d(me.z, g.loz, g.inv_h, g.dz);
  TopK<K> best;
  best.init();
  const int rmax = max(max(g.dx, g.dy), g.dz);
  for (int r = 0; r <= rmax; r++) {
    const int z0 = cz - r, z1 = cz + r, y0 = cy - r, y1 = cy + r;
    for (int z = max(z0, 0); z <= min(z1, g.dz - 1); z++) {
      const bool zface = (z == z0) | (z == z1);
      for (int y = max(y0, 0); y <= min(y1, g.dy - 1); y++) {
        const bool face = zface | (y == y0) | (y == y1);
        // on a y/z face of the shell: the whole x run; otherwise only the two x end cells
        const int nseg = (face || r == 0) ? 1 : 2;
        for (int sgi = 0; sgi < nseg; sgi++) {
          int xa, xb;
          if (nseg == 1) { xa = cx - r; xb = cx + r; }
          else { xa = xb = (sgi == 0 ? cx - r : cx + r); }
          if (xb < 0 || xa >= g.dx) continue;
          xa = max(xa, 0);
          xb = min(xb, g.dx - 1);
          const int row = (z * g.dy + y) * g.dx;
          const int pb = st[row + xa], pe = st[row + xb + 1];
          for (int p = pb; p < pe; p++) {
            const float4 o = sq[p];
            const int j = __float_as_int(o.w);
            const float ddx = me.x - o.x, ddy = me.y - o.y, ddz = me.z - o.z;
            const float dd = (ddx * ddx + ddy * ddy) + ddz * ddz;
            if (j != self) best.push(dd, j);
          }
        }
      }
    }
    // everything outside the visited cube is farther than r*h along some axis (0.1% slack
    // covers float binning round-off); strict '<' so equal-distance ties are never skipped
    const double lb = (double)r * g.h * 0.999;
    if (best.full() && (double)best.d(K - 1) < lb * lb) break;
  }
  int32_t *oi = idx + (q * n + self) * idx_stride;
#pragma unroll
  for (int s = 0; s < K; s++) oi[s] = best.id(s) == 0x7fffffff ? -1 : best.id(s);
  if (d2out) {
    float *od = d2out + (q * n + self) * K;
#pragma unroll
    for (int s = 0; s < K; s++) od[s] = best.d(s);
  }
}

template <int K>
__global__ void __launch_bounds__(128) k_knn_search(const float4 *__restrict__ sorted, int64_t n, int64_t total,
                                                    const KnnGrid *__restrict__ grids, int cell_stride,
                                                    const int32_t *__restrict__ cell_start,
                                                    int32_t *__restrict__ idx, int idx_stride,
                                                    float *__restrict__ d2out, int64_t first) {
  const int64_t t = first + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  knn_ring_search<K>(t, sorted, n, grids, cell_stride, cell_start, idx, idx_stride, d2out);
}

// ---- direct search: one thread per cell-sorted point, survivors in shared-memory columns ---------
// Thread t owns sorted point t (neighbouring threads sit in the same or adjacent cells, so their
// candidate reads hit the same L1 lines).  It scans the nine runs of its 27 cells straight from
// memory and appends every candidate with d2 < (0.999 h)^2 branch-free to its own column of (d2, index)
// keys in shared memory; the top-k insertion network then runs over the ~20 survivors only.  Exact
// whenever at least k survived; sparse points retry once over the 5 x 5 x 5 cells with radius
// 2 * 0.999 h, and what is still short goes to the ring-search list.
#ifndef CTP_KNN_DIRECT_THREADS
#define CTP_KNN_DIRECT_THREADS 256  // measured: 64 -> 2.92, 128 -> 2.80, 256 -> 2.73, 512 -> 2.82 ms per C2 step
#endif
#ifndef CTP_KNN_DIRECT_KEEP
#define CTP_KNN_DIRECT_KEEP 48
#endif
#define CTP_KNN_SMALL_LIST 8192
#define CTP_KNN_DIRECT_SMEM (CTP_KNN_DIRECT_KEEP * CTP_KNN_DIRECT_THREADS * 4)

// Passes: 0 = the 3 x 3 x 3 cells with radius 0.999 h; 1, 2 = wider balls for sparse points, only as
// wide as the density seen so far suggests is needed for ~2k points (kept in a ball of radius lb ->
// 2k in a ball of radius lb * cbrt(2k / kept)), at most (pass + 1) cells.  k_knn_direct runs pass 0
// for every point and lists the sparse ones with their survivor count; k_knn_direct_list runs the
// wider passes for the listed points only, so dense points never wait for sparse lanes.
#ifndef CTP_KNN_PRUNE0
#define CTP_KNN_PRUNE0 0  // pruning in pass 0: 0 none (2.31 ms per C2 step), 1 rows and x-ranges (2.37), 2 corner rows only (2.34)
#endif
template <int K, int PRUNE>
__device__ __forceinline__ void knn_direct_point(int64_t t, int tid, int32_t (*s_keep)[CTP_KNN_DIRECT_THREADS],
                                                 const float4 *__restrict__ sorted, int64_t n,
                                                 const KnnGrid *__restrict__ grids, int cell_stride,
                                                 const int32_t *__restrict__ cell_start, int32_t *__restrict__ idx,
                                                 int idx_stride, float *__restrict__ d2out, int pass_lo, int pass_hi,
                                                 int kept_in, int32_t *__restrict__ fb_list,
                                                 uint8_t *__restrict__ fb_kept, int32_t *__restrict__ fb_count) {
  const int64_t q = t / n;
  const KnnGrid g = grids[q];
  const float4 me = sorted[t];
  const int self = __float_as_int(me.w);
  const float4 *sq = sorted + q * n;
  const int32_t *st = cell_start + q * (cell_stride + 1);
  const int cx = knn_cell_1d(me.x, g.lox, g.inv_h, g.dx);
  const int cy = knn_cell_1d(me.y, g.loy, g.inv_h, g.dy);
  const int cz = knn_cell_1d(me.z, g.loz, g.inv_h, g.dz);
  const float lb = (float)(g.h * 0.999);  // rounding here only shifts which points take the wider search
  const float gx = (me.x - g.lox) * g.inv_h, gy = (me.y - g.loy) * g.inv_h, gz = (me.z - g.loz) * g.inv_h;
  const int selfpos = (int)(t - q * n);
  int kept = 0, kept_prev = kept_in;  // kept counts the point itself as well (d2 = 0 always passes the filter)
  bool done = false;
  float thr2 = 0.f;
  for (int pass = pass_lo; pass <= pass_hi; pass++) {
    const float grow = pass == 0 ? 1.f : fminf((float)(pass + 1), cbrtf((float)(2 * K) / (float)max(kept_prev, 1)));
    const int R = (pass == 0 || grow <= 1.f) ? 1 : (grow > 2.f ? 3 : 2);  // overflowed points shrink: 27 cells suffice
    const float wide = grow * lb;
    thr2 = wide * wide;
    kept = 0;
    const int xa = max(cx - R, 0), xb = min(cx + R, g.dx - 1);
    // Pruning by the ball itself (PRUNE: the wider passes over 125 or 343 cells -- on the 27 cells of pass 0 the extra
    // arithmetic per row and the less uniform runs cost more than the cells saved: 1.72 against 1.67 ms on C2): a cell
    // row (z, y) whose nearest face is farther than `wide` holds no survivor, and in the other rows only the cells the
    // ball's chord reaches.  In cell units, with the point's own float grid coordinate
    // (the expression the binning floors; a point binned into cell c has a coordinate in [c, c + 1), the last cell of an
    // axis also what lies beyond it) and 10^-3 cells of margin for the float roundings of the binning and the distances.
    const float wc = wide * g.inv_h * 1.0001f + 1e-3f, wc2 = wc * wc;
    for (int z = max(cz - R, 0); z <= min(cz + R, g.dz - 1); z++) {
      const float az = fmaxf((z == cz ? 0.f : (z > cz ? (float)z - gz : gz - (float)(z + 1))) - 1e-3f, 0.f);
      for (int y = max(cy - R, 0); y <= min(cy + R, g.dy - 1); y++) {
        int xa2 = xa, xb2 = xb;
        if (PRUNE == 2) {  // pass 0: only the four corner rows are tested, whole rows
          if (z != cz && y != cy) {
            const float ay = fmaxf((y > cy ? (float)y - gy : gy - (float)(y + 1)) - 1e-3f, 0.f);
            if (!(az * az + ay * ay < wc2)) continue;
          }
        }
        if (PRUNE == 1) {
          const float ay = fmaxf((y == cy ? 0.f : (y > cy ? (float)y - gy : gy - (float)(y + 1))) - 1e-3f, 0.f);
          const float rem = wc2 - (az * az + ay * ay);
          if (!(rem > 0.f)) continue;
          const float xr = sqrtf(rem);
          xa2 = max(xa, min((int)floorf(gx - xr), g.dx - 1));
          xb2 = min(xb, max((int)floorf(gx + xr), 0));
          if (xa2 > xb2) continue;
        }
        const int row = (z * g.dy + y) * g.dx;
        const int pb = st[row + xa2], pe = st[row + xb2 + 1];
        for (int c = pb; c < pe; c++) {
          const float4 o = sq[c];
          const float ddx = me.x - o.x, ddy = me.y - o.y, ddz = me.z - o.z;
          // filter only: fused, so a few ulp off the distance the selection uses below; the margin on the
          // k-th survivor after the selection makes up for it
          const float dd = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
          const bool keep = dd < thr2;
          if (keep) s_keep[min(kept, CTP_KNN_DIRECT_KEEP - 1)][tid] = c;  // position in the sorted array
          kept += keep ? 1 : 0;
        }
      }
    }
    if (kept > K) {  // k others besides the point itself
      done = kept <= CTP_KNN_DIRECT_KEEP;
      break;
    }
    kept_prev = kept - 1;
  }
  TopK<K> best;
  if (done) {
    // kept > K survivors (the point itself among them, with the key of an absent entry).  The first K go straight into
    // the K slots and through a sorting network (a third of the instructions of K insertions, and their gathers are
    // all in flight together); the others, four gathers at a time, are inserted only if they beat the current k-th.
    auto key_of = [&](int pos) {
      const float4 o = __ldg(sq + pos);
      const float ddx = me.x - o.x, ddy = me.y - o.y, ddz = me.z - o.z;
      const unsigned long long kk =
          ((unsigned long long)__float_as_uint((ddx * ddx + ddy * ddy) + ddz * ddz) << 32) | (unsigned)__float_as_int(o.w);
      return pos == selfpos ? 0x7f8000007fffffffull : kk;
    };
#pragma unroll
    for (int u = 0; u < K; u++) best.key[u] = key_of(s_keep[u][tid]);
    best.sort();
    for (int u0 = K; u0 < kept; u0 += 4) {
      unsigned long long kq[4];
#pragma unroll
      for (int i = 0; i < 4; i++) kq[i] = u0 + i < kept ? key_of(s_keep[min(u0 + i, CTP_KNN_DIRECT_KEEP - 1)][tid]) : 0x7f8000007fffffffull;
#pragma unroll
      for (int i = 0; i < 4; i++) best.push_key(kq[i]);
    }
    // a candidate the fused filter dropped has d2 >= thr2 * (1 - 2^-22): with the k-th survivor below
    // thr2 * (1 - 1e-5) none of them can belong to the k nearest
    done = best.d(K - 1) < thr2 * 0.99999f;
  }
  if (!done) {
    const int pos = atomicAdd(fb_count, 1);
    fb_list[pos] = (int32_t)t;
    if (fb_kept) fb_kept[pos] = (uint8_t)min(max(kept - 1, 0), 255);
    return;
  }
  int32_t *oi = idx + (q * n + self) * idx_stride;
#pragma unroll
  for (int s2 = 0; s2 < K; s2++) oi[s2] = best.id(s2);
  if (d2out) {
    float *od = d2out + (q * n + self) * K;
#pragma unroll
    for (int s2 = 0; s2 < K; s2++) od[s2] = best.d(s2);
  }
}

template <int K>
__global__ void __launch_bounds__(CTP_KNN_DIRECT_THREADS)
k_knn_direct(const float4 *__restrict__ sorted, int64_t n, int64_t total, const KnnGrid *__restrict__ grids,
             int cell_stride, const int32_t *__restrict__ cell_start, int32_t *__restrict__ idx, int idx_stride,
             float *__restrict__ d2out, int32_t *__restrict__ fb_list, uint8_t *__restrict__ fb_kept,
             int32_t *__restrict__ fb_count, int64_t first) {
  extern __shared__ int32_t s_keep_raw[];  // CTP_KNN_DIRECT_SMEM bytes, passed at launch
  int32_t(*s_keep)[CTP_KNN_DIRECT_THREADS] = reinterpret_cast<int32_t(*)[CTP_KNN_DIRECT_THREADS]>(s_keep_raw);
  // positions [first, total) of the sorted array (first > 0: one GPU's slice of a query split over several)
  const int64_t t = first + blockIdx.x * (int64_t)CTP_KNN_DIRECT_THREADS + threadIdx.x;
  if (t >= total) return;
  knn_direct_point<K, CTP_KNN_PRUNE0>(t, threadIdx.x, s_keep, sorted, n, grids, cell_stride, cell_start, idx, idx_stride, d2out, 0, 0, 0,
                      fb_list, fb_kept, fb_count);
}

// the wider passes over a list of positions in the sorted array (kept = survivors of the pass before)
template <int K>
__global__ void __launch_bounds__(CTP_KNN_DIRECT_THREADS)
k_knn_direct_list(const float4 *__restrict__ sorted, int64_t n, const int32_t *__restrict__ list,
                  const uint8_t *__restrict__ list_kept, const int32_t *__restrict__ count, int pass_lo,
                  const KnnGrid *__restrict__ grids, int cell_stride, const int32_t *__restrict__ cell_start,
                  int32_t *__restrict__ idx, int idx_stride, float *__restrict__ d2out,
                  int32_t *__restrict__ fb_list, int32_t *__restrict__ fb_count) {
  extern __shared__ int32_t s_keep_raw[];  // CTP_KNN_DIRECT_SMEM bytes, passed at launch
  int32_t(*s_keep)[CTP_KNN_DIRECT_THREADS] = reinterpret_cast<int32_t(*)[CTP_KNN_DIRECT_THREADS]>(s_keep_raw);
  const int cnt = *count;
  if (cnt <= CTP_KNN_SMALL_LIST) {
    // latency regime (a single query lists a few hundred points): one thread per point would walk
    // ~10^2 cells serially; hand them to the warp-cooperative ring search instead
    for (int i = blockIdx.x * CTP_KNN_DIRECT_THREADS + threadIdx.x; i < cnt; i += gridDim.x * CTP_KNN_DIRECT_THREADS)
      fb_list[atomicAdd(fb_count, 1)] = list[i];
    return;
  }
  for (int i = blockIdx.x * CTP_KNN_DIRECT_THREADS + threadIdx.x; i < cnt; i += gridDim.x * CTP_KNN_DIRECT_THREADS)
    knn_direct_point<K, 1>(list[i], threadIdx.x, s_keep, sorted, n, grids, cell_stride, cell_start, idx, idx_stride, d2out,
                        pass_lo, 2, list_kept ? list_kept[i] : 0, fb_list, nullptr, fb_count);
}

// Radius-neighbour candidates (north_star's wording; the reference itself takes the k nearest, :329-339): with a radius
// set, a neighbour farther than it is dropped from the table (slot = -1, the "no neighbour" value every later stage
// knows), so a node's candidates are its nearest ones WITHIN the radius, at most k.  The distance is the float32 one
// the search itself ranks by ((dx*dx + dy*dy) + dz*dz on the coordinates narrowed to float).
__global__ void __launch_bounds__(256) k_knn_radius_mask(const double *__restrict__ nodes, int64_t n, int64_t total_rows, int k,
                                                         float r2, int32_t *__restrict__ idx, int idx_stride,
                                                         float *__restrict__ d2out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= total_rows * k) return;
  const int64_t row = e / k;
  const int s = (int)(e - row * k);
  const int j = idx[row * idx_stride + s];
  if (j < 0) return;
  const int64_t qbase = (row / n) * n;
  const double *a = nodes + 3 * row, *b = nodes + 3 * (qbase + j);
  const float dx = (float)a[0] - (float)b[0], dy = (float)a[1] - (float)b[1], dz = (float)a[2] - (float)b[2];
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  if (d2 > r2) {
    idx[row * idx_stride + s] = -1;
    if (d2out) d2out[row * k + s] = __int_as_float(0x7f800000);
  }
}

// ---- warp-level selection helpers: (d2, index) keys, bitonic networks over the lanes ------------
__device__ __forceinline__ unsigned long long knn_min64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long knn_max64(unsigned long long a, unsigned long long b) { return a < b ? b : a; }

// 32 keys, one per lane, ascending over lanes
__device__ __forceinline__ unsigned long long knn_sort32(unsigned long long key, int lane) {
#pragma unroll
  for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, jj);
      const bool take_min = (((lane & kk) == 0) == ((lane & jj) == 0));
      key = take_min ? knn_min64(other, key) : knn_max64(other, key);
    }
  }
  return key;
}

// 64 keys, element e = lane (key) or lane + 32 (key2), ascending; afterwards `key` of lanes 0..31 holds
// the 32 smallest in order
__device__ __noinline__ unsigned long long knn_sort64_low(unsigned long long key, unsigned long long key2, int lane) {
#pragma unroll
  for (int kk = 2; kk <= 64; kk <<= 1) {
#pragma unroll
    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
      if (jj == 32) {  // partner is the other register of the same lane (kk == 64: ascending)
        const unsigned long long mn = knn_min64(key, key2), mx = knn_max64(key, key2);
        key = mn;
        key2 = mx;
      } else {
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, key, jj), o2 = __shfl_xor_sync(0xffffffffu, key2, jj);
        const bool lower = (lane & jj) == 0;
        const bool up1 = (lane & kk) == 0, up2 = ((lane + 32) & kk) == 0;
        key = (up1 == lower) ? knn_min64(o1, key) : knn_max64(o1, key);
        key2 = (up2 == lower) ? knn_min64(o2, key2) : knn_max64(o2, key2);
      }
    }
  }
  return key;
}

// warp-wide: append the lanes' surviving (dd, j) keys; with more than 32 pending keep the K best
template <int K>
__device__ __forceinline__ void knn_offer(bool surv, float dd, int j, int &base, float &thr, unsigned long long *keys,
                                          int lane, unsigned lt_mask) {
  const unsigned b = __ballot_sync(0xffffffffu, surv);
  if (b == 0) return;
  const int pos = base + __popc(b & lt_mask);
  if (surv) keys[pos] = ((unsigned long long)__float_as_uint(dd) << 32) | (unsigned)j;
  base += __popc(b);
  if (base > 32) {
    __syncwarp();
    unsigned long long k1 = lane < base ? keys[lane] : 0xffffffffffffffffull;
    unsigned long long k2 = lane + 32 < base ? keys[lane + 32] : 0xffffffffffffffffull;
    k1 = knn_sort64_low(k1, k2, lane);
    if (lane < K) keys[lane] = k1;
    base = K;
    // later candidates must beat (or tie in distance with) the current k-th
    thr = __uint_as_float((unsigned)(__shfl_sync(0xffffffffu, k1, K - 1) >> 32));
    __syncwarp();
  }
}

// ---- ring search, one warp per listed point ------------------------------------------------------
// What the direct passes could not finish (isolated points, very uneven density) is few but needs
// many cells: a warp walks the Chebyshev rings around the point's cell, lanes over the candidates of
// each row, survivors kept with knn_offer; after each ring the k best so far are sorted and the search
// stops as soon as the k-th is strictly nearer than the ring just completed (0.999 r h).
#define CTP_KNN_RW_WARPS 4

template <int K>
__global__ void __launch_bounds__(32 * CTP_KNN_RW_WARPS)
k_knn_ring_warp(const float4 *__restrict__ sorted, int64_t n, const int32_t *__restrict__ list,
                const int32_t *__restrict__ count, const KnnGrid *__restrict__ grids, int cell_stride,
                const int32_t *__restrict__ cell_start, int32_t *__restrict__ idx, int idx_stride,
                float *__restrict__ d2out) {
  __shared__ unsigned long long s_key[CTP_KNN_RW_WARPS][64];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned long long *keys = s_key[w];
  const int cnt = *count;
  const int n_warps = gridDim.x * CTP_KNN_RW_WARPS;
  for (int li = blockIdx.x * CTP_KNN_RW_WARPS + w; li < cnt; li += n_warps) {
    const int64_t t = list[li];
    const int64_t q = t / n;
    const KnnGrid g = grids[q];
    const float4 me = sorted[t];
    const int self = __float_as_int(me.w);
    const float4 *sq = sorted + q * n;
    const int32_t *st = cell_start + q * (cell_stride + 1);
    const int cx = knn_cell_1d(me.x, g.lox, g.inv_h, g.dx);
    const int cy = knn_cell_1d(me.y, g.loy, g.inv_h, g.dy);
    const int cz = knn_cell_1d(me.z, g.loz, g.inv_h, g.dz);
    int base = 0;
    float thr = __int_as_float(0x7f800000);
    unsigned long long key = 0xffffffffffffffffull;
    const int rmax = max(max(g.dx, g.dy), g.dz);
    __syncwarp();
    for (int r = 0; r <= rmax; r++) {
      const int z0 = cz - r, z1 = cz + r, y0 = cy - r, y1 = cy + r;
      for (int z = max(z0, 0); z <= min(z1, g.dz - 1); z++) {
        const bool zface = (z == z0) | (z == z1);
        for (int y = max(y0, 0); y <= min(y1, g.dy - 1); y++) {
          const bool face = zface | (y == y0) | (y == y1);
          const int nseg = (face || r == 0) ? 1 : 2;  // inside the shell only the two x end cells are new
          for (int sgi = 0; sgi < nseg; sgi++) {
            int xa, xb;
            if (nseg == 1) { xa = cx - r; xb = cx + r; }
            else { xa = xb = (sgi == 0 ? cx - r : cx + r); }
            if (xb < 0 || xa >= g.dx) continue;
            xa = max(xa, 0);
            xb = min(xb, g.dx - 1);
            const int row = (z * g.dy + y) * g.dx;
            const int b0 = st[row + xa], len = st[row + xb + 1] - b0;
            for (int i0 = 0; i0 < len; i0 += 32) {
              const int i = i0 + lane;
              bool surv = false;
              float dd = 0.f;
              int j = -1;
              if (i < len) {
                const float4 o = sq[b0 + i];
                const float ddx = me.x - o.x, ddy = me.y - o.y, ddz = me.z - o.z;
                dd = (ddx * ddx + ddy * ddy) + ddz * ddz;
                j = __float_as_int(o.w);
                surv = (dd <= thr) & (j != self);
              }
              knn_offer<K>(surv, dd, j, base, thr, keys, lane, lt_mask);
            }
          }
        }
      }
      if (base >= K) {  // k best so far, sorted; tighten; stop when the ring just completed bounds the k-th
        __syncwarp();
        key = knn_sort32(lane < base ? keys[lane] : 0xffffffffffffffffull, lane);
        __syncwarp();
        if (lane < K) keys[lane] = key;
        base = K;
        thr = __uint_as_float((unsigned)(__shfl_sync(full, key, K - 1) >> 32));
        __syncwarp();
        const double lbd = (double)r * g.h * 0.999;
        if ((double)thr < lbd * lbd) break;
      }
    }
    if (base < K) {  // fewer than k other points exist: sorted prefix, padded with -1
      __syncwarp();
      key = knn_sort32(lane < base ? keys[lane] : 0xffffffffffffffffull, lane);
    }
    if (lane < K) {
      const bool valid = lane < base;
      idx[(q * n + self) * idx_stride + lane] = valid ? (int)(unsigned)key : -1;
      if (d2out) d2out[(q * n + self) * K + lane] = valid ? __uint_as_float((unsigned)(key >> 32)) : __int_as_float(0x7f800000);
    }
    __syncwarp();
  }
}

// =========================== symmetric CSR emit (:374-385) ================================
// {i,j} is an edge iff free(i->j) with j in kNN(i), or free(j->i) with i in kNN(j); both map
// entries are written on success, so the adjacency is the symmetric union.
//
// Working layout: one 64-byte (NS = 16 ints; 128-byte for k = 16) record per node holding its k
// neighbour ids and, in the last int, the bit mask of its free directed edges.  "Does row j list
// i as a free neighbour" is then NS/4 16-byte loads of one record instead of 2k scalar loads.

__host__ __device__ __forceinline__ int csr_ns(int k) { return k <= 15 ? 16 : 32; }

struct CsrSplit {
  CtpDiv dk, dn;  // by k (edge -> row), by n (row -> query)
  int small;
};
__device__ __forceinline__ void csr_edge_split(int64_t e, int64_t n, int k, const CsrSplit &sp, int64_t &row, int &slot,
                                               int64_t &qbase, int &i) {
  if (sp.small) {
    const uint32_t r = ctp_div_u31((uint32_t)e, sp.dk);
    const uint32_t q = ctp_div_u31(r, sp.dn);
    row = r;
    slot = (int)((uint32_t)e - r * (uint32_t)k);
    qbase = (int64_t)(q * (uint32_t)n);
    i = (int)(r - q * (uint32_t)n);
  } else {
    row = e / k;
    slot = (int)(e - row * k);
    qbase = (row / n) * n;
    i = (int)(row - qbase);
  }
}
__device__ __forceinline__ int64_t csr_row_query(int64_t row, int64_t n, const CsrSplit &sp) {
  return sp.small ? (int64_t)ctp_div_u31((uint32_t)row, sp.dn) : row / n;
}

// free bytes of a row -> mask in the record; own[] = entries the row contributes to itself.  A slot without a
// neighbour (id < 0) never counts as free, whatever the caller's verdict byte says.  upper != 0: only entries with
// col > row are emitted (each undirected edge once), so own[] counts only those.  When every node id fits 16 bits
// (n <= 65535, k <= 14) the row is also written as a 32-byte record of 16-bit ids (slot 15 = mask, absent
// neighbours = 0xffff), which halves the gather of the reverse-edge test.
__global__ void __launch_bounds__(256) k_csr_rowmask(const uint8_t *__restrict__ fr, int64_t n, int k, int ns,
                                                     int64_t total_rows, CsrSplit sp, int upper, int32_t *__restrict__ rec,
                                                     int32_t *__restrict__ own, uint16_t *__restrict__ rec16) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= total_rows) return;
  const int i = (int)(row - csr_row_query(row, n, sp) * n);
  unsigned mask = 0, gt = 0;
  int ids[CTP_KNN_MAXK];
#pragma unroll
  for (int s = 0; s < CTP_KNN_MAXK; s++) {
    ids[s] = s < k ? rec[row * ns + s] : -1;
    if (s < k && fr[row * k + s] && ids[s] >= 0) mask |= 1u << s;
    if (ids[s] > i) gt |= 1u << s;
  }
  rec[row * ns + ns - 1] = (int)mask;
  own[row] = __popc(upper ? (mask & gt) : mask);
  if (rec16) {
    unsigned short v[16];
#pragma unroll
    for (int s = 0; s < 15; s++) v[s] = (unsigned short)ids[s];  // -1 -> 0xffff
    v[15] = (unsigned short)mask;
    uint4 lo, hi;
    lo.x = v[0] | (v[1] << 16); lo.y = v[2] | (v[3] << 16); lo.z = v[4] | (v[5] << 16); lo.w = v[6] | (v[7] << 16);
    hi.x = v[8] | (v[9] << 16); hi.y = v[10] | (v[11] << 16); hi.z = v[12] | (v[13] << 16); hi.w = v[14] | (v[15] << 16);
    uint4 *o = reinterpret_cast<uint4 *>(rec16 + row * 16);
    o[0] = lo;
    o[1] = hi;
  }
}

// the 32-byte record of a row in one 256-bit load (sm_100 LDG.256)
__device__ __forceinline__ void csr_load_rec16(const uint16_t *__restrict__ rec16, int64_t row, unsigned w[8]) {
  asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
      : "l"(rec16 + row * 16));
}
// does a 16-bit record list `target` as a free neighbour?  mask_out = the record's free mask
__device__ __forceinline__ bool csr_lists_free16(const unsigned w[8], int k, int target, unsigned &mask_out) {
  const unsigned mask = (w[7] >> 16) & ((1u << k) - 1u);
  const unsigned t = (unsigned)target & 0xffffu;
  unsigned hit = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    hit |= ((w[i] & 0xffffu) == t ? 1u : 0u) << (2 * i);
    hit |= ((w[i] >> 16) == t ? 1u : 0u) << (2 * i + 1);
  }
  mask_out = mask;
  return (hit & mask) != 0;
}

template <int NS>
__device__ __forceinline__ bool csr_lists_free(const int32_t *__restrict__ rec, int64_t row, int k, int target,
                                               unsigned &mask_out) {
  const int4 *r4 = reinterpret_cast<const int4 *>(rec + row * NS);
  int v[NS];
#pragma unroll
  for (int t = 0; t < NS / 4; t++) {
    const int4 q = __ldg(r4 + t);
    v[4 * t] = q.x; v[4 * t + 1] = q.y; v[4 * t + 2] = q.z; v[4 * t + 3] = q.w;
  }
  const unsigned mask = (unsigned)v[NS - 1] & ((1u << k) - 1u);
  unsigned hit = 0;
#pragma unroll
  for (int s = 0; s < NS - 1; s++) hit |= (v[s] == target ? 1u : 0u) << s;
  mask_out = mask;
  return (hit & mask) != 0;
}

// Extra entries.  A free directed edge i -> j always puts j into row i (an "own" entry, known from row i's mask);
// it also puts i into row j unless row j lists i as a free neighbour itself.  Those extra entries are pushed here:
// one thread per row i walks its free slots, gathers the record of each neighbour j (one 32-byte load) and, where
// j does not list i, bumps xcnt[j] and parks i in row j's staging line (CTP_CSR_XCAP ids per row; the value the
// atomic returns is the slot).  A row whose line overflows is put on the `big` list and finished by k_csr_emit_big
// (its count stays exact).  upper != 0: row j only holds ids above j, so only slots with j < i push anything.
#define CTP_CSR_XCAP 16

template <int NS, typename XT, bool R16>
__global__ void __launch_bounds__(256) k_csr_count(const int32_t *__restrict__ rec, const uint16_t *__restrict__ rec16,
                                                   int64_t n, int k, int64_t total_rows, CsrSplit sp, int upper,
                                                   int32_t *__restrict__ xcnt, XT *__restrict__ xtra,
                                                   int32_t *__restrict__ big_list, int32_t *__restrict__ big_cnt,
                                                   int2 *__restrict__ big_pairs, int pair_cap) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= total_rows) return;
  const int64_t q = csr_row_query(row, n, sp);
  const int64_t qbase = q * n;
  const int i = (int)(row - qbase);
  int ids[CTP_KNN_MAXK];
  unsigned mask;
  if (R16) {
    unsigned w[8];
    csr_load_rec16(rec16, row, w);
#pragma unroll
    for (int s = 0; s < 15; s++) ids[s] = (int)((s & 1) ? (w[s >> 1] >> 16) : (w[s >> 1] & 0xffffu));
    ids[15] = 0;
    mask = (w[7] >> 16) & ((1u << k) - 1u);
  } else {
    const int4 *r4 = reinterpret_cast<const int4 *>(rec + row * NS);
    int v[NS];
#pragma unroll
    for (int t = 0; t < NS / 4; t++) {
      const int4 qq = __ldg(r4 + t);
      v[4 * t] = qq.x; v[4 * t + 1] = qq.y; v[4 * t + 2] = qq.z; v[4 * t + 3] = qq.w;
    }
#pragma unroll
    for (int s = 0; s < CTP_KNN_MAXK; s++) ids[s] = v[s];
    mask = (unsigned)v[NS - 1] & ((1u << k) - 1u);
  }
  if (upper) {
#pragma unroll
    for (int s = 0; s < CTP_KNN_MAXK; s++)
      if (ids[s] >= i) mask &= ~(1u << s);
  }
  // four neighbour records in flight at a time
#pragma unroll
  for (int s0 = 0; s0 < CTP_KNN_MAXK; s0 += 4) {
    if (!((mask >> s0) & 15u)) continue;
    bool rev[4];
    unsigned mj[4];
    (void)mj;
    if (R16) {
      unsigned w[4][8];
#pragma unroll
      for (int u = 0; u < 4; u++)
        if ((mask >> (s0 + u)) & 1u) csr_load_rec16(rec16, qbase + ids[s0 + u], w[u]);
#pragma unroll
      for (int u = 0; u < 4; u++)
        if ((mask >> (s0 + u)) & 1u) rev[u] = csr_lists_free16(w[u], k, i, mj[u]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; u++)
        if ((mask >> (s0 + u)) & 1u) rev[u] = csr_lists_free<NS>(rec, qbase + ids[s0 + u], k, i, mj[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (!((mask >> (s0 + u)) & 1u) || rev[u]) continue;
      const int j = ids[s0 + u];
      const int pos = atomicAdd(xcnt + qbase + j, 1);
      if (pos < CTP_CSR_XCAP) {
        xtra[(qbase + j) * CTP_CSR_XCAP + pos] = (XT)i;
      } else {  // the line is full: row j goes on the `big` list once, the id into the overflow pairs
        if (pos == CTP_CSR_XCAP) big_list[atomicAdd(big_cnt, 1)] = (int32_t)(qbase + j);
        const int p = atomicAdd(big_cnt + 1, 1);
        if (p < pair_cap) big_pairs[p] = make_int2((int)(qbase + j), i);
      }
    }
  }
}

// per query: row_ptr (relative) = exclusive scan of degrees; nnz per query; CL CTAs (a cluster) per query
template <int CL>
__global__ void __launch_bounds__(1024) k_csr_rowptr(const int32_t *__restrict__ own, const int32_t *__restrict__ xcnt,
                                                     int64_t n, int32_t *__restrict__ row_ptr, int64_t *__restrict__ nnz) {
  const int64_t q = blockIdx.x / CL;
  const int rank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
  const int32_t *dq = own + q * n, *xq = xcnt + q * n;
  int32_t *rp = row_ptr + q * (n + 1);
  const long long total = cluster_excl_scan<CL>(
      rank, n, [&](int64_t i) { return dq[i] + xq[i]; }, [&](int64_t i, int pre) { rp[i] = pre; },
      [&](int64_t i, int off) { rp[i] += off; });
  if (threadIdx.x == 0 && rank == 0) {
    rp[n] = (int32_t)total;
    nnz[q] = total;
  }
}

// single CTA: nnz_off = exclusive scan over queries (int64), nnz_off[q] = total
__global__ void __launch_bounds__(1024) k_csr_query_offsets(const int64_t *__restrict__ nnz, int64_t q,
                                                            int64_t *__restrict__ nnz_off) {
  __shared__ long long wsum[32];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < q; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const long long v = i < q ? nnz[i] : 0;
    long long s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) >= o) s += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      long long w = wsum[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += t;
      }
      wsum[threadIdx.x] = w;
    }
    __syncthreads();
    const long long woff = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0;
    const long long c = carry;
    if (i < q) nnz_off[i] = c + woff + s - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + woff + s;
    __syncthreads();
  }
  if (threadIdx.x == 0) nnz_off[q] = carry;
}

// ascending sorting network over N registers (N a power of two), fully unrolled
template <int N> __device__ __forceinline__ void csr_reg_sort(unsigned (&v)[N]) {
#pragma unroll
  for (int kk = 2; kk <= N; kk <<= 1) {
#pragma unroll
    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
#pragma unroll
      for (int i = 0; i < N; i++) {
        const int l = i ^ jj;
        if (l > i) {
          const unsigned a = v[i], b = v[l];
          const unsigned lo = min(a, b), hi = max(a, b);
          const bool up = (i & kk) == 0;
          v[i] = up ? lo : hi;
          v[l] = up ? hi : lo;
        }
      }
    }
  }
}

// Emit: a CTA takes CTP_CSR_EMIT_ROWS consecutive rows.  Phase 1, one thread per row: the row's own free ids and
// its parked extra ids go through a 32-wide sorting network in registers (absent slots = 0xffffffff sort to the
// tail) and the row's ascending ids land in shared memory, back to back in row order.  Phase 2, one thread per
// ENTRY: the FP64 weight ||to - from|| (:369-371) from the staged own coordinates and one gathered neighbour, then
// col / w go out in final order -- every global load and store of the kernel is coalesced except that one gather.
#ifndef CTP_CSR_EMIT_ROWS
#define CTP_CSR_EMIT_ROWS 128
#endif

template <int NS, typename XT>
__global__ void __launch_bounds__(CTP_CSR_EMIT_ROWS)
k_csr_emit(const double *__restrict__ nodes, const int32_t *__restrict__ rec, const uint16_t *__restrict__ rec16,
           const XT *__restrict__ xtra, int64_t n, int k, int64_t total_rows, CsrSplit sp, int upper,
           const int32_t *__restrict__ row_ptr, const int64_t *__restrict__ nnz_off, int64_t cap,
           int32_t *__restrict__ col, double *__restrict__ w) {
  __shared__ unsigned s_col[CTP_CSR_EMIT_ROWS * 32];
  __shared__ uint8_t s_own[CTP_CSR_EMIT_ROWS * 32];
  __shared__ double s_node[CTP_CSR_EMIT_ROWS][3];
  __shared__ long long s_gbeg[CTP_CSR_EMIT_ROWS];  // global position of the row's first entry
  __shared__ int s_loff[CTP_CSR_EMIT_ROWS + 1];    // position of the row's first entry in s_col
  __shared__ long long s_qbase[CTP_CSR_EMIT_ROWS];
  __shared__ int s_wsum[CTP_CSR_EMIT_ROWS / 32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int64_t row = (int64_t)blockIdx.x * CTP_CSR_EMIT_ROWS + t;
  unsigned v[32];
  int cnt = 0;
  if (row < total_rows) {
    const int64_t q = csr_row_query(row, n, sp);
    const int64_t qbase = q * n;
    const int i = (int)(row - qbase);
    const int32_t *rp = row_ptr + q * (n + 1);
    const int rb = rp[i], deg_i = rp[i + 1] - rb;
    s_gbeg[t] = nnz_off[q] + rb;
    s_qbase[t] = qbase;
    s_node[t][0] = nodes[3 * row];
    s_node[t][1] = nodes[3 * row + 1];
    s_node[t][2] = nodes[3 * row + 2];
    unsigned mask;
    if (rec16) {
      unsigned ww[8];
      csr_load_rec16(rec16, row, ww);
      mask = (ww[7] >> 16) & ((1u << k) - 1u);
#pragma unroll
      for (int s = 0; s < 16; s++) {
        const unsigned id = (s & 1) ? (ww[s >> 1] >> 16) : (ww[s >> 1] & 0xffffu);
        if (upper && (int)id <= i) mask &= ~(1u << s);
        v[s] = ((mask >> s) & 1u) ? id : 0xffffffffu;
      }
    } else {
      const int4 *r4 = reinterpret_cast<const int4 *>(rec + row * NS);
      int rr[NS];
#pragma unroll
      for (int u = 0; u < NS / 4; u++) {
        const int4 qq = __ldg(r4 + u);
        rr[4 * u] = qq.x; rr[4 * u + 1] = qq.y; rr[4 * u + 2] = qq.z; rr[4 * u + 3] = qq.w;
      }
      mask = (unsigned)rr[NS - 1] & ((1u << k) - 1u);
#pragma unroll
      for (int s = 0; s < 16; s++) {
        if (upper && rr[s] <= i) mask &= ~(1u << s);
        v[s] = ((mask >> s) & 1u) ? (unsigned)rr[s] : 0xffffffffu;
      }
    }
    const int nx = deg_i - __popc(mask);
    if (nx <= CTP_CSR_XCAP) {  // else: a `big` row, written by k_csr_emit_big
      cnt = deg_i;
      const XT *xr = xtra + row * CTP_CSR_XCAP;
      if (sizeof(XT) == 2) {
        const uint4 *x4 = reinterpret_cast<const uint4 *>(xr);
        const uint4 a = nx > 0 ? x4[0] : make_uint4(0, 0, 0, 0), b = nx > 8 ? x4[1] : make_uint4(0, 0, 0, 0);
        const unsigned ww[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int s = 0; s < 16; s++) {
          const unsigned id = (s & 1) ? (ww[s >> 1] >> 16) : (ww[s >> 1] & 0xffffu);
          v[16 + s] = s < nx ? id : 0xffffffffu;
        }
      } else {
        const uint4 *x4 = reinterpret_cast<const uint4 *>(xr);
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint4 a = nx > 4 * u ? x4[u] : make_uint4(0, 0, 0, 0);
          v[16 + 4 * u] = 4 * u < nx ? a.x : 0xffffffffu;
          v[16 + 4 * u + 1] = 4 * u + 1 < nx ? a.y : 0xffffffffu;
          v[16 + 4 * u + 2] = 4 * u + 2 < nx ? a.z : 0xffffffffu;
          v[16 + 4 * u + 3] = 4 * u + 3 < nx ? a.w : 0xffffffffu;
        }
      }
      csr_reg_sort<32>(v);
    }
  }
  // exclusive scan of cnt over the CTA
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_wsum[wid] = incl;
  __syncthreads();
  int woff = 0;
#pragma unroll
  for (int u = 0; u < CTP_CSR_EMIT_ROWS / 32; u++) woff += u < wid ? s_wsum[u] : 0;
  const int loff = woff + incl - cnt;
  s_loff[t] = loff;
  if (t == CTP_CSR_EMIT_ROWS - 1) s_loff[CTP_CSR_EMIT_ROWS] = loff + cnt;
  // ids below 2^25 leave room for the owning row (7 bits) in the same word: one shared load per entry instead of two
  const bool pack = n < ((int64_t)1 << 25);
#pragma unroll
  for (int s = 0; s < 32; s++)
    if (s < cnt) {
      s_col[loff + s] = pack ? (v[s] | ((unsigned)t << 25)) : v[s];
      if (!pack) s_own[loff + s] = (uint8_t)t;
    }
  // The usual CTA: all rows of one query, none of them a `big` row -- then the CTA's entries are one contiguous run of
  // col / w that starts at the first row's position, and the per-row tables are not needed to place an entry.
  const bool row_ok = row >= total_rows || (s_qbase[t] == s_qbase[0] && s_gbeg[t] == s_gbeg[0] + loff);
  const bool flat = __syncthreads_and(row_ok) != 0;  // (also the barrier between the phases)
  const int total = s_loff[CTP_CSR_EMIT_ROWS];
  const long long g0 = s_gbeg[0], q0 = s_qbase[0];
  // four entries per thread and step: their neighbour gathers (the only scattered loads of the kernel) are issued
  // together, the FP64 norms follow
  for (int x0 = t; x0 < total; x0 += 4 * CTP_CSR_EMIT_ROWS) {
    int ow[4];
    unsigned c[4];
    double ox[4], oy[4], oz[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int x = x0 + u * CTP_CSR_EMIT_ROWS;
      const bool in = x < total;
      const unsigned word = in ? s_col[x] : 0u;
      ow[u] = pack ? (int)(word >> 25) : (in ? s_own[x] : 0);
      c[u] = pack ? (word & 0x1ffffffu) : word;
      const double *o = nodes + 3 * ((flat ? q0 : s_qbase[ow[u]]) + (int64_t)c[u]);
      ox[u] = in ? o[0] : 0.0;
      oy[u] = in ? o[1] : 0.0;
      oz[u] = in ? o[2] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int x = x0 + u * CTP_CSR_EMIT_ROWS;
      if (x >= total) continue;
      const int64_t pos = flat ? g0 + x : s_gbeg[ow[u]] + (x - s_loff[ow[u]]);
      if (pos < cap) {
        col[pos] = (int32_t)c[u];
        if (w) {
          const double dx = ox[u] - s_node[ow[u]][0], dy = oy[u] - s_node[ow[u]][1], dz = oz[u] - s_node[ow[u]][2];
          w[pos] = sqrt((dx * dx + dy * dy) + dz * dz);
        }
      }
    }
  }
}

// Rows with more than CTP_CSR_XCAP extra entries (a node many others list without being listed back: never seen
// on roadmaps of 10^4 nodes, a few per query at 10^6).  One CTA per listed row j: own free ids, the ids parked in
// its staging line, then those that overflowed it (the pair list k_csr_count appended to; should that list itself
// have overflowed: every i of the query whose record lists j free while j's does not list i), collected straight into
// the row's col range; ranked into order through a per-CTA scratch line (tmp: gridDim.x lines of n + 32 ids);
// weights last.
template <int NS, typename XT>
__global__ void __launch_bounds__(256) k_csr_emit_big(const double *__restrict__ nodes, const int32_t *__restrict__ rec,
                                                      int64_t n, int k, CsrSplit sp, int upper, const int32_t *__restrict__ big_list,
                                                      const int32_t *__restrict__ big_cnt, const XT *__restrict__ xtra,
                                                      const int2 *__restrict__ big_pairs, int pair_cap,
                                                      const int32_t *__restrict__ row_ptr,
                                                      const int64_t *__restrict__ nnz_off, int64_t cap,
                                                      int32_t *__restrict__ tmp_all, int32_t *__restrict__ col,
                                                      double *__restrict__ w) {
  __shared__ int s_fill;
  const int nbig = *big_cnt;
  int32_t *tmp = tmp_all + (size_t)blockIdx.x * (size_t)(n + 32);
  for (int b = blockIdx.x; b < nbig; b += gridDim.x) {
    const int64_t row = big_list[b];
    const int64_t q = csr_row_query(row, n, sp), qbase = q * n;
    const int j = (int)(row - qbase);
    const int32_t *rp = row_ptr + q * (n + 1);
    const int64_t gb = nnz_off[q] + rp[j];
    const int deg_j = rp[j + 1] - rp[j];
    if (gb + deg_j > cap) continue;  // uniform over the CTA
    const unsigned kmask = (1u << k) - 1u;
    const unsigned mj = (unsigned)rec[row * NS + NS - 1] & kmask;
    if (threadIdx.x == 0) {
      int c = 0;
      for (int s = 0; s < k; s++)
        if (((mj >> s) & 1u) && (!upper || rec[row * NS + s] > j)) col[gb + c++] = rec[row * NS + s];
      s_fill = c;
    }
    __syncthreads();
    const int npair = big_cnt[1];
    if (npair <= pair_cap) {
      for (int x = threadIdx.x; x < CTP_CSR_XCAP; x += blockDim.x)
        col[gb + atomicAdd(&s_fill, 1)] = (int32_t)xtra[row * CTP_CSR_XCAP + x];
      for (int p = threadIdx.x; p < npair; p += blockDim.x) {
        const int2 pr = big_pairs[p];
        if (pr.x == (int)row) col[gb + atomicAdd(&s_fill, 1)] = pr.y;
      }
    } else
    for (int64_t i = (upper ? j + 1 : 0) + threadIdx.x; i < n; i += blockDim.x) {
      const int32_t *ri = rec + (qbase + i) * NS;
      const unsigned mi = (unsigned)ri[NS - 1] & kmask;
      bool lists = false;
      for (int s = 0; s < k; s++) lists |= ((mi >> s) & 1u) && ri[s] == j;
      if (!lists) continue;
      bool back = false;
      for (int s = 0; s < k; s++) back |= ((mj >> s) & 1u) && rec[row * NS + s] == (int)i;
      if (!back) col[gb + atomicAdd(&s_fill, 1)] = (int32_t)i;
    }
    __syncthreads();
    for (int x = threadIdx.x; x < deg_j; x += blockDim.x) {
      const int32_t c = col[gb + x];
      int rank = 0;
      for (int y = 0; y < deg_j; y++) rank += col[gb + y] < c ? 1 : 0;
      tmp[rank] = c;
    }
    __syncthreads();
    const double *me = nodes + 3 * row;
    for (int x = threadIdx.x; x < deg_j; x += blockDim.x) {
      const int32_t c = tmp[x];
      col[gb + x] = c;
      if (w) {
        const double *o = nodes + 3 * (qbase + c);
        const double dx = o[0] - me[0], dy = o[1] - me[1], dz = o[2] - me[2];
        w[gb + x] = sqrt((dx * dx + dy * dy) + dz * dz);
      }
    }
    __syncthreads();
  }
}

// dense k-NN table (stride k) -> neighbour records (stride NS)
__global__ void __launch_bounds__(256) k_nbr_to_rec(const int32_t *__restrict__ nbr, int ns, int k, int64_t total_edges,
                                                    int32_t *__restrict__ rec) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= total_edges) return;
  const int64_t row = e / k;
  rec[row * ns + (e - row * k)] = nbr[e];
}

// neighbour records (stride NS) -> dense user-visible k-NN table (stride k)
__global__ void __launch_bounds__(256) k_rec_to_nbr(const int32_t *__restrict__ rec, int ns, int k, int64_t total_edges,
                                                    int32_t *__restrict__ nbr) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= total_edges) return;
  const int64_t row = e / k;
  nbr[e] = rec[row * ns + (e - row * k)];
}
