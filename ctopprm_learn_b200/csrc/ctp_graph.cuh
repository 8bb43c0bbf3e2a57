// ctp_graph.cuh -- shortest paths over the CSR roadmaps, batched over queries: the graph search the
// planner runs right after the PRM build (findShortestPath = dijkstra.findPath(0, {1}, nodes_),
// include/topological_prm_clustering.hpp:2172, include/dijkstra.hpp:83-176) and the multi-source
// flood that seeds the clusters (wavefrontFill, :789-962).  SURVEY.md section 8f rank 1.
//
// One CTA per query.  Distances live in shared memory as the bit patterns of non-negative doubles
// (ordered like the numbers, so a 64-bit atomicMin relaxes them); a node is re-expanded whenever its
// distance dropped since it was last expanded (frontier Bellman-Ford).  The fixed point equals
// Dijkstra's result: every distance is the FP64 sum of the edge weights along a shortest path in path
// order, exactly what the reference's `distance_from_start + edge` accumulates.  Predecessors are
// chosen afterwards as the smallest neighbour u with dist[u] + w(u,v) == dist[v] (bitwise), which
// makes them deterministic; labels (index of the nearest source) follow the predecessor chains.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CTP_SSSP_THREADS 1024
#define CTP_SSSP_INF 0x7ff0000000000000ull

// dist: n x u64 (shared or global), chg: n x u8
__device__ __forceinline__ void sssp_run(int n, const int32_t *__restrict__ rp, const int32_t *__restrict__ col,
                                         const double *__restrict__ w, unsigned long long *dist, uint8_t *chg) {
  for (;;) {
    int any = 0;
    for (int u = threadIdx.x; u < n; u += blockDim.x) {
      if (!chg[u]) continue;
      chg[u] = 0;
      __threadfence_block();
      const double du = __longlong_as_double((long long)dist[u]);
      const int eb = rp[u], ee = rp[u + 1];
      // four edges per step: their (col, w) loads are issued together, the relaxations follow
      for (int e = eb; e < ee; e += 4) {
        int vv[4];
        double ww[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const bool in = e + t < ee;
          vv[t] = in ? __ldg(col + e + t) : -1;
          ww[t] = in ? __ldg(w + e + t) : 0.0;
        }
#pragma unroll
        for (int t = 0; t < 4; t++) {
          if (vv[t] < 0) continue;
          const int v = vv[t];
          const unsigned long long nd = (unsigned long long)__double_as_longlong(du + ww[t]);
          if (nd < dist[v]) {
            const unsigned long long old = atomicMin(&dist[v], nd);
            if (nd < old) {
              chg[v] = 1;
              any = 1;
            }
          }
        }
      }
    }
    if (!__syncthreads_or(any)) break;
  }
}

// ---- shared-memory variant with an explicit frontier: edge-parallel relaxation ---------------------
// The flag scan above leaves most threads idle (a frontier is a few hundred of the 10^4 nodes) and walks
// each node's edges serially at L2 latency.  Here every sweep (1) turns the "distance dropped" bit mask
// into a queue of node ids, (2) takes the queue in tiles of CTP_SSSP_TILE nodes, scans their degrees and
// (3) hands the tile's EDGES out to all threads, so the (col, w) loads of a sweep are issued together.
// Same fixed point, hence the same distances bit for bit.
#define CTP_SSSP_TILE 256

__device__ __forceinline__ int sssp_block_scan_excl(int v, int *s_warp, int &total) {
  // exclusive scan over the CTA's threads (blockDim.x a multiple of 32, at most 1024), result valid on every thread
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int ws = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += t;
    }
    s_warp[lane] = ws;
  }
  __syncthreads();
  total = s_warp[31];
  const int base = wid ? s_warp[wid - 1] : 0;
  __syncthreads();  // s_warp is reused by the next scan
  return base + incl - v;
}

__global__ void __launch_bounds__(CTP_SSSP_THREADS)
k_sssp_frontier(int64_t n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                const double *__restrict__ w, const int64_t *__restrict__ nnz_off, const int32_t *__restrict__ src, int ns,
                double *__restrict__ dist_out, int32_t *__restrict__ pred_out, int32_t *__restrict__ label_out,
                const double *__restrict__ dist_init) {
  extern __shared__ unsigned long long s_mem[];
  __shared__ int s_warp[32];
  __shared__ int s_off[CTP_SSSP_TILE + 1];
  __shared__ int s_eb[CTP_SSSP_TILE];
  __shared__ double s_du[CTP_SSSP_TILE];
  const int64_t q = blockIdx.x;
  const int32_t *rp = row_ptr + q * (n + 1);
  const int32_t *cq = col + nnz_off[q];
  const double *wq = w + nnz_off[q];
  const int nn = (int)n, words = (nn + 31) >> 5;
  unsigned long long *dist = s_mem;
  unsigned *bits = reinterpret_cast<unsigned *>(s_mem + nn);
  uint16_t *queue = reinterpret_cast<uint16_t *>(bits + words);
  // dist_init (optional): upper bounds the search starts from instead of +inf (the re-flood of addCentroid,
  // include/topological_prm_clustering.hpp:1315-1331, only ever lowers the distances of the previous flood)
  for (int u = threadIdx.x; u < nn; u += blockDim.x)
    dist[u] = dist_init ? (unsigned long long)__double_as_longlong(dist_init[q * n + u]) : CTP_SSSP_INF;
  for (int t = threadIdx.x; t < words; t += blockDim.x) bits[t] = 0;
  __syncthreads();
  if (threadIdx.x < ns) {
    const int s = src[q * ns + threadIdx.x];
    if (s >= 0 && s < nn) {
      dist[s] = 0ull;
      atomicOr(&bits[s >> 5], 1u << (s & 31));
    }
  }
  __syncthreads();
  for (;;) {
    // (1) bit mask -> queue (node ids ascending), mask cleared
    int qlen = 0;
    for (int t0 = 0; t0 < words; t0 += blockDim.x) {
      const int t = t0 + threadIdx.x;
      unsigned m = 0;
      if (t < words) {
        m = bits[t];
        bits[t] = 0;
      }
      int tot;
      int pos = qlen + sssp_block_scan_excl(__popc(m), s_warp, tot);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        queue[pos++] = (uint16_t)((t << 5) + b);
      }
      qlen += tot;
    }
    __syncthreads();
    if (qlen == 0) break;
    // (2) tiles of the queue, (3) edge-parallel relaxation
    for (int tile = 0; tile < qlen; tile += CTP_SSSP_TILE) {
      const int cnt = min(CTP_SSSP_TILE, qlen - tile);
      int deg = 0;
      if (threadIdx.x < cnt) {
        const int u = queue[tile + threadIdx.x];
        const int eb = rp[u];
        deg = rp[u + 1] - eb;
        s_eb[threadIdx.x] = eb;
        s_du[threadIdx.x] = __longlong_as_double((long long)dist[u]);
      }
      int total;
      const int off = sssp_block_scan_excl(deg, s_warp, total);
      if (threadIdx.x <= CTP_SSSP_TILE) s_off[min((int)threadIdx.x, CTP_SSSP_TILE)] = threadIdx.x < cnt ? off : total;
      __syncthreads();
      for (int x = threadIdx.x; x < total; x += blockDim.x) {
        int lo = 0;  // owner = last tile entry whose offset is <= x
#pragma unroll
        for (int step = CTP_SSSP_TILE >> 1; step > 0; step >>= 1)
          if (lo + step < cnt && s_off[lo + step] <= x) lo += step;
        const int e = s_eb[lo] + (x - s_off[lo]);
        const int v = __ldg(cq + e);
        const unsigned long long nd = (unsigned long long)__double_as_longlong(s_du[lo] + __ldg(wq + e));
        if (nd < dist[v]) {
          const unsigned long long old = atomicMin(&dist[v], nd);
          if (nd < old) atomicOr(&bits[v >> 5], 1u << (v & 31));
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // predecessors: smallest neighbour that is strictly nearer and realises the distance
  int32_t *pred = pred_out + q * n;
  for (int v = threadIdx.x; v < nn; v += blockDim.x) {
    const unsigned long long dv = dist[v];
    int best = -1;
    if (dv != CTP_SSSP_INF && dv != 0ull) {
      for (int e = rp[v]; e < rp[v + 1]; e++) {
        const int u = cq[e];
        const unsigned long long du = dist[u];
        if (du >= dv) continue;
        const unsigned long long cand = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)du) + wq[e]);
        if (cand == dv && (best < 0 || u < best)) best = u;
      }
    }
    pred[v] = best;
    dist_out[q * n + v] = __longlong_as_double((long long)dv);
  }
  if (label_out) {
    int32_t *label = label_out + q * n;
    for (int v = threadIdx.x; v < nn; v += blockDim.x) label[v] = 0x7fffffff;
    __syncthreads();
    if (threadIdx.x < ns) {
      const int s = src[q * ns + threadIdx.x];
      if (s >= 0 && s < nn) atomicMin(&label[s], (int)threadIdx.x);
    }
    __syncthreads();
    for (int v = threadIdx.x; v < nn; v += blockDim.x) {
      if (dist[v] == CTP_SSSP_INF || dist[v] == 0ull) continue;
      int u = pred[v];
      while (u >= 0 && *(volatile int32_t *)&label[u] == 0x7fffffff) u = pred[u];
      if (u >= 0) label[v] = *(volatile int32_t *)&label[u];
    }
    __syncthreads();
    for (int v = threadIdx.x; v < nn; v += blockDim.x)
      if (label[v] == 0x7fffffff) label[v] = -1;
  }
}

// ---- one query spread over a thread-block cluster (distributed shared memory) ------------------------
// The same frontier algorithm with the distance array of ONE query cut into CTP_SSSP_CL slices, one per CTA of a
// cluster: a CTA owns the nodes [rank * slice, (rank + 1) * slice), keeps their distances, frontier bits and
// queue in its own shared memory and walks the out-edges of its frontier; a relaxation that lands in another
// slice is a 64-bit atomicMin on that CTA's shared memory through the cluster's distributed-shared-memory
// window.  One cluster barrier per sweep (two frontier bit arrays take turns, a rotating flag per CTA says whether
// anything was improved); within a sweep a warp takes a frontier word and its
// lanes the out-edges of one node at a time -- the per-CTA frontier is a few dozen nodes, so this beats the
// scan-and-search distribution of the one-CTA kernel.  Used when few queries are in flight (a single query keeps
// 8 SMs busy instead of one) and when a query's distances do not fit one CTA's shared memory (up to ~170 000
// nodes stay on chip).  Same fixed point, so the distances are Dijkstra's bit for bit.
#define CTP_SSSP_CL 8        // portable cluster size
#define CTP_SSSP_CL_BIG 16   // opt-in cluster size: twice the shared memory, enough to hold a 10^4-node roadmap's rows
#ifndef CTP_SSSP_CL_THREADS
#define CTP_SSSP_CL_THREADS 1024
#endif
// bytes of dynamic shared memory for a slice of `slice` nodes and room for `cap_e` staged CSR entries
__host__ __device__ inline size_t sssp_cluster_smem(int slice, int cap_e) {
  return (size_t)slice * 8 + (size_t)cap_e * 8 + (size_t)(slice / 32) * 8 + (size_t)(slice + 1) * 4 + (size_t)cap_e * 4 + 16;
}
// CL = CTAs per cluster.  cap_e > 0: a CTA whose rows hold at most cap_e entries copies them (row offsets, column
// ids, weights) into its shared memory first, so that a sweep reads the roadmap on chip instead of paying two L2
// round trips per frontier node after every barrier (the barrier's acquire invalidates the L1).
template <int CL>
__global__ void __launch_bounds__(CTP_SSSP_CL_THREADS)
k_sssp_cluster(int64_t n, int slice, int cap_e, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
               const double *__restrict__ w, const int64_t *__restrict__ nnz_off, const int32_t *__restrict__ src, int ns,
               double *__restrict__ dist_out, int32_t *__restrict__ pred_out, const double *__restrict__ dist_init) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ unsigned long long s_mem[];
  __shared__ int s_work[3];
  const int rank = (int)cluster.block_rank();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int64_t q = blockIdx.x / CL;
  const int32_t *rp = row_ptr + q * (n + 1);
  const int32_t *cq = col + nnz_off[q];
  const double *wq = w + nnz_off[q];
  const int nn = (int)n;
  const int lo = rank * slice;
  const int mine = max(0, min(slice, nn - lo));  // nodes this CTA owns
  const int words = slice >> 5;
  unsigned long long *dist = s_mem;
  double *s_w = reinterpret_cast<double *>(s_mem + slice);
  // two frontier bit arrays: sweep k consumes B[k & 1] (its owner only) while relaxations set bits in B[(k + 1) & 1]
  // of whichever CTA owns the target; s_work[k % 3] of the owner is raised along with a bit, so that after the ONE
  // cluster barrier of the sweep everybody learns from the CL flags whether another sweep is needed
  unsigned *bits0 = reinterpret_cast<unsigned *>(s_w + cap_e);
  unsigned *bits1 = bits0 + words;
  int32_t *s_rp = reinterpret_cast<int32_t *>(bits1 + words);
  int32_t *s_col = s_rp + slice + 1;
  for (int u = threadIdx.x; u < slice; u += blockDim.x)
    dist[u] = (dist_init && u < mine) ? (unsigned long long)__double_as_longlong(dist_init[q * n + lo + u]) : CTP_SSSP_INF;
  for (int t = threadIdx.x; t < 2 * words; t += blockDim.x) bits0[t] = 0;
  if (threadIdx.x < 3) s_work[threadIdx.x] = 0;
  // stage the rows this CTA owns
  const int e_base = mine > 0 ? rp[lo] : 0;
  const int e_cnt = mine > 0 ? rp[lo + mine] - e_base : 0;
  const bool staged = cap_e > 0 && e_cnt <= cap_e;
  if (staged) {
    for (int i = threadIdx.x; i <= mine; i += blockDim.x) s_rp[i] = rp[lo + i] - e_base;
    for (int i = threadIdx.x; i < e_cnt; i += blockDim.x) {
      s_col[i] = __ldg(cq + e_base + i);
      s_w[i] = __ldg(wq + e_base + i);
    }
  }
  cluster.sync();
  if (rank == 0)
    for (int i = threadIdx.x; i < ns; i += blockDim.x) {
      const int s = src[q * ns + i];
      if (s >= 0 && s < nn) {
        const int owner = s / slice, ls = s - owner * slice;
        cluster.map_shared_rank(dist, owner)[ls] = 0ull;
        atomicOr(cluster.map_shared_rank(bits0, owner) + (ls >> 5), 1u << (ls & 31));
        cluster.map_shared_rank(s_work, owner)[2] = 1;  // "sweep -1"
      }
    }
  cluster.sync();
  for (int k = 0;; k++) {
    // did the previous sweep (or the sources) leave work anywhere in the cluster?
    int work = 0;
    if (threadIdx.x < CL) work = cluster.map_shared_rank(s_work, threadIdx.x)[(k + 2) % 3];
    if (__syncthreads_or(work) == 0) break;  // the same CL flags on every CTA
    if (threadIdx.x == 0) s_work[(k + 1) % 3] = 0;  // last read before the previous barrier, next raised after this sweep's
    unsigned *cur = (k & 1) ? bits1 : bits0;
    unsigned *nxt = (k & 1) ? bits0 : bits1;
    int *flag = s_work + (k % 3);
    // a warp per frontier word, its lanes over the out-edges of one node at a time
    for (int word = warp; word < words; word += n_warps) {
      unsigned m = cur[word];
      __syncwarp();
      if (lane == 0 && m) cur[word] = 0;
      while (m) {  // a frontier word rarely holds more than one or two nodes: the whole warp takes one node at a time
        const int lu = (word << 5) + __ffs(m) - 1;  // (8 or 16 lanes per node measured slower: 0.42 / 0.34 ms against 0.29 ms)
        m &= m - 1;
        const double du = __longlong_as_double((long long)dist[lu]);
        const int eb = staged ? s_rp[lu] : rp[lo + lu], ee = staged ? s_rp[lu + 1] : rp[lo + lu + 1];
        for (int e = eb + lane; e < ee; e += 32) {
          const int v = staged ? s_col[e] : __ldg(cq + e);
          const double we = staged ? s_w[e] : __ldg(wq + e);
          const unsigned long long nd = (unsigned long long)__double_as_longlong(du + we);
          const int owner = v / slice, lv = v - owner * slice;
          unsigned long long *rd = cluster.map_shared_rank(dist, owner) + lv;
          // 64-bit minimum as a compare-and-swap loop: a 64-bit atomicMin on shared memory is emulated under a
          // CTA-local lock, which a relaxation arriving from another CTA of the cluster would not respect
          unsigned long long cur_d = *reinterpret_cast<volatile unsigned long long *>(rd);
          while (nd < cur_d) {
            const unsigned long long prev = atomicCAS(rd, cur_d, nd);
            if (prev == cur_d) {
              atomicOr(cluster.map_shared_rank(nxt, owner) + (lv >> 5), 1u << (lv & 31));
              *reinterpret_cast<volatile int *>(cluster.map_shared_rank(flag, owner)) = 1;
              break;
            }
            cur_d = prev;
          }
        }
      }
    }
    cluster.sync();
  }
  // predecessors: smallest neighbour that is strictly nearer and realises the distance
  int32_t *pred = pred_out + q * n;
  for (int lv = threadIdx.x; lv < mine; lv += blockDim.x) {
    const int v = lo + lv;
    const unsigned long long dv = dist[lv];
    int best = -1;
    if (dv != CTP_SSSP_INF && dv != 0ull) {
      const int eb = staged ? s_rp[lv] : rp[v], ee = staged ? s_rp[lv + 1] : rp[v + 1];
      for (int e = eb; e < ee; e++) {
        const int u = staged ? s_col[e] : cq[e];
        const int owner = u / slice;
        const unsigned long long du = cluster.map_shared_rank(dist, owner)[u - owner * slice];
        if (du >= dv) continue;
        const double we = staged ? s_w[e] : wq[e];
        const unsigned long long cand = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)du) + we);
        if (cand == dv && (best < 0 || u < best)) best = u;
      }
    }
    pred[v] = best;
    dist_out[q * n + v] = __longlong_as_double((long long)dv);
  }
  cluster.sync();  // nobody leaves while a neighbour may still read its slice
}

template <bool SMEM>
__global__ void __launch_bounds__(CTP_SSSP_THREADS)
k_sssp(int64_t n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, const double *__restrict__ w,
       const int64_t *__restrict__ nnz_off, const int32_t *__restrict__ src, int ns, double *__restrict__ dist_out,
       int32_t *__restrict__ pred_out, int32_t *__restrict__ label_out, unsigned long long *__restrict__ g_dist,
       uint8_t *__restrict__ g_chg, const double *__restrict__ dist_init) {
  extern __shared__ unsigned long long s_mem[];
  const int64_t q = blockIdx.x;
  const int32_t *rp = row_ptr + q * (n + 1);
  const int32_t *cq = col + nnz_off[q];
  const double *wq = w + nnz_off[q];
  unsigned long long *dist = SMEM ? s_mem : g_dist + q * n;
  uint8_t *chg = SMEM ? reinterpret_cast<uint8_t *>(s_mem + n) : g_chg + q * n;
  for (int u = threadIdx.x; u < n; u += blockDim.x) {
    dist[u] = dist_init ? (unsigned long long)__double_as_longlong(dist_init[q * n + u]) : CTP_SSSP_INF;
    chg[u] = 0;
  }
  __syncthreads();
  if (threadIdx.x < ns) {
    const int s = src[q * ns + threadIdx.x];
    if (s >= 0 && s < n) {
      dist[s] = 0ull;
      chg[s] = 1;
    }
  }
  __syncthreads();
  sssp_run((int)n, rp, cq, wq, dist, chg);
  __syncthreads();
  // predecessors: smallest neighbour that is strictly nearer and realises the distance; sources and
  // unreachable nodes: -1
  int32_t *pred = pred_out + q * n;
  for (int v = threadIdx.x; v < n; v += blockDim.x) {
    const unsigned long long dv = dist[v];
    int best = -1;
    if (dv != CTP_SSSP_INF && dv != 0ull) {
      for (int e = rp[v]; e < rp[v + 1]; e++) {
        const int u = cq[e];
        const unsigned long long du = dist[u];
        if (du >= dv) continue;
        const unsigned long long cand = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)du) + wq[e]);
        if (cand == dv && (best < 0 || u < best)) best = u;
      }
    }
    pred[v] = best;
    dist_out[q * n + v] = __longlong_as_double((long long)dv);
  }
  if (label_out) {
    // label = index (in the src list) of the source at the root of the predecessor chain
    int32_t *label = label_out + q * n;
    for (int v = threadIdx.x; v < n; v += blockDim.x) label[v] = 0x7fffffff;
    __syncthreads();
    if (threadIdx.x < ns) {
      const int s = src[q * ns + threadIdx.x];
      if (s >= 0 && s < n) atomicMin(&label[s], (int)threadIdx.x);  // listed twice: the first index wins
    }
    __syncthreads();
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
      if (dist[v] == CTP_SSSP_INF || dist[v] == 0ull) continue;
      // distances strictly decrease along the chain, so it ends at a source; labels that other threads
      // have already filled in on the way are final and equal to what the root would give
      int u = pred[v];
      while (u >= 0 && *(volatile int32_t *)&label[u] == 0x7fffffff) u = pred[u];
      if (u >= 0) label[v] = *(volatile int32_t *)&label[u];
    }
    __syncthreads();
    for (int v = threadIdx.x; v < n; v += blockDim.x)
      if (label[v] == 0x7fffffff) label[v] = -1;
  }
}

// start -> goal path of every query from the predecessor array: node ids (in the query's node
// numbering) and coordinates, written start first.  path_n = 0 when the goal is unreachable or the
// path does not fit path_cap.
__global__ void __launch_bounds__(128) k_extract_paths(int64_t q, int64_t n, const double *__restrict__ nodes,
                                                       const double *__restrict__ dist, const int32_t *__restrict__ pred,
                                                       int src_node, int goal_node, int path_cap,
                                                       int32_t *__restrict__ path_ids, double *__restrict__ path_xyz,
                                                       int32_t *__restrict__ path_n, double *__restrict__ length) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= q) return;
  const double d = dist[i * n + goal_node];
  const int32_t *pr = pred + i * n;
  length[i] = d;
  int cnt = 0;
  if (d < __longlong_as_double(0x7ff0000000000000LL)) {
    cnt = 1;
    for (int u = goal_node; u != src_node && cnt <= n; u = pr[u]) {
      if (pr[u] < 0) { cnt = 0; break; }
      cnt++;
    }
    if (cnt > path_cap) cnt = 0;
  }
  path_n[i] = cnt;
  int32_t *ids = path_ids + i * path_cap;
  double *xyz = path_xyz + 3 * i * path_cap;
  int u = goal_node;
  for (int p = cnt - 1; p >= 0; p--) {
    ids[p] = u;
    const double *c = nodes + 3 * (i * n + u);
    xyz[3 * p] = c[0];
    xyz[3 * p + 1] = c[1];
    xyz[3 * p + 2] = c[2];
    u = p > 0 ? pr[u] : u;
  }
  for (int p = cnt; p < path_cap; p++) ids[p] = -1;
}

