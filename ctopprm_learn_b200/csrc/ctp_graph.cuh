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
  // exclusive scan over the CTA's threads (blockDim.x = 1024), result valid on every thread
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
    int ws = s_warp[lane];
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
                double *__restrict__ dist_out, int32_t *__restrict__ pred_out, int32_t *__restrict__ label_out) {
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
  for (int u = threadIdx.x; u < nn; u += blockDim.x) dist[u] = CTP_SSSP_INF;
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

template <bool SMEM>
__global__ void __launch_bounds__(CTP_SSSP_THREADS)
k_sssp(int64_t n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, const double *__restrict__ w,
       const int64_t *__restrict__ nnz_off, const int32_t *__restrict__ src, int ns, double *__restrict__ dist_out,
       int32_t *__restrict__ pred_out, int32_t *__restrict__ label_out, unsigned long long *__restrict__ g_dist,
       uint8_t *__restrict__ g_chg) {
  extern __shared__ unsigned long long s_mem[];
  const int64_t q = blockIdx.x;
  const int32_t *rp = row_ptr + q * (n + 1);
  const int32_t *cq = col + nnz_off[q];
  const double *wq = w + nnz_off[q];
  unsigned long long *dist = SMEM ? s_mem : g_dist + q * n;
  uint8_t *chg = SMEM ? reinterpret_cast<uint8_t *>(s_mem + n) : g_chg + q * n;
  for (int u = threadIdx.x; u < n; u += blockDim.x) {
    dist[u] = CTP_SSSP_INF;
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

