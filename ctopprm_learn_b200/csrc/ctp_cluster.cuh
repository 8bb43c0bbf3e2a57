// ctp_cluster.cuh -- the floods of CTopPRM's cluster stage on the device CSR, batched over queries:
// wavefrontFill's multi-source Dijkstra with its inter-cluster connection records
// (include/topological_prm_clustering.hpp:826-962) and addCentroid's re-flood from a new seed (:1276-1374).
//
// The reference interleaves three things in one sequential loop: relaxing distances, assigning clusters, and
// recording a "connection" whenever the node being expanded meets a neighbour of another cluster -- with the
// neighbour's distance AS IT STANDS AT THAT MOMENT, which may still be tentative.  Here the distances come from the
// shortest-path kernels of ctp_graph.cuh (the fixed point is Dijkstra's, bit for bit), and everything that depends on
// the pop order is reconstructed per node / per edge from the total order the sequential loop follows,
//     order(a) < order(b)  <=>  (dist[a], a) < (dist[b], b),
// (ties by node id, neighbours in ascending id: the deterministic rule the oracle restatement uses as well):
//   * the predecessor of v is the first node in that order whose relaxation produced dist[v];
//   * the state of a neighbour nb when ex is expanded is its final one if nb was expanded before ex, else the best
//     relaxation it has received from nodes expanded up to and including ex (first one wins a tie);
//   * the number of successful relaxations (the reference's `cnt`, which caps addCentroid at :1319) is the number of
//     strict running minima among a node's incoming relaxations in that order.
// Records come out unordered (one atomic append each); the host layer sorts them by (dist[ex], ex, nb), the order
// in which the sequential loop would have produced them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CTP_CL_INF 0x7ff0000000000000ull

__device__ __forceinline__ bool cl_before(double da, int a, double db, int b) { return da < db || (da == db && a < b); }

// ---- predecessors, labels of the improved nodes, update counts -----------------------------------------------
// mode 0 (full flood): dist_old is ignored (every reached node counts as improved), seeds get label = their index.
// mode 1 (one more seed, cluster id c = n_seeds - 1): only nodes with dist_new < dist_old are touched.
// LPN lanes per node (1 or 32): the floods of ONE query leave a thread-per-node launch on 40 CTAs, each thread walking
// its ~15 x 15 dependent loads alone; with a warp per node the lanes share the node's edges and meet in a shuffle
// reduction (one flood of a 10 000-node roadmap: see DESIGN section 5).
template <int LPN>
__global__ void __launch_bounds__(256)
k_cluster_prev(int64_t q, int64_t n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
               const double *__restrict__ w, const int64_t *__restrict__ nnz_off, const int32_t *__restrict__ seeds,
               int max_seeds, const int32_t *__restrict__ n_seeds, int mode, const double *__restrict__ dist_old,
               const double *__restrict__ dist_new, int32_t *__restrict__ prev, int32_t *__restrict__ label,
               unsigned long long *__restrict__ n_updates) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t t = tid / LPN;  // the node; all LPN lanes of it take the same branches below
  const int sub = (int)(tid - t * LPN);
  if (t >= q * n) return;
  const int64_t qi = t / n;
  const int v = (int)(t - qi * n);
  const int ns = n_seeds[qi];
  if (ns <= 0) return;  // inactive query
  const double inf = __longlong_as_double(CTP_CL_INF);
  const double *dn = dist_new + qi * n;
  const double dv = dn[v];
  const double d0 = mode == 0 ? inf : dist_old[qi * n + v];
  const int32_t *sd = seeds + qi * max_seeds;
  // seeds of this flood
  if (mode == 0) {
    for (int s = 0; s < ns; s++)
      if (sd[s] == v) {  // listed twice: the first index wins, as the reference's loop :836-842 would leave the last
        prev[t] = -1;    // one -- callers never list a seed twice
        label[t] = s;
        return;
      }
    if (!(dv < inf)) {
      prev[t] = -1;
      label[t] = -1;
      return;
    }
  } else {
    if (sd[ns - 1] == v) {
      prev[t] = -1;
      label[t] = ns - 1;
      return;
    }
    if (!(dv < d0)) return;  // not claimed by the new cluster: distance, predecessor and label stay
  }
  const int32_t *rp = row_ptr + qi * (n + 1);
  const int32_t *cq = col + nnz_off[qi];
  const double *wq = w + nnz_off[qi];
  const int eb = rp[v], ee = rp[v + 1];
  // the predecessor: first node in pop order whose relaxation equals the final distance
  int best = -1;
  double bd = 0;
  int succ = 0;
  for (int e = eb + sub; e < ee; e += LPN) {
    const int u = cq[e];
    const double du = dn[u];
    if (!(du < inf)) continue;
    if (mode == 1 && !(du < dist_old[qi * n + u]) && u != sd[ns - 1]) continue;  // u is not expanded by this flood
    const double cand = du + wq[e];
    if (cand == dv && cl_before(du, u, dv, v) && (best < 0 || cl_before(du, u, bd, best))) {
      best = u;
      bd = du;
    }
    // was this relaxation a strict running minimum when it happened?  (the reference's cnt++)
    if (cand < d0) {
      bool low = true;
      for (int e2 = eb; e2 < ee && low; e2++) {
        if (e2 == e) continue;
        const int u2 = cq[e2];
        const double du2 = dn[u2];
        if (!(du2 < inf)) continue;
        if (mode == 1 && !(du2 < dist_old[qi * n + u2]) && u2 != sd[ns - 1]) continue;
        if (cl_before(du2, u2, du, u) && du2 + wq[e2] <= cand) low = false;
      }
      succ += low ? 1 : 0;
    }
  }
  if (LPN > 1) {  // (a node's lanes are one whole warp: 256 threads per CTA, LPN = 32)
#pragma unroll
    for (int o = LPN / 2; o > 0; o >>= 1) {
      const int ob = __shfl_xor_sync(0xffffffffu, best, o);
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      succ += __shfl_xor_sync(0xffffffffu, succ, o);
      if (ob >= 0 && (best < 0 || cl_before(od, ob, bd, best))) {
        best = ob;
        bd = od;
      }
    }
    if (sub) return;
  }
  prev[t] = best;
  if (mode == 1) label[t] = ns - 1;
  if (succ) atomicAdd(n_updates + qi, (unsigned long long)succ);
}

// labels of a full flood: walk the predecessor chain to its seed
__global__ void __launch_bounds__(256)
k_cluster_label(int64_t q, int64_t n, const int32_t *__restrict__ n_seeds, const int32_t *__restrict__ prev,
                const double *__restrict__ dist, int32_t *__restrict__ label) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= q * n) return;
  const int64_t qi = t / n;
  if (n_seeds[qi] <= 0) return;
  if (!(dist[t] < __longlong_as_double(CTP_CL_INF))) return;  // unreachable: -1, set by k_cluster_prev
  const int32_t *pq = prev + qi * n;
  int u = (int)(t - qi * n);
  int guard = 0;
  while (pq[u] >= 0 && guard++ < n) u = pq[u];
  // the root of a reached node is a seed; k_cluster_prev wrote its label (roots are never rewritten here)
  if (pq[(int)(t - qi * n)] >= 0) label[t] = label[qi * n + u];
}

// ---- connection records -------------------------------------------------------------------------------------------
// One thread per expanded node ex; for every neighbour nb the state nb had when ex was expanded (after ex's own
// relaxations, :859-872 / :1315-1331) and, if that state belongs to another cluster, a record
// (ex, nb, cluster_state(nb), dist_state(nb) + dist[ex] + w) -- :930-938 / :1352-1360.
// mode 0: label = labels of this flood.  mode 1: dist_old / label_old = the arrays before this flood, c = new cluster.
// LPN lanes per expanded node, as k_cluster_prev.
template <int LPN>
__global__ void __launch_bounds__(256)
k_cluster_records(int64_t q, int64_t n, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                  const double *__restrict__ w, const int64_t *__restrict__ nnz_off, const int32_t *__restrict__ seeds,
                  int max_seeds, const int32_t *__restrict__ n_seeds, int mode, const double *__restrict__ dist_old,
                  const int32_t *__restrict__ label_old, const double *__restrict__ dist_new,
                  const int32_t *__restrict__ label_new, const int32_t *__restrict__ prev_new, int rec_cap,
                  int32_t *__restrict__ rec_nodes, int32_t *__restrict__ rec_label, double *__restrict__ rec_dist,
                  int32_t *__restrict__ rec_count) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t t = tid / LPN;
  const int sub = (int)(tid - t * LPN);
  if (t >= q * n) return;
  const int64_t qi = t / n;
  const int ex = (int)(t - qi * n);
  const int ns = n_seeds[qi];
  if (ns <= 0) return;
  const double inf = __longlong_as_double(CTP_CL_INF);
  const double *dn = dist_new + qi * n;
  const double dex = dn[ex];
  if (!(dex < inf)) return;  // never expanded
  const int new_seed = seeds[qi * max_seeds + ns - 1];
  if (mode == 1 && !(dex < dist_old[t]) && ex != new_seed) return;  // only the new cluster's nodes are expanded
  const int32_t *rp = row_ptr + qi * (n + 1);
  const int32_t *cq = col + nnz_off[qi];
  const double *wq = w + nnz_off[qi];
  const int32_t *ln = label_new + qi * n;
  const int lex = ln[ex];
  for (int e = rp[ex] + sub; e < rp[ex + 1]; e += LPN) {
    const int nb = cq[e];
    const double dnb = dn[nb];
    double sd_ = 0;
    int sl = -1;
    if (mode == 0) {
      const bool nb_seed = dnb == 0.0 && prev_new[qi * n + nb] < 0;
      if (nb_seed || cl_before(dnb, nb, dex, ex)) {
        sd_ = dnb;
        sl = ln[nb];
      } else {
        // best relaxation received from nodes expanded up to and including ex; the first one wins a tie
        int bu = -1;
        double bc = 0, bdu = 0;
        for (int e2 = rp[nb]; e2 < rp[nb + 1]; e2++) {
          const int u = cq[e2];
          const double du = dn[u];
          if (!(du < inf)) continue;
          if (!(u == ex || cl_before(du, u, dex, ex))) continue;
          const double cand = du + wq[e2];
          if (bu < 0 || cand < bc || (cand == bc && cl_before(du, u, bdu, bu))) {
            bu = u;
            bc = cand;
            bdu = du;
          }
        }
        sd_ = bc;
        sl = ln[bu];  // bu >= 0: ex itself qualifies
      }
      if (sl == lex) continue;
    } else {
      const double d0 = dist_old[qi * n + nb];
      const bool nb_claimed = nb == new_seed || dnb < d0;
      if (nb_claimed && cl_before(dnb, nb, dex, ex)) continue;  // already in the new cluster
      bool taken = false;  // has a node expanded so far (ex included) pulled nb into the new cluster?
      for (int e2 = rp[nb]; e2 < rp[nb + 1] && !taken; e2++) {
        const int u = cq[e2];
        const double du = dn[u];
        if (!(du < inf)) continue;
        if (!(u == new_seed || du < dist_old[qi * n + u])) continue;  // u is not expanded by this flood
        if (!(u == ex || cl_before(du, u, dex, ex))) continue;
        if (du + wq[e2] < d0) taken = true;
      }
      sl = label_old[qi * n + nb];
      if (taken || sl < 0) continue;  // unassigned nodes join the expanding cluster at once (:1319)
      sd_ = d0;
    }
    const int pos = atomicAdd(rec_count + qi, 1);
    if (pos < rec_cap) {
      rec_nodes[2 * (qi * rec_cap + pos)] = ex;
      rec_nodes[2 * (qi * rec_cap + pos) + 1] = nb;
      rec_label[qi * rec_cap + pos] = sl;  // the cluster nb belonged to at that moment
      rec_dist[qi * rec_cap + pos] = (sd_ + dex) + wq[e];
    }
  }
}

// sources of the shortest-path launch of a flood: mode 0 all seeds, mode 1 the newest one (-1 = inactive query)
__global__ void k_cluster_sources(int64_t q, const int32_t *__restrict__ seeds, int max_seeds,
                                  const int32_t *__restrict__ n_seeds, int mode, int32_t *__restrict__ src) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= q * max_seeds) return;
  const int64_t qi = t / max_seeds;
  const int s = (int)(t - qi * max_seeds);
  const int ns = n_seeds[qi];
  int v = -1;
  if (mode == 0) v = s < ns ? seeds[t] : -1;
  else if (s == 0 && ns > 0) v = seeds[qi * max_seeds + ns - 1];
  src[t] = v;
}

// after the records of a flood have been taken: the new distances (and, for a re-flood, labels) become the current ones
__global__ void k_cluster_commit(int64_t total, int64_t n, const int32_t *__restrict__ n_seeds,
                                 const double *__restrict__ dist_new, const int32_t *__restrict__ label_new,
                                 double *__restrict__ dist, int32_t *__restrict__ label) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total || n_seeds[t / n] <= 0) return;
  dist[t] = dist_new[t];
  if (label_new) label[t] = label_new[t];
}
