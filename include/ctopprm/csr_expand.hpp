// csr_expand.hpp -- host-side mirror of the compact adjacency ctp_sample_multiple_ex / ctp_build_prm_ex_dev hand
// back (CTP_CSR_UPPER | CTP_CSR_NO_WEIGHTS: every undirected edge once, no weights) into the full symmetric,
// weighted CSR that sampleMultiple's callers see (include/topological_prm_clustering.hpp:374-385: both
// visibility_node_ids entries, weight ||to - from|| from the double coordinates, :369-371).  Pure bookkeeping: no
// ESDF lookups happen here.  Compile the including unit with -ffp-contract=off so that the weights carry one
// rounding per operation like the device's.
#pragma once
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace ctopprm {

// one roadmap: rp_up (n+1) / col_up = entries with col > row, ascending within a row.
// Outputs: row_ptr (n+1), col / w (2 * nnz_up entries), ascending within a row.
inline void csr_expand_upper_one(const double *nodes, int64_t n, const int32_t *rp_up, const int32_t *col_up,
                                 int32_t *row_ptr, int32_t *col, double *w) {
  std::vector<int32_t> cur((size_t)n + 1, 0);
  for (int64_t i = 0; i < n; i++)
    for (int32_t e = rp_up[i]; e < rp_up[i + 1]; e++) cur[(size_t)col_up[e] + 1]++;  // entries below the diagonal of row col
  row_ptr[0] = 0;
  for (int64_t i = 0; i < n; i++) row_ptr[i + 1] = row_ptr[i] + cur[(size_t)i + 1] + (rp_up[i + 1] - rp_up[i]);
  for (int64_t i = 0; i < n; i++) cur[(size_t)i] = row_ptr[i];
  auto weight = [&](int64_t from, int64_t to) {
    const double dx = nodes[3 * to] - nodes[3 * from], dy = nodes[3 * to + 1] - nodes[3 * from + 1],
                 dz = nodes[3 * to + 2] - nodes[3 * from + 2];
    return std::sqrt((dx * dx + dy * dy) + dz * dz);
  };
  // rows are visited in ascending order, so the mirrored entries of a row arrive in ascending order as well and
  // all of them (ids below the row) precede its own upper entries
  for (int64_t i = 0; i < n; i++)
    for (int32_t e = rp_up[i]; e < rp_up[i + 1]; e++) {
      const int32_t c = col_up[e];
      const int32_t p = cur[(size_t)c]++;
      col[p] = (int32_t)i;
      if (w) w[p] = weight(c, i);
    }
  for (int64_t i = 0; i < n; i++)
    for (int32_t e = rp_up[i]; e < rp_up[i + 1]; e++) {
      const int32_t p = cur[(size_t)i]++;
      col[p] = col_up[e];
      if (w) w[p] = weight(i, col_up[e]);
    }
}

// q roadmaps packed as the C ABI packs them (row_ptr q x (n+1) relative, col packed, nnz_off q+1); `threads`
// host threads share the queries.  nnz_off[i] = 2 * off_up[i].
inline void csr_expand_upper(const double *nodes, int64_t q, int64_t n, const int32_t *rp_up, const int32_t *col_up,
                             const int64_t *off_up, int32_t *row_ptr, int32_t *col, double *w, int64_t *nnz_off,
                             int threads = 1) {
  for (int64_t i = 0; i <= q; i++) nnz_off[i] = 2 * off_up[i];
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; i++)
      csr_expand_upper_one(nodes + 3 * i * n, n, rp_up + i * (n + 1), col_up + off_up[i], row_ptr + i * (n + 1),
                           col + nnz_off[i], w ? w + nnz_off[i] : nullptr);
  };
  if (threads <= 1 || q <= 1) {
    work(0, q);
    return;
  }
  if (threads > q) threads = (int)q;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) pool.emplace_back(work, q * t / threads, q * (t + 1) / threads);
  for (auto &th : pool) th.join();
}

// The wire formats CTP_NODES_AS_INDEX / CTP_CSR_COL16 of ctp_sample_multiple_ex back into what csr_expand_upper takes:
// node coordinates gathered from the candidate streams the host already holds (cand q x m x 3; -2 = start, -3 = goal,
// -1 = row not filled -> zeros, as the coordinate form leaves them), 16-bit column ids widened.
inline void nodes_from_index(const int32_t *node_src, int64_t q, int64_t n, const double *cand, int64_t m,
                             const double *start, const double *goal, double *nodes) {
  for (int64_t i = 0; i < q; i++)
    for (int64_t r = 0; r < n; r++) {
      const int32_t s = node_src[i * n + r];
      const double *p = s >= 0 ? cand + 3 * (i * m + s) : (s == -2 ? start + 3 * i : (s == -3 ? goal + 3 * i : nullptr));
      double *o = nodes + 3 * (i * n + r);
      o[0] = p ? p[0] : 0.0;
      o[1] = p ? p[1] : 0.0;
      o[2] = p ? p[2] : 0.0;
    }
}
inline void widen_col16(const uint16_t *col16, int64_t nnz, int32_t *col) {
  for (int64_t e = 0; e < nnz; e++) col[e] = (int32_t)col16[e];
}

}  // namespace ctopprm
