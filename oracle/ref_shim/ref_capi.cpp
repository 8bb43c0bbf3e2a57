// C API over the reference's own ESDFMap / BaseMap code (src/esdf_map.cpp, src/base_map.cpp
// compiled where they lie, against the shim headers in this directory).  Output:
// oracle/_ref/libctopprm_ref.so.  TEST INFRASTRUCTURE: used to pin the restatement in
// oracle/ctp_oracle.c and as the "reference"-kind CPU baseline.  Never linked into the product.
#include <cstring>
#include <memory>
#include "esdf_map.hpp"
#include "../ctp_oracle.h"

namespace {
thread_local double g_u = 0.5;
thread_local double g_g[3] = {1, 0, 0};
typedef HeapNode<Vector<3>> Node;
Vector<3> V(const double *p) { return Vector<3>(p[0], p[1], p[2]); }
void put(const Vector<3> &v, double *o) { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
struct Poly {
  std::vector<Node> store;
  std::vector<Node *> ptrs;
  Poly(const double *pts, int n) : store(n) {
    for (int i = 0; i < n; i++) { store[i].data = V(pts + 3 * i); store[i].id = i; }
    for (int i = 0; i < n; i++) ptrs.push_back(&store[i]);
  }
};
}  // namespace

// src/common.cpp:256 needs Eigen::JacobiSVD; served by the oracle restatement with the random
// draws injected through ref_set_draws() so BaseMap::fillRandomStateInEllipse can still run.
Vector<3> random_ellipsoid_point(Vector<3> A, Vector<3> B, double two_major_length) {
  double o[3];
  orc_ellipsoid_point(A.v, B.v, two_major_length, g_u, g_g, o);
  return V(o);
}

extern "C" {
void *ref_map_load(const char *filename) {
  try {
    ESDFMap *m = new ESDFMap();
    m->load(filename);
    return m;
  } catch (...) { return nullptr; }
}
void ref_map_free(void *h) { delete static_cast<ESDFMap *>(h); }
void ref_min_pos(void *h, double *o) { put(static_cast<ESDFMap *>(h)->getMinPos(), o); }
void ref_max_pos(void *h, double *o) { put(static_cast<ESDFMap *>(h)->getMaxPos(), o); }
double ref_clearance(void *h, const double *p) { return static_cast<ESDFMap *>(h)->getClearence(V(p)); }
double ref_interpolate(void *h, const double *q) { return static_cast<ESDFMap *>(h)->interpolateMapData(V(q)); }
void ref_clearance_batch(void *h, const double *p, int64_t n, double *out) {
  ESDFMap *m = static_cast<ESDFMap *>(h);
  for (int64_t i = 0; i < n; i++) out[i] = m->getClearence(V(p + 3 * i));
}
void ref_gradient_batch(void *h, const double *p, int64_t n, double *g, double *c) {
  ESDFMap *m = static_cast<ESDFMap *>(h);
  for (int64_t i = 0; i < n; i++) {
    auto r = m->gradientInVoxelCenter(V(p + 3 * i));
    put(r.first, g + 3 * i);
    put(r.second, c + 3 * i);
  }
}
void ref_in_collision_batch(void *h, const double *p, int64_t n, double clr, uint8_t *out) {
  ESDFMap *m = static_cast<ESDFMap *>(h);
  for (int64_t i = 0; i < n; i++) out[i] = m->isInCollision(V(p + 3 * i), clr);
}
// returns number of free edges
int64_t ref_edge_nodes_batch(void *h, const double *a, const double *b, int64_t n, double clr, double cdc,
                             uint8_t *free_out, double *hit) {
  ESDFMap *m = static_cast<ESDFMap *>(h);
  int64_t nfree = 0;
  for (int64_t i = 0; i < n; i++) {
    auto r = m->isSimplePathFreeBetweenNodes(V(a + 3 * i), V(b + 3 * i), clr, cdc);
    if (free_out) free_out[i] = r.first;
    if (hit) put(r.second, hit + 3 * i);
    nfree += r.first;
  }
  return nfree;
}
int64_t ref_edge_index_batch(void *h, const double *nodes, const int32_t *from, const int32_t *to, int64_t n,
                             double clr, double cdc, uint8_t *free_out) {
  ESDFMap *m = static_cast<ESDFMap *>(h);
  int64_t nfree = 0;
  for (int64_t i = 0; i < n; i++) {
    bool f = m->isSimplePathFreeBetweenNodes(V(nodes + 3 * (int64_t)from[i]), V(nodes + 3 * (int64_t)to[i]), clr, cdc).first;
    if (free_out) free_out[i] = f;
    nfree += f;
  }
  return nfree;
}
void ref_edge_guards_batch(void *h, const double *a, const double *b, int64_t n, double clr, double cdc,
                           uint8_t *free_out) {
  ESDFMap *m = static_cast<ESDFMap *>(h);
  for (int64_t i = 0; i < n; i++) free_out[i] = m->isSimplePathFreeBetweenGuards(V(a + 3 * i), V(b + 3 * i), clr, cdc);
}
void ref_sample_path(void *h, const double *pts, int npts, double len, double num, double *out) {
  Poly p(pts, npts);
  auto s = static_cast<ESDFMap *>(h)->samplePath(p.ptrs, len, num);
  for (size_t i = 0; i < s.size(); i++) put(s[i], out + 3 * i);
}
int ref_deformable(void *h, const double *p1, int n1, double l1, const double *p2, int n2, double l2, double num,
                   double clr, double cdc) {
  Poly a(p1, n1), b(p2, n2);
  return static_cast<ESDFMap *>(h)->isDeformablePath(a.ptrs, l1, b.ptrs, l2, num, clr, cdc);
}
int ref_deformable_between(void *h, const double *s, const double *e, const double *b1, const double *b2, double clr,
                           double cdc) {
  Node ns(V(s)), ne(V(e)), n1(V(b1)), n2(V(b2));
  return static_cast<ESDFMap *>(h)->isDeformablePathBetween(&ns, &ne, &n1, &n2, clr, cdc);
}
int ref_path_collision_free(void *h, const double *pts, int npts, double clr, double cdc, double *hit) {
  Poly p(pts, npts);
  path_with_length<Vector<3>> pl;
  pl.plan = p.ptrs;
  pl.length = 0;
  auto r = static_cast<ESDFMap *>(h)->isPathCollisionFree(pl, clr, cdc);
  put(r.second, hit);
  return r.first;
}
// Whole TopologicalPRMClustering::sampleMultiple (include/topological_prm_clustering.hpp:288-401)
// for q independent queries with the REFERENCE'S OWN isInCollision / isSimplePathFreeBetweenNodes
// doing every clearance test; the planner header itself cannot be compiled (FLANN, Eigen, yaml-cpp
// absent; common.hpp broken), so the loop structure (:299-327 rejection, :355-397 directed checks and
// symmetric adjacency) is driven from here and the neighbour stage is the oracle's exact k-NN.
// OpenMP over queries: the reference is single-threaded, but its map is read-only, so independent
// queries can use every host core ("reference" CPU baseline of bench.py).
int ref_sample_multiple_batch(void *h, const double *start, const double *goal, const double *cand, int q, int64_t M,
                              int n, int k, double clr, double cdc, int nthreads, int32_t *n_filled, int64_t *n_free,
                              int64_t *nnz, double *nodes_out, int32_t *row_ptr_out, int32_t *col_out, double *w_out) {
  ESDFMap *m = static_cast<ESDFMap *>(h);
  int short_q = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : short_q) num_threads(nthreads > 1 ? nthreads : 1)
  for (int qi = 0; qi < q; qi++) {
    const size_t cap = 2 * (size_t)n * k;
    std::vector<double> nodes_v;
    std::vector<int32_t> rp_v, col_v;
    std::vector<double> w_v;
    double *nodes = nodes_out ? nodes_out + 3 * (size_t)qi * n : (nodes_v.resize(3 * (size_t)n), nodes_v.data());
    int32_t *rp = row_ptr_out ? row_ptr_out + (size_t)qi * (n + 1) : (rp_v.resize((size_t)n + 1), rp_v.data());
    int32_t *col = col_out ? col_out + (size_t)qi * cap : (col_v.resize(cap), col_v.data());
    double *w = w_out ? w_out + (size_t)qi * cap : (w_v.resize(cap), w_v.data());
    std::vector<int32_t> nbr((size_t)n * k);
    std::vector<uint8_t> fr((size_t)n * k);
    memcpy(nodes, start + 3 * (size_t)qi, 24);
    memcpy(nodes + 3, goal + 3 * (size_t)qi, 24);
    int filled = 2;
    const double *cq = cand + 3 * (size_t)qi * M;
    for (int64_t c = 0; c < M && filled < n; c++)
      if (!m->isInCollision(V(cq + 3 * c), clr)) {
        memcpy(nodes + 3 * (size_t)filled, cq + 3 * c, 24);
        filled++;
      }
    n_filled[qi] = filled;
    if (filled < n) {
      short_q++;
      n_free[qi] = nnz[qi] = 0;
      continue;
    }
    orc_knn_grid(nodes, n, k, nbr.data(), nullptr, 1);
    int64_t f = 0;
    for (size_t e = 0; e < (size_t)n * k; e++) {
      const int to = nbr[e];
      bool ok = false;
      if (to >= 0) ok = m->isSimplePathFreeBetweenNodes(V(nodes + 3 * (e / k)), V(nodes + 3 * (size_t)to), clr, cdc).first;
      fr[e] = ok;
      f += ok;
    }
    n_free[qi] = f;
    nnz[qi] = orc_csr_from_directed(nodes, n, k, nbr.data(), fr.data(), rp, col, w);
  }
  return short_q;
}

void ref_set_draws(double u, const double *g) { g_u = u; memcpy(g_g, g, 24); }
void ref_fill_random_state_in_ellipse(void *h, const double *s, const double *g, double ratio, int planar, double *out) {
  Node ns(V(s)), ng(V(g)), nn;
  static_cast<ESDFMap *>(h)->fillRandomStateInEllipse(&nn, &ns, &ng, ratio, planar != 0);
  put(nn.data, out);
}
}
