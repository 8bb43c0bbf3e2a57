/*
 * ctp_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see ctp_oracle.h for the rules and the parity-pinning statement).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 */
#include "ctp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PRECISION_BASE_MAP (10e-6) /* include/base_map.hpp:18 */
#define PRECISION_SHORTEN (1E-4)   /* include/topological_prm_clustering.hpp:27 */

/* ---- tiny 3-vector helpers, Eigen semantics restated (see header: parity unpinned) ---- */
static inline double v3_norm(const double v[3]) {
  return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
}
static inline void v3_sub(const double a[3], const double b[3], double o[3]) {
  o[0] = a[0] - b[0];
  o[1] = a[1] - b[1];
  o[2] = a[2] - b[2];
}
static inline void v3_cross(const double a[3], const double b[3], double o[3]) {
  double x = a[1] * b[2] - a[2] * b[1];
  double y = a[2] * b[0] - a[0] * b[2];
  double z = a[0] * b[1] - a[1] * b[0];
  o[0] = x;
  o[1] = y;
  o[2] = z;
}
/* Eigen >= 3.3 normalize(): divide by sqrt(squaredNorm) only when squaredNorm > 0 */
static inline void v3_normalize(double v[3]) {
  double z = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
  if (z > 0.0) {
    double s = sqrt(z);
    v[0] /= s;
    v[1] /= s;
    v[2] /= s;
  }
}
static inline double v3_dist(const double a[3], const double b[3]) {
  double d[3];
  v3_sub(a, b, d);
  return v3_norm(d);
}

/* ---------------------------------- ESDF map ------------------------------------------ */

void orc_map_init(orc_map *m, const float *vox, int n, const double c[3], const double e[3]) {
  m->vox = vox;
  m->n = n;
  m->N = (double)n; /* src/esdf_map.cpp:24 */
  for (int i = 0; i < 3; i++) {
    m->c[i] = c[i];
    m->e[i] = e[i];
  }
  m->max_e = fmax(fmax(e[0], e[1]), e[2]); /* src/esdf_map.cpp:27 */
}

void orc_min_pos(const orc_map *m, double out[3]) { /* src/esdf_map.cpp:75 */
  for (int i = 0; i < 3; i++) out[i] = m->c[i] - m->e[i] / 2;
}
void orc_max_pos(const orc_map *m, double out[3]) { /* src/esdf_map.cpp:76 */
  for (int i = 0; i < 3; i++) out[i] = m->c[i] + m->e[i] / 2;
}

static inline double vox_at(const orc_map *m, int x, int y, int z) {
  /* map_data[x][y][z] filled from data[x*s0*s1 + y*s1 + z] (include/esdf_map.hpp:55-66) */
  return (double)m->vox[((size_t)x * m->n + y) * m->n + z];
}

double orc_interpolate(const orc_map *m, const double q[3]) { /* src/esdf_map.cpp:78-121 */
  double xi = q[0], yi = q[1], zi = q[2];
  int x0 = (int)floor(xi), x1 = (int)ceil(xi);
  int y0 = (int)floor(yi), y1 = (int)ceil(yi);
  int z0 = (int)floor(zi), z1 = (int)ceil(zi);
  if (x0 < 0) return NAN;
  if (y0 < 0) return NAN;
  if (z0 < 0) return NAN;
  if (x1 >= m->N) return NAN;
  if (y1 >= m->N) return NAN;
  if (z1 >= m->N) return NAN;
  /* integer coordinate => x1 == x0 => 0/0 = NaN, kept on purpose (:108-110) */
  double xd = (xi - x0) / (x1 - x0);
  double yd = (yi - y0) / (y1 - y0);
  double zd = (zi - z0) / (z1 - z0);
  double c00 = (1 - xd) * vox_at(m, x0, y0, z0) + xd * vox_at(m, x1, y0, z0);
  double c01 = (1 - xd) * vox_at(m, x0, y0, z1) + xd * vox_at(m, x1, y0, z1);
  double c10 = (1 - xd) * vox_at(m, x0, y1, z0) + xd * vox_at(m, x1, y1, z0);
  double c11 = (1 - xd) * vox_at(m, x0, y1, z1) + xd * vox_at(m, x1, y1, z1);
  double c0 = (1 - yd) * c00 + yd * c10;
  double c1 = (1 - yd) * c01 + yd * c11;
  return (1 - zd) * c0 + zd * c1;
}

/* world -> index coordinates, src/esdf_map.cpp:136-139 (and :169-172) */
static inline void world_to_index(const orc_map *m, const double p[3], double q[3]) {
  for (int i = 0; i < 3; i++) {
    double t = p[i] - m->c[i];
    t *= 2.0 / m->max_e;
    t += 1.0;
    t *= m->N / 2.0;
    q[i] = t;
  }
}

double orc_clearance(const orc_map *m, const double p[3]) { /* src/esdf_map.cpp:130-165 */
  for (int i = 0; i < 3; i++) {
    double lo = m->c[i] - m->e[i] / 2, hi = m->c[i] + m->e[i] / 2;
    if (p[i] < lo || p[i] > hi) return NAN;
  }
  double q[3];
  world_to_index(m, p, q);
  /* int xi = round(q): NaN converts to INT_MIN on x86-64, i.e. "< 0" => NaN result */
  for (int i = 0; i < 3; i++) {
    if (isnan(q[i])) return NAN;
    double r = round(q[i]);
    if (r < 0 || r >= m->N) return NAN;
  }
  return orc_interpolate(m, q);
}

void orc_real_pos_from_index(const orc_map *m, const double q[3], double out[3]) {
  /* src/esdf_map.cpp:123-128 */
  double dp = m->max_e / (m->N - 1.0);
  for (int i = 0; i < 3; i++) {
    double min_ext = m->c[i] - 1.0 * m->max_e / 2.0;
    out[i] = min_ext + q[i] * dp;
  }
}

void orc_gradient(const orc_map *m, const double p[3], double grad[3], double centre[3]) {
  /* src/esdf_map.cpp:167-227 */
  double q[3];
  world_to_index(m, p, q);
  int r[3];
  for (int i = 0; i < 3; i++) {
    double rr = round(q[i]);
    /* (int)NaN / out-of-range is INT_MIN on x86-64 */
    r[i] = (isnan(rr) || rr < -2147483648.0 || rr > 2147483647.0) ? (int)0x80000000 : (int)rr;
  }
  orc_real_pos_from_index(m, q, centre); /* :181 "voxel centre" is the rescaled pos itself */
  double d[3];
  for (int ax = 0; ax < 3; ax++) {
    double dval = 0, dd = 0;
    long long rp = (long long)r[ax] + 1, rm = (long long)r[ax] - 1;
    /* r+1 / r-1 wrap in 32-bit int in the reference; INT_MIN-1 is UB there, treat as
     * out of range (the only way to get here is a NaN/absurd position). */
    if (rp >= 0 && rp < m->N) {
      double qq[3] = {q[0], q[1], q[2]};
      qq[ax] = q[ax] + 1;
      dval += orc_interpolate(m, qq);
      dd += m->N;
    }
    if (rm >= 0 && rm < m->N) {
      double qq[3] = {q[0], q[1], q[2]};
      qq[ax] = q[ax] + -1;
      dval -= orc_interpolate(m, qq);
      dd += m->N;
    }
    d[ax] = dval / dd;
  }
  v3_normalize(d); /* .normalized() (:226) */
  grad[0] = d[0];
  grad[1] = d[1];
  grad[2] = d[2];
}

int orc_in_collision(const orc_map *m, const double p[3], double clr) {
  /* src/base_map.cpp:7-12 */
  double c = orc_clearance(m, p);
  return !isfinite(c) || c < clr;
}

/* ------------------------------ segment collision checks ------------------------------ */

int orc_edge_nodes(const orc_map *m, const double a[3], const double b[3], double clr,
                   double cdc, double hit[3], int32_t *hit_idx, int32_t *lookups) {
  /* src/base_map.cpp:49-74 */
  double v[3];
  v3_sub(b, a, v);
  double d = v3_norm(v);
  int32_t nl = 0;
  if (hit) hit[0] = hit[1] = hit[2] = NAN;
  if (hit_idx) *hit_idx = -1;
  if (d < cdc) {
    if (lookups) *lookups = 0;
    return 1;
  }
  double npts = ceil((d / cdc) + 1.0);
  for (int index = 1; index < npts; index += 1) {
    double t = ((double)index / (double)npts);
    double p[3] = {a[0] + v[0] * t, a[1] + v[1] * t, a[2] + v[2] * t};
    nl++;
    if (orc_in_collision(m, p, clr)) {
      if (hit) {
        hit[0] = p[0];
        hit[1] = p[1];
        hit[2] = p[2];
      }
      if (hit_idx) *hit_idx = index;
      if (lookups) *lookups = nl;
      return 0;
    }
  }
  if (lookups) *lookups = nl;
  return 1;
}

int orc_edge_guards(const orc_map *m, const double a[3], const double b[3], double clr,
                    double cdc, int32_t *lookups) {
  /* src/base_map.cpp:77-97: marches from `neigbour` (b) toward `actual` (a), no short-edge
   * shortcut */
  double v[3];
  v3_sub(a, b, v);
  double d = v3_norm(v);
  double npts = ceil((d / cdc) + 1.0);
  int32_t nl = 0;
  for (int index = 1; index < npts; index += 1) {
    double t = ((double)index / (double)npts);
    double p[3] = {b[0] + v[0] * t, b[1] + v[1] * t, b[2] + v[2] * t};
    nl++;
    if (orc_in_collision(m, p, clr)) {
      if (lookups) *lookups = nl;
      return 0;
    }
  }
  if (lookups) *lookups = nl;
  return 1;
}

void orc_clearance_batch(const orc_map *m, const double *xyz, int64_t cnt, double *out,
                         int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 1 ? nthreads : 1)
  for (int64_t i = 0; i < cnt; i++) out[i] = orc_clearance(m, xyz + 3 * i);
}

void orc_gradient_batch(const orc_map *m, const double *xyz, int64_t cnt, double *grad,
                        double *centre, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 1 ? nthreads : 1)
  for (int64_t i = 0; i < cnt; i++) orc_gradient(m, xyz + 3 * i, grad + 3 * i, centre + 3 * i);
}

int64_t orc_edge_nodes_batch(const orc_map *m, const double *a, const double *b, int64_t cnt,
                             double clr, double cdc, uint8_t *free_out, double *hit,
                             int32_t *hit_idx, int32_t *lookups, int nthreads) {
  int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : total) \
    num_threads(nthreads > 1 ? nthreads : 1)
  for (int64_t i = 0; i < cnt; i++) {
    int32_t nl = 0, hi = -1;
    double h[3];
    int f = orc_edge_nodes(m, a + 3 * i, b + 3 * i, clr, cdc, h, &hi, &nl);
    if (free_out) free_out[i] = (uint8_t)f;
    if (hit) memcpy(hit + 3 * i, h, sizeof h);
    if (hit_idx) hit_idx[i] = hi;
    if (lookups) lookups[i] = nl;
    total += nl;
  }
  return total;
}

int64_t orc_edge_guards_batch(const orc_map *m, const double *a, const double *b, int64_t cnt,
                              double clr, double cdc, uint8_t *free_out, int nthreads) {
  int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : total) \
    num_threads(nthreads > 1 ? nthreads : 1)
  for (int64_t i = 0; i < cnt; i++) {
    int32_t nl = 0;
    int f = orc_edge_guards(m, a + 3 * i, b + 3 * i, clr, cdc, &nl);
    if (free_out) free_out[i] = (uint8_t)f;
    total += nl;
  }
  return total;
}

int64_t orc_edge_index_batch(const orc_map *m, const double *nodes, const int32_t *from,
                             const int32_t *to, int64_t cnt, double clr, double cdc,
                             uint8_t *free_out, int32_t *hit_idx, int nthreads) {
  int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : total) \
    num_threads(nthreads > 1 ? nthreads : 1)
  for (int64_t i = 0; i < cnt; i++) {
    int32_t nl = 0, hi = -1;
    int f = orc_edge_nodes(m, nodes + 3 * (int64_t)from[i], nodes + 3 * (int64_t)to[i], clr, cdc,
                           NULL, &hi, &nl);
    if (free_out) free_out[i] = (uint8_t)f;
    if (hit_idx) hit_idx[i] = hi;
    total += nl;
  }
  return total;
}

/* ----------------------------------- polylines ---------------------------------------- */

double orc_path_length(const double *pts, int npts) { /* include/dijkstra.hpp:27-38 */
  double len = 0;
  for (int i = 1; i < npts; i++) len += v3_dist(pts + 3 * i, pts + 3 * (i - 1));
  return len;
}

int orc_sample_path(const double *pts, int npts, double length_tot, double num_samples,
                    double *out) {
  /* src/base_map.cpp:105-157.  NB: the cursor advances at most ONE segment per sample. */
  int cur = 0;
  double seg[3];
  v3_sub(pts + 3, pts, seg);
  double cur_len = v3_norm(seg);
  double cur_start = 0;
  int k = 0;
  for (double i = 0; i < num_samples; ++i, ++k) {
    double at = length_tot * (i / (num_samples - 1));
    if (at > cur_start + cur_len) {
      cur += 1;
      if (cur + 1 >= npts) {
        if (at - length_tot > PRECISION_BASE_MAP) {
          return -1; /* reference: ERROR + exit(1) */
        } else {
          cur -= 1;
          at = cur_start + cur_len;
        }
      } else {
        cur_start += cur_len;
        v3_sub(pts + 3 * (cur + 1), pts + 3 * cur, seg);
        cur_len = v3_norm(seg);
      }
    }
    const double *p0 = pts + 3 * cur, *p1 = pts + 3 * (cur + 1);
    double f = ((at - cur_start) / cur_len);
    out[3 * k + 0] = p0[0] + (p1[0] - p0[0]) * f;
    out[3 * k + 1] = p0[1] + (p1[1] - p0[1]) * f;
    out[3 * k + 2] = p0[2] + (p1[2] - p0[2]) * f;
  }
  return 0;
}

int orc_deformable(const orc_map *m, const double *p1, int n1, double len1, const double *p2,
                   int n2, double len2, double num_check, double clr, double cdc) {
  /* src/base_map.cpp:160-192 */
  int K = (int)num_check;
  if (K < 0) K = 0;
  double *s1 = (double *)malloc(sizeof(double) * 3 * (size_t)(K + 1));
  double *s2 = (double *)malloc(sizeof(double) * 3 * (size_t)(K + 1));
  int rc = 1;
  if (orc_sample_path(p1, n1, len1, num_check, s1) != 0 ||
      orc_sample_path(p2, n2, len2, num_check, s2) != 0) {
    rc = -1;
  } else {
    for (int i = 0; i < K; i++) {
      if (!orc_edge_nodes(m, s1 + 3 * i, s2 + 3 * i, clr, cdc, NULL, NULL, NULL)) {
        rc = 0;
        break;
      }
    }
  }
  free(s1);
  free(s2);
  return rc;
}

int orc_path_collision_free(const orc_map *m, const double *pts, int npts, double clr,
                            double cdc, double hit[3]) {
  /* src/base_map.cpp:200-214 */
  int ok = 1;
  hit[0] = hit[1] = hit[2] = NAN;
  for (int i = 1; i < npts; i++) {
    ok = orc_edge_nodes(m, pts + 3 * (i - 1), pts + 3 * i, clr, cdc, hit, NULL, NULL);
    if (!ok) return 0;
  }
  return ok;
}

int orc_deformable_between(const orc_map *m, const double s[3], const double e[3],
                           const double b1[3], const double b2[3], double clr, double cdc) {
  /* src/base_map.cpp:217-240 */
  double l1 = v3_dist(s, b1) + v3_dist(b1, e);
  double l2 = v3_dist(s, b2) + v3_dist(b2, e);
  double larger = l1 > l2 ? l1 : l2; /* std::max */
  double num = ceil(larger / cdc) + 1;
  double pa[9], pb[9];
  memcpy(pa, s, 24);
  memcpy(pa + 3, b1, 24);
  memcpy(pa + 6, e, 24);
  memcpy(pb, s, 24);
  memcpy(pb + 3, b2, 24);
  memcpy(pb + 6, e, 24);
  return orc_deformable(m, pa, 3, l1, pb, 3, l2, num, clr, cdc);
}

/* ------------------------------------ sampling ---------------------------------------- */

void orc_ellipsoid_point(const double A[3], const double B[3], double two_major, double u,
                         const double g[3], double out[3]) {
  /* src/common.cpp:256-312.  C is a proper rotation with C*e1 = a1; the reference obtains
   * it from JacobiSVD(a1*e1^T) (absent third-party code).  Restated with an explicit
   * orthonormal completion: the two minor semi-axes are equal, so any completion yields the
   * same distribution (parity unpinned at Eigen, see header). */
  double ab[3];
  v3_sub(B, A, ab);
  double two_focal = v3_norm(ab);
  double a1[3] = {ab[0] / two_focal, ab[1] / two_focal, ab[2] / two_focal};
  double centre[3] = {(A[0] + B[0]) / 2.0, (A[1] + B[1]) / 2.0, (A[2] + B[2]) / 2.0};
  /* completion: pick the world axis least aligned with a1 */
  int k = 0;
  if (fabs(a1[1]) < fabs(a1[k])) k = 1;
  if (fabs(a1[2]) < fabs(a1[k])) k = 2;
  double ek[3] = {0, 0, 0};
  ek[k] = 1.0;
  double a2[3], a3[3];
  v3_cross(a1, ek, a2);
  v3_normalize(a2);
  v3_cross(a1, a2, a3);
  double axis = sqrt(two_major * two_major - two_focal * two_focal);
  double L0 = two_major / 2.0, L1 = axis / 2.0;
  double radii = 1.0 * pow(u, 1.0 / 3.0);
  double gn = v3_norm(g);
  double ball[3] = {g[0] * radii / gn, g[1] * radii / gn, g[2] * radii / gn};
  double s0 = L0 * ball[0], s1 = L1 * ball[1], s2 = L1 * ball[2];
  for (int i = 0; i < 3; i++) out[i] = ((a1[i] * s0 + a2[i] * s1) + a3[i] * s2) + centre[i];
}

void orc_fill_random_state_in_ellipse(const double start[3], const double goal[3],
                                      double ratio, int planar, double u, const double g[3],
                                      double out[3]) {
  /* src/base_map.cpp:29-43 */
  double dist = v3_dist(goal, start);
  orc_ellipsoid_point(start, goal, dist * ratio, u, g, out);
  if (planar) out[2] = start[2];
}

int orc_sample_reject(const orc_map *m, const double start[3], const double goal[3],
                      const double *cand, int64_t M, double clr, int n, double *nodes,
                      uint8_t *accept, int64_t *n_used) {
  /* include/topological_prm_clustering.hpp:299-327 */
  memcpy(nodes, start, 24);
  memcpy(nodes + 3, goal, 24);
  int filled = 2;
  int64_t c = 0;
  for (; c < M && filled < n; c++) {
    int coll = orc_in_collision(m, cand + 3 * c, clr);
    if (accept) accept[c] = (uint8_t)!coll;
    if (!coll) {
      memcpy(nodes + 3 * (size_t)filled, cand + 3 * c, 24);
      filled++;
    }
  }
  if (n_used) *n_used = c;
  return filled;
}

/* --------------------------------- k-NN + PRM build ----------------------------------- */

void orc_knn(const double *pts64, int n, int k, int32_t *idx, float *d2, int nthreads) {
  /* include/topological_prm_clustering.hpp:300-339 restated as exact float32 k-NN.
   * FLANN L2<float> accumulates diff*diff in float, axis order x,y,z. */
  float *p = (float *)malloc(sizeof(float) * 3 * (size_t)n);
  for (size_t i = 0; i < (size_t)n * 3; i++) p[i] = (float)pts64[i];
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 1 ? nthreads : 1)
  for (int i = 0; i < n; i++) {
    float bd[64];
    int32_t bi[64];
    int cnt = 0;
    float xi = p[3 * i], yi = p[3 * i + 1], zi = p[3 * i + 2];
    for (int j = 0; j < n; j++) {
      if (j == i) continue; /* slot 0 "is the same point" (:361) */
      float dx = xi - p[3 * j], dy = yi - p[3 * j + 1], dz = zi - p[3 * j + 2];
      float dd = (dx * dx + dy * dy) + dz * dz;
      if (cnt == k && !(dd < bd[k - 1])) continue; /* j ascending: ties keep lower index */
      int pos = cnt < k ? cnt : k - 1;
      while (pos > 0 && bd[pos - 1] > dd) {
        bd[pos] = bd[pos - 1];
        bi[pos] = bi[pos - 1];
        pos--;
      }
      bd[pos] = dd;
      bi[pos] = j;
      if (cnt < k) cnt++;
    }
    for (int s = 0; s < k; s++) {
      idx[(size_t)i * k + s] = s < cnt ? bi[s] : -1;
      if (d2) d2[(size_t)i * k + s] = s < cnt ? bd[s] : INFINITY;
    }
  }
  free(p);
}

/* Same result as orc_knn (k smallest by (d2, index), self excluded), found through a uniform
 * grid instead of the O(n^2) scan, so the CPU baseline is not charged for a neighbour search
 * the reference does with a KD-forest.  tests/test_oracle_prm.py checks it against orc_knn. */
static inline int knn_less(float da, int32_t ia, float db, int32_t ib) {
  return da < db || (da == db && ia < ib);
}

void orc_knn_grid(const double *pts64, int n, int k, int32_t *idx, float *d2, int nthreads) {
  if (n < 64 || k > 64) {
    orc_knn(pts64, n, k, idx, d2, nthreads);
    return;
  }
  float *p = (float *)malloc(sizeof(float) * 3 * (size_t)n);
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      float v = (float)pts64[3 * (size_t)i + a];
      p[3 * (size_t)i + a] = v;
      if (v < lo[a]) lo[a] = v;
      if (v > hi[a]) hi[a] = v;
    }
  double ext[3], emax = 0;
  for (int a = 0; a < 3; a++) {
    ext[a] = (double)hi[a] - (double)lo[a];
    if (!(ext[a] >= 0)) ext[a] = 0; /* NaN coordinates: everything lands in cell 0 */
    if (ext[a] > emax) emax = ext[a];
  }
  if (!(emax > 0) || !isfinite(emax)) {
    free(p);
    orc_knn(pts64, n, k, idx, d2, nthreads);
    return;
  }
  double vol = 1;
  for (int a = 0; a < 3; a++) vol *= ext[a] > emax * 1e-6 ? ext[a] : emax * 1e-6;
  double h = cbrt(vol / ((double)n / 2.0));
  int dim[3];
  for (;;) {
    double cells = 1;
    for (int a = 0; a < 3; a++) {
      double c = floor(ext[a] / h) + 1;
      dim[a] = c > 2048 ? 2048 : (int)c;
      cells *= dim[a];
    }
    if (cells <= 4.0 * n + 64) break;
    h *= 1.25;
  }
  const int ncell = dim[0] * dim[1] * dim[2];
  int32_t *cell_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  int32_t *start = (int32_t *)calloc((size_t)ncell + 1, sizeof(int32_t));
  int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  for (int i = 0; i < n; i++) {
    int c[3];
    for (int a = 0; a < 3; a++) {
      double t = floor(((double)p[3 * (size_t)i + a] - (double)lo[a]) / h);
      c[a] = !(t >= 0) ? 0 : (t >= dim[a] ? dim[a] - 1 : (int)t);
    }
    cell_of[i] = (c[2] * dim[1] + c[1]) * dim[0] + c[0];
    start[cell_of[i] + 1]++;
  }
  for (int c = 0; c < ncell; c++) start[c + 1] += start[c];
  int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * (size_t)ncell);
  memcpy(cur, start, sizeof(int32_t) * (size_t)ncell);
  for (int i = 0; i < n; i++) order[cur[cell_of[i]]++] = i;
  free(cur);
  int rmax = dim[0] > dim[1] ? dim[0] : dim[1];
  if (dim[2] > rmax) rmax = dim[2];
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 1 ? nthreads : 1)
  for (int i = 0; i < n; i++) {
    float bd[64];
    int32_t bi[64];
    int cnt = 0;
    const float xi = p[3 * (size_t)i], yi = p[3 * (size_t)i + 1], zi = p[3 * (size_t)i + 2];
    const int ci = cell_of[i];
    const int cx = ci % dim[0], cy = (ci / dim[0]) % dim[1], cz = ci / (dim[0] * dim[1]);
    for (int r = 0; r <= rmax; r++) {
      for (int z = cz - r; z <= cz + r; z++) {
        if (z < 0 || z >= dim[2]) continue;
        for (int y = cy - r; y <= cy + r; y++) {
          if (y < 0 || y >= dim[1]) continue;
          const int shell_yz = (z == cz - r) || (z == cz + r) || (y == cy - r) || (y == cy + r);
          for (int x = cx - r; x <= cx + r; x += (shell_yz || r == 0) ? 1 : 2 * r) {
            if (x < 0 || x >= dim[0]) continue;
            const int c = (z * dim[1] + y) * dim[0] + x;
            for (int s = start[c]; s < start[c + 1]; s++) {
              const int j = order[s];
              if (j == i) continue;
              const float dx = xi - p[3 * (size_t)j], dy = yi - p[3 * (size_t)j + 1], dz = zi - p[3 * (size_t)j + 2];
              const float dd = (dx * dx + dy * dy) + dz * dz;
              if (cnt == k && !knn_less(dd, j, bd[k - 1], bi[k - 1])) continue;
              int pos = cnt < k ? cnt : k - 1;
              while (pos > 0 && knn_less(dd, j, bd[pos - 1], bi[pos - 1])) {
                bd[pos] = bd[pos - 1];
                bi[pos] = bi[pos - 1];
                pos--;
              }
              bd[pos] = dd;
              bi[pos] = j;
              if (cnt < k) cnt++;
            }
          }
        }
      }
      /* every unvisited point is more than r*h away along some axis (0.1% slack for the
       * float binning); strict '<' so equal-distance ties are never cut off */
      const double lb = (double)r * h * 0.999;
      if (cnt == k && (double)bd[k - 1] < lb * lb) break;
    }
    for (int s = 0; s < k; s++) {
      idx[(size_t)i * k + s] = s < cnt ? bi[s] : -1;
      if (d2) d2[(size_t)i * k + s] = s < cnt ? bd[s] : INFINITY;
    }
  }
  free(p);
  free(cell_of);
  free(start);
  free(order);
}

static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

/* symmetric union of the free directed edges as CSR (cols ascending):
 * visibility_node_ids[to] = visibility_node_ids[from] = norm (:378,384) */
int64_t orc_csr_from_directed(const double *nodes, int n, int k, const int32_t *nbr,
                              const uint8_t *directed_free, int32_t *row_ptr, int32_t *col,
                              double *w) {
  int32_t *deg = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
  for (int64_t e = 0; e < (int64_t)n * k; e++)
    if (directed_free[e]) {
      deg[e / k]++;
      deg[nbr[e]]++;
    }
  int32_t *tmp_ptr = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 1));
  tmp_ptr[0] = 0;
  for (int i = 0; i < n; i++) tmp_ptr[i + 1] = tmp_ptr[i] + deg[i];
  int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(tmp_ptr[n] > 0 ? tmp_ptr[n] : 1));
  memset(deg, 0, sizeof(int32_t) * ((size_t)n + 1));
  for (int64_t e = 0; e < (int64_t)n * k; e++)
    if (directed_free[e]) {
      int a = (int)(e / k), b = nbr[e];
      tmp[tmp_ptr[a] + deg[a]++] = b;
      tmp[tmp_ptr[b] + deg[b]++] = a;
    }
  int64_t nnz = 0;
  row_ptr[0] = 0;
  for (int i = 0; i < n; i++) {
    int32_t *r = tmp + tmp_ptr[i];
    int len = deg[i];
    qsort(r, (size_t)len, sizeof(int32_t), cmp_i32);
    int32_t prev = -1;
    for (int s = 0; s < len; s++) {
      if (r[s] == prev) continue;
      prev = r[s];
      col[nnz] = r[s];
      /* weight = (to->data - from->data).norm() on the double coordinates (:369-371) */
      w[nnz] = v3_dist(nodes + 3 * (size_t)r[s], nodes + 3 * (size_t)i);
      nnz++;
    }
    row_ptr[i + 1] = (int32_t)nnz;
  }
  free(deg);
  free(tmp_ptr);
  free(tmp);
  return nnz;
}

int64_t orc_build_prm(const orc_map *m, const double *nodes, int n, int k, double clr,
                      double cdc, int32_t *nbr, uint8_t *directed_free, int32_t *row_ptr,
                      int32_t *col, double *w, int64_t *lookups, int nthreads) {
  /* include/topological_prm_clustering.hpp:329-397 */
  orc_knn_grid(nodes, n, k, nbr, NULL, nthreads);
  int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : total) \
    num_threads(nthreads > 1 ? nthreads : 1)
  for (int64_t e = 0; e < (int64_t)n * k; e++) {
    int from = (int)(e / k);
    int to = nbr[e];
    int32_t nl = 0;
    int f = 0;
    if (to >= 0) f = orc_edge_nodes(m, nodes + 3 * (size_t)from, nodes + 3 * (size_t)to, clr, cdc,
                                     NULL, NULL, &nl);
    directed_free[e] = (uint8_t)f;
    total += nl;
  }
  if (lookups) *lookups = total;
  return orc_csr_from_directed(nodes, n, k, nbr, directed_free, row_ptr, col, w);
}

/* Whole sampleMultiple (:288-401) for q independent queries, OpenMP over queries: the CPU
 * "port" baseline bench.py times.  cand: q x M x 3.  Per query outputs: n_filled, number of
 * free directed edges, nnz, lookups; optional full outputs (may be NULL): nodes q x n x 3,
 * row_ptr q x (n+1), col/w q x 2nk.  returns the number of queries that ran out of candidates. */
int orc_sample_multiple_batch(const orc_map *m, const double *start, const double *goal,
                              const double *cand, int q, int64_t M, int n, int k, double clr,
                              double cdc, int nthreads, int32_t *n_filled, int64_t *n_free,
                              int64_t *nnz, int64_t *lookups, double *nodes_out,
                              int32_t *row_ptr_out, int32_t *col_out, double *w_out) {
  int short_q = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : short_q) \
    num_threads(nthreads > 1 ? nthreads : 1)
  for (int qi = 0; qi < q; qi++) {
    const size_t cap = 2 * (size_t)n * k;
    double *nodes = nodes_out ? nodes_out + 3 * (size_t)qi * n : (double *)malloc(24 * (size_t)n);
    int32_t *nbr = (int32_t *)malloc(4 * (size_t)n * k);
    uint8_t *fr = (uint8_t *)malloc((size_t)n * k);
    int32_t *rp = row_ptr_out ? row_ptr_out + (size_t)qi * (n + 1) : (int32_t *)malloc(4 * ((size_t)n + 1));
    int32_t *col = col_out ? col_out + (size_t)qi * cap : (int32_t *)malloc(4 * cap);
    double *w = w_out ? w_out + (size_t)qi * cap : (double *)malloc(8 * cap);
    int64_t used = 0, lk = 0;
    const int filled = orc_sample_reject(m, start + 3 * (size_t)qi, goal + 3 * (size_t)qi,
                                         cand + 3 * (size_t)qi * M, M, clr, n, nodes, NULL, &used);
    n_filled[qi] = filled;
    if (filled < n) {
      short_q++;
      n_free[qi] = nnz[qi] = 0;
      if (lookups) lookups[qi] = 0;
    } else {
      nnz[qi] = orc_build_prm(m, nodes, n, k, clr, cdc, nbr, fr, rp, col, w, &lk, 1);
      int64_t f = 0;
      for (size_t e = 0; e < (size_t)n * k; e++) f += fr[e];
      n_free[qi] = f;
      if (lookups) lookups[qi] = lk;
    }
    if (!nodes_out) free(nodes);
    free(nbr);
    free(fr);
    if (!row_ptr_out) free(rp);
    if (!col_out) free(col);
    if (!w_out) free(w);
  }
  return short_q;
}

/* ---------------------------------- path shortening ----------------------------------- */

int orc_shorten_path(const orc_map *m, const double *pts, int npts, double length, int forward,
                     double clr, double cdc, double *out, int32_t *origin, int cap,
                     double *out_length, int *status, int64_t *lookups) {
  /* include/topological_prm_clustering.hpp:2313-2479 */
  int64_t nl_total = 0;
  *status = 0;
  double num_samples = ceil(length / cdc) + 1;
  int S = (int)num_samples;
  if (S < 0) S = 0;
  double *sampled = (double *)malloc(sizeof(double) * 3 * (size_t)(S + 1));
  int nout = 0;
  double slen = 0;
#define RETURN_ORIGINAL(code)                                   \
  do {                                                          \
    int c_ = npts < cap ? npts : cap;                           \
    memcpy(out, pts, sizeof(double) * 3 * (size_t)c_);          \
    for (int q_ = 0; q_ < c_; q_++) origin[q_] = q_;            \
    *out_length = length;                                       \
    *status = (code);                                           \
    if (lookups) *lookups = nl_total;                           \
    free(sampled);                                              \
    return c_;                                                  \
  } while (0)
  if (orc_sample_path(pts, npts, length, num_samples, sampled) != 0) RETURN_ORIGINAL(-1);
  if (!forward) { /* std::reverse(sampled) (:2344) */
    for (int i = 0, j = S - 1; i < j; i++, j--)
      for (int c = 0; c < 3; c++) {
        double t = sampled[3 * i + c];
        sampled[3 * i + c] = sampled[3 * j + c];
        sampled[3 * j + c] = t;
      }
  }
  int first = forward ? 0 : npts - 1;
  memcpy(out, pts + 3 * (size_t)first, 24);
  origin[0] = first;
  nout = 1;
  for (int i = 1; i < S; i++) {
    const double *anchor = out + 3 * (size_t)(nout - 1);
    double hit[3];
    int32_t nl = 0;
    int is_free = orc_edge_nodes(m, anchor, sampled + 3 * i, clr, cdc, hit, NULL, &nl);
    nl_total += nl;
    if (!is_free) {
      double grad[3], vc[3];
      orc_gradient(m, hit, grad, vc);
      nl_total += 6;
      double npv[3], ngn[3], vfc[3];
      v3_sub(sampled + 3 * i, anchor, npv);
      v3_cross(grad, npv, ngn);
      v3_cross(npv, ngn, vfc);
      v3_normalize(vfc);
      double ds = cdc;
      int added = 0;
      for (double tci = clr; tci <= clr * 4; tci += ds) {
        double np[3] = {vc[0] + vfc[0] * tci, vc[1] + vfc[1] * tci, vc[2] + vfc[2] * tci};
        nl_total++;
        if (!orc_in_collision(m, np, clr)) {
          if (nout >= cap - 1) RETURN_ORIGINAL(-3);
          double dist_new = v3_dist(anchor, np);
          slen += dist_new;
          memcpy(out + 3 * (size_t)nout, np, 24);
          origin[nout] = -1;
          nout++;
          added = 1;
          break;
        }
      }
      if (!added) RETURN_ORIGINAL(1); /* "no point added ..." -> return path (:2423-2427) */
    }
  }
  int last = forward ? npts - 1 : 0;
  if (nout >= cap) RETURN_ORIGINAL(-3);
  slen += v3_dist(out + 3 * (size_t)(nout - 1), pts + 3 * (size_t)last);
  memcpy(out + 3 * (size_t)nout, pts + 3 * (size_t)last, 24);
  origin[nout] = last;
  nout++;
  if (!forward) {
    for (int i = 0, j = nout - 1; i < j; i++, j--) {
      for (int c = 0; c < 3; c++) {
        double t = out[3 * i + c];
        out[3 * i + c] = out[3 * j + c];
        out[3 * j + c] = t;
      }
      int32_t t = origin[i];
      origin[i] = origin[j];
      origin[j] = t;
    }
  }
  /* self-check (:2463-2473): reference exits when it fails */
  double calc = orc_path_length(out, nout);
  const double *lo = out + 3 * (size_t)(nout - 1), *li = pts + 3 * (size_t)(npts - 1);
  if (fabs(calc - slen) > PRECISION_SHORTEN || lo[0] != li[0] || lo[1] != li[1] || lo[2] != li[2])
    *status = -2;
  *out_length = slen;
  if (lookups) *lookups = nl_total;
  free(sampled);
  return nout;
#undef RETURN_ORIGINAL
}

int orc_remove_equivalent(const orc_map *m, const double *pts, const int32_t *offs,
                          const double *lengths, int P, double clr, double cdc,
                          int32_t *keep_src) {
  /* include/topological_prm_clustering.hpp:2237-2298, replayed on indices */
  int cnt = P;
  for (int i = 0; i < P; i++) keep_src[i] = i;
  if (cnt > 1) {
    int i = 0;
    int32_t *rm = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
    while (i < cnt) {
      int shortest = i;
      double shortest_len = lengths[keep_src[i]];
      int nrm = 0;
      for (int j = i + 1; j < cnt; j++) {
        int a = keep_src[i], b = keep_src[j];
        double larger = lengths[a] > lengths[b] ? lengths[a] : lengths[b];
        double num = ceil(larger / cdc) + 1;
        int def = orc_deformable(m, pts + 3 * (size_t)offs[a], offs[a + 1] - offs[a], lengths[a],
                                 pts + 3 * (size_t)offs[b], offs[b + 1] - offs[b], lengths[b], num,
                                 clr, cdc);
        if (def == 1) {
          rm[nrm++] = j;
          if (lengths[b] < shortest_len) {
            shortest = j;
            shortest_len = lengths[b];
          }
        }
      }
      if (shortest != i) keep_src[i] = keep_src[shortest];
      for (int t = nrm - 1; t >= 0; t--) {
        for (int q = rm[t]; q + 1 < cnt; q++) keep_src[q] = keep_src[q + 1];
        cnt--;
      }
      i++;
    }
    free(rm);
  }
  return cnt;
}

/* ------------------------------------ graph search ------------------------------------ */
/* Dijkstra over a CSR graph (binary heap), restating include/dijkstra.hpp:83-176 with every node as a
 * potential goal (no early stop) and any number of zero-distance sources (the multi-source flood of
 * wavefrontFill, include/topological_prm_clustering.hpp:789-962).  dist accumulates
 * distance_from_start + edge in FP64 exactly as expandBestNode does.  pred / label follow the
 * deterministic rule of include/ctopprm_b200.h (ctp_sssp): smallest strictly nearer neighbour that
 * realises the distance; label = source index at the root of the chain. */
typedef struct { double d; int32_t v; } orc_hitem;

static void heap_push(orc_hitem *h, int *cnt, double d, int32_t v) {
  int i = (*cnt)++;
  h[i].d = d;
  h[i].v = v;
  while (i > 0) {
    int p = (i - 1) / 2;
    if (h[p].d <= h[i].d) break;
    orc_hitem t = h[p];
    h[p] = h[i];
    h[i] = t;
    i = p;
  }
}

static orc_hitem heap_pop(orc_hitem *h, int *cnt) {
  orc_hitem top = h[0];
  h[0] = h[--(*cnt)];
  int i = 0;
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    if (l < *cnt && h[l].d < h[m].d) m = l;
    if (r < *cnt && h[r].d < h[m].d) m = r;
    if (m == i) break;
    orc_hitem t = h[m];
    h[m] = h[i];
    h[i] = t;
    i = m;
  }
  return top;
}

void orc_sssp(int n, const int32_t *row_ptr, const int32_t *col, const double *w, const int32_t *src, int ns,
              double *dist, int32_t *pred, int32_t *label) {
  int64_t nnz = row_ptr[n];
  orc_hitem *heap = (orc_hitem *)malloc(sizeof(orc_hitem) * (size_t)(nnz + ns + 1));
  int cnt = 0;
  for (int i = 0; i < n; i++) dist[i] = INFINITY;
  for (int i = 0; i < ns; i++)
    if (src[i] >= 0 && src[i] < n && dist[src[i]] != 0.0) {
      dist[src[i]] = 0.0;
      heap_push(heap, &cnt, 0.0, src[i]);
    }
  while (cnt > 0) {
    orc_hitem it = heap_pop(heap, &cnt);
    if (it.d > dist[it.v]) continue;
    for (int32_t e = row_ptr[it.v]; e < row_ptr[it.v + 1]; e++) {
      double nd = dist[it.v] + w[e];
      if (nd < dist[col[e]]) {
        dist[col[e]] = nd;
        heap_push(heap, &cnt, nd, col[e]);
      }
    }
  }
  free(heap);
  for (int v = 0; v < n; v++) {
    int best = -1;
    if (isfinite(dist[v]) && dist[v] != 0.0)
      for (int32_t e = row_ptr[v]; e < row_ptr[v + 1]; e++) {
        int u = col[e];
        if (!(dist[u] < dist[v])) continue;
        if (dist[u] + w[e] == dist[v] && (best < 0 || u < best)) best = u;
      }
    pred[v] = best;
  }
  if (label) {
    for (int v = 0; v < n; v++) label[v] = -1;
    for (int i = ns - 1; i >= 0; i--)
      if (src[i] >= 0 && src[i] < n) label[src[i]] = i; /* listed twice: the first index wins */
    for (int v = 0; v < n; v++) {
      if (!isfinite(dist[v]) || dist[v] == 0.0) continue;
      int u = pred[v];
      while (u >= 0 && pred[u] >= 0) u = pred[u];
      label[v] = u >= 0 ? label[u] : -1;
    }
  }
}


/* ---- exact EDT by definition (checker for the device EDT; see ctp_oracle.h) ------------------------- */
static void edt_axis_min(const int64_t *g, int64_t *out, int n, size_t stride, size_t n_lines_a, size_t stride_a,
                         size_t n_lines_b, size_t stride_b) {
  for (size_t a = 0; a < n_lines_a; a++)
    for (size_t b = 0; b < n_lines_b; b++) {
      const size_t base = a * stride_a + b * stride_b;
      for (int u = 0; u < n; u++) {
        int64_t best = INT64_MAX;
        for (int i = 0; i < n; i++) {
          const int64_t v = (int64_t)(u - i) * (u - i) + g[base + (size_t)i * stride];
          if (v < best) best = v;
        }
        out[base + (size_t)u * stride] = best;
      }
    }
}

void orc_edt_signed(const unsigned char *occ, int n, double pitch, float *vox, int32_t *d2s) {
  const size_t cnt = (size_t)n * n * n;
  const int64_t INF = (int64_t)1 << 40;
  int64_t *a = (int64_t *)malloc(cnt * sizeof(int64_t));
  int64_t *b = (int64_t *)malloc(cnt * sizeof(int64_t));
  for (int feat = 1; feat >= 0; feat--) {
    for (size_t o = 0; o < cnt; o++) a[o] = ((occ[o] != 0) == (feat != 0)) ? 0 : INF;
    edt_axis_min(a, b, n, 1, (size_t)n, (size_t)n * n, (size_t)n, (size_t)n);            /* along z */
    edt_axis_min(b, a, n, (size_t)n, (size_t)n, (size_t)n * n, (size_t)n, 1);            /* along y */
    edt_axis_min(a, b, n, (size_t)n * n, (size_t)n, (size_t)n, (size_t)n, 1);            /* along x */
    for (size_t o = 0; o < cnt; o++) {
      if ((occ[o] != 0) == (feat != 0)) continue;
      int64_t v = b[o];
      if (v >= INF) v = 3 * (int64_t)n * n;
      const double d = pitch * sqrt((double)v);
      if (vox) vox[o] = (float)(feat ? d : -d);
      if (d2s) d2s[o] = (int32_t)(feat ? v : -v);
    }
  }
  free(a);
  free(b);
}
