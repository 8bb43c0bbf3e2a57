/*
 * ctp_oracle.h -- CPU restatement of CTopPRM's ESDF / collision / PRM-build hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ctopprm_learn_b200/, include/)
 * may include, link or call this file; it exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs have a checker and a CPU baseline.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference tree: src/esdf_map.cpp, src/base_map.cpp, src/common.cpp,
 * include/topological_prm_clustering.hpp).  FP64 throughout, grid values stored as f32,
 * compiled with -O2 -ffp-contract=off so no fused multiply-adds are formed.
 *
 * Parity pinning: oracle/_ref (the reference's own src/esdf_map.cpp and src/base_map.cpp
 * compiled against shim headers, see oracle/Makefile) is compared against this
 * restatement by tests/test_oracle_vs_ref.py and tools/make_golden.py; the resulting
 * vectors are committed under tests/golden/.  Pieces of the path that live in absent
 * third-party code are "parity unpinned" and are restated from their published behaviour:
 *   - Eigen (version not pinned by the reference): Vector<3>::norm() reduction order is
 *     taken as ((x*x + y*y) + z*z); JacobiSVD in random_ellipsoid_point only fixes the
 *     orientation of the two minor axes, which are equal, so the distribution is the same.
 *   - FLANN (version not pinned): KDTreeIndexParams(4)/SearchParams(128) is approximate and
 *     randomized; restated here as exact float32 k-NN with (distance, index) tie-break.
 *   - libstdc++ default_random_engine seeded with time(NULL): replaced by a shared
 *     candidate-sample file (north_star: "fed from a shared sample file").
 */
#ifndef CTP_ORACLE_H
#define CTP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_map {
  const float *vox; /* flat, index x*n*n + y*n + z  (include/esdf_map.hpp:55-66) */
  int n;            /* voxels per axis = shape[0]   (src/esdf_map.cpp:24)        */
  double N;         /* num_vexels_per_axis is a double in the reference (esdf_map.hpp:46) */
  double c[3];      /* centroid */
  double e[3];      /* extents */
  double max_e;     /* extents.maxCoeff() (src/esdf_map.cpp:27) */
} orc_map;

void orc_map_init(orc_map *m, const float *vox, int n, const double c[3], const double e[3]);

/* src/esdf_map.cpp:75-76 */
void orc_min_pos(const orc_map *m, double out[3]);
void orc_max_pos(const orc_map *m, double out[3]);
/* src/esdf_map.cpp:78-121 */
double orc_interpolate(const orc_map *m, const double q[3]);
/* src/esdf_map.cpp:130-165 */
double orc_clearance(const orc_map *m, const double p[3]);
/* src/esdf_map.cpp:123-128 */
void orc_real_pos_from_index(const orc_map *m, const double q[3], double out[3]);
/* src/esdf_map.cpp:167-227 */
void orc_gradient(const orc_map *m, const double p[3], double grad[3], double centre[3]);
/* src/base_map.cpp:7-12 */
int orc_in_collision(const orc_map *m, const double p[3], double clr);
/* src/base_map.cpp:49-74.  returns 1 if free.  hit = first colliding sample or NaN^3,
 * hit_idx = its index (1-based march index) or -1, lookups = clearance evaluations made. */
int orc_edge_nodes(const orc_map *m, const double a[3], const double b[3], double clr,
                   double cdc, double hit[3], int32_t *hit_idx, int32_t *lookups);
/* src/base_map.cpp:77-97 */
int orc_edge_guards(const orc_map *m, const double a[3], const double b[3], double clr,
                    double cdc, int32_t *lookups);

/* batch drivers (nthreads<=1: plain loop; otherwise OpenMP over items) */
void orc_clearance_batch(const orc_map *m, const double *xyz, int64_t cnt, double *out,
                         int nthreads);
void orc_gradient_batch(const orc_map *m, const double *xyz, int64_t cnt, double *grad,
                        double *centre, int nthreads);
/* returns total lookups */
int64_t orc_edge_nodes_batch(const orc_map *m, const double *a, const double *b, int64_t cnt,
                             double clr, double cdc, uint8_t *free_out, double *hit,
                             int32_t *hit_idx, int32_t *lookups, int nthreads);
int64_t orc_edge_guards_batch(const orc_map *m, const double *a, const double *b, int64_t cnt,
                              double clr, double cdc, uint8_t *free_out, int nthreads);
/* edges given as index pairs into a node table */
int64_t orc_edge_index_batch(const orc_map *m, const double *nodes, const int32_t *from,
                             const int32_t *to, int64_t cnt, double clr, double cdc,
                             uint8_t *free_out, int32_t *hit_idx, int nthreads);

/* src/base_map.cpp:105-157.  pts: npts x 3, out: (int)num_samples x 3.
 * returns 0, or -1 where the reference would exit(1) (cursor overrun > PRECISION_BASE_MAP). */
int orc_sample_path(const double *pts, int npts, double length_tot, double num_samples,
                    double *out);
/* src/base_map.cpp:160-192; returns 1 deformable, 0 not, -1 samplePath failure */
int orc_deformable(const orc_map *m, const double *p1, int n1, double len1, const double *p2,
                   int n2, double len2, double num_check, double clr, double cdc);
/* src/base_map.cpp:200-214; returns 1 if the polyline is free; hit as orc_edge_nodes */
int orc_path_collision_free(const orc_map *m, const double *pts, int npts, double clr,
                            double cdc, double hit[3]);
/* src/base_map.cpp:217-240 */
int orc_deformable_between(const orc_map *m, const double s[3], const double e[3],
                           const double b1[3], const double b2[3], double clr, double cdc);
/* include/dijkstra.hpp:27-38 */
double orc_path_length(const double *pts, int npts);

/* src/common.cpp:256-312 with the four random draws passed in (u ~ U(0,1), g ~ N(0,1)^3);
 * src/base_map.cpp:29-43 planar override */
void orc_ellipsoid_point(const double A[3], const double B[3], double two_major, double u,
                         const double g[3], double out[3]);
void orc_fill_random_state_in_ellipse(const double start[3], const double goal[3],
                                      double ratio, int planar, double u, const double g[3],
                                      double out[3]);

/* include/topological_prm_clustering.hpp:299-327: rejection loop replayed over a candidate
 * stream.  nodes: n x 3, rows 0,1 = start, goal (not tested), rows 2.. = first accepted
 * candidates in stream order.  accept (optional, len M) = per-candidate bit for the prefix
 * that was consumed.  returns number of nodes filled (== n on success); n_used = candidates
 * consumed. */
int orc_sample_reject(const orc_map *m, const double start[3], const double goal[3],
                      const double *cand, int64_t M, double clr, int n, double *nodes,
                      uint8_t *accept, int64_t *n_used);

/* :300-339 restated as exact float32 k-NN (self excluded, k others, sorted by (d2, idx)).
 * pts64: n x 3 doubles, narrowed to float exactly as :300-326 does.  idx: n x k, d2: n x k. */
void orc_knn(const double *pts64, int n, int k, int32_t *idx, float *d2, int nthreads);

/* same result through a uniform grid (used by orc_build_prm; checked against orc_knn) */
void orc_knn_grid(const double *pts64, int n, int k, int32_t *idx, float *d2, int nthreads);
/* :374-385 symmetric union of the free directed edges as CSR; returns nnz */
int64_t orc_csr_from_directed(const double *nodes, int n, int k, const int32_t *nbr,
                              const uint8_t *directed_free, int32_t *row_ptr, int32_t *col,
                              double *w);
/* whole sampleMultiple (:288-401) for q queries, OpenMP over queries (CPU "port" baseline) */
int orc_sample_multiple_batch(const orc_map *m, const double *start, const double *goal,
                              const double *cand, int q, int64_t M, int n, int k, double clr,
                              double cdc, int nthreads, int32_t *n_filled, int64_t *n_free,
                              int64_t *nnz, int64_t *lookups, double *nodes_out,
                              int32_t *row_ptr_out, int32_t *col_out, double *w_out);

/* :355-397 directed checks + symmetric adjacency as CSR (cols ascending).
 * directed_free: n x k.  row_ptr: n+1, col/w sized for 2*n*k.  returns nnz. */
int64_t orc_build_prm(const orc_map *m, const double *nodes, int n, int k, double clr,
                      double cdc, int32_t *nbr, uint8_t *directed_free, int32_t *row_ptr,
                      int32_t *col, double *w, int64_t *lookups, int nthreads);

/* :2313-2479.  in: npts x 3 polyline with length.  out: up to cap x 3 points,
 * origin[i] = index into the input polyline for reused nodes, -1 for new push-off nodes.
 * returns number of output points; *status 0 ok, 1 = "no point added" (original returned),
 * -1 = samplePath failure, -2 = length self-check failed (reference exit(1)), -3 cap. */
int orc_shorten_path(const orc_map *m, const double *pts, int npts, double length, int forward,
                     double clr, double cdc, double *out, int32_t *origin, int cap,
                     double *out_length, int *status, int64_t *lookups);

/* :2237-2298.  paths given as a concatenated point table with offsets (P+1) and lengths (P).
 * keep_src[k] = index of the input path that ends up in output slot k.  returns count. */
int orc_remove_equivalent(const orc_map *m, const double *pts, const int32_t *offs,
                          const double *lengths, int P, double clr, double cdc,
                          int32_t *keep_src);

/* include/dijkstra.hpp:83-176 over a CSR graph, all nodes as goals, ns zero-distance sources;
 * pred / label by the deterministic rule documented for ctp_sssp (label may be NULL) */
void orc_sssp(int n, const int32_t *row_ptr, const int32_t *col, const double *w, const int32_t *src,
              int ns, double *dist, int32_t *pred, int32_t *label);

/* ---- cluster stage (ctp_oracle_cluster.c; include/topological_prm_clustering.hpp:589-1952) ---- */
/* one flood: mode 0 = wavefrontFill's multi-source flood (:826-962), mode 1 = addCentroid's re-flood from
 * seeds[nseeds-1] (:1276-1374) on top of the dist / prev / cluster arrays passed in.  Records (expanding node,
 * neighbour, distance) in the reference's append order; returns their number. */
int orc_cluster_flood(int n, const int32_t *rp, const int32_t *col, const double *w, const int32_t *seeds, int nseeds,
                      int mode, double *dist, int32_t *prev, int32_t *cluster, int cap, int32_t *rec_nodes,
                      int32_t *rec_label, double *rec_dist, int64_t *n_updates);
/* clusterGraph (:1832-1952): stitched cluster paths (points back to back, offsets, DFS lengths), final seeds and
 * labels.  Returns the number of paths, -1 capacity, -2 where the reference would exit(1). */
int orc_cluster_graph(const orc_map *m, int n, const double *xyz, const int32_t *rp, const int32_t *col, const double *w,
                      double clr, double cdc, int num_clusters, int max_clusters, int min_clusters, double max_path_length,
                      const int32_t *seed_candidates, int n_cand, int cap_paths, int cap_pts, double *out_pts,
                      int32_t *out_off, double *out_len, int32_t *seeds_out, int32_t *nseeds_out, int32_t *cluster_out);
/* re-floods of the last orc_cluster_graph call in which the node-count cap of addCentroid (:1319) fired */
int orc_cluster_capped_refloods(void);
/* removeTooLongPaths (:2193-2222) */
void orc_remove_too_long(const double *lengths, int P, double min_allowed, double cutoff_ratio, uint8_t *keep);

/* exact signed squared Euclidean distance transform of an n^3 occupancy grid (.npy order), by definition
 * (separable brute-force minimisation, O(n^4)): the checker for ctp_map_create_occupancy.  The map pipeline's
 * own distance field comes from mesh_to_sdf (map.py:49-73, not in this image): parity unpinned at that
 * boundary, pinned here against scipy.ndimage.distance_transform_edt in tests/test_oracle_prm.py.
 * vox (optional): pitch * sqrt(d2) signed, as f32; d2s (optional): signed squared voxel distances. */
void orc_edt_signed(const unsigned char *occ, int n, double pitch, float *vox, int32_t *d2s);

#ifdef __cplusplus
}
#endif
#endif
