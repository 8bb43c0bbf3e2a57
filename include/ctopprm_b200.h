/*
 * ctopprm_b200.h -- C ABI of libctopprm_b200.so: CTopPRM's data-parallel hot path
 * (ESDF clearance, segment collision checks, PRM build, polyline deformability and
 * shortening) as hand-written sm_100a CUDA kernels.
 *
 * The reference (Myracle1/CTopPRM_Learn, an annotated fork of ctu-mrs/CTopPRM) has no FFI
 * layer: its seam is the C++ class BaseMap/ESDFMap plus three TopologicalPRMClustering
 * methods.  Each entry point below names the reference symbol (file:line, relative to the
 * reference tree) it replaces; include/ctopprm/(asterisk).hpp re-creates the class surface on top.
 *
 * Conventions
 *   - every function returns CTP_OK (0) or a negative CTP_ERR_* code; the message of the last
 *     failure on the calling thread is available from ctp_last_error().  Nothing here ever
 *     calls exit() or throws across the boundary.
 *   - plain pointers and sizes only.  `ctp_x(...)` takes HOST buffers and is synchronous
 *     (copies in, launches, copies out, waits).  `ctp_x_dev(...)` takes DEVICE buffers that
 *     live on the map's device and only enqueues work on `stream` (a cudaStream_t passed as
 *     void*; NULL = the legacy default stream).
 *   - one handle per (map, device).  Calls on one handle share a scratch workspace and must
 *     be issued on one stream at a time; distinct handles may be driven from distinct host
 *     threads (one per GPU when work is sharded).
 *   - points are packed xyz doubles (AoS, 24 B), exactly the layout of the reference's
 *     Vector<3> (include/common.hpp:657-658).  NaN clearance = outside the map, and NaN^3 is
 *     the "no hit" sentinel, both preserved (src/base_map.cpp:10,58,73).
 *   - there is no CPU fallback: without a CUDA device every call fails with CTP_ERR_CUDA.
 */
#ifndef CTOPPRM_B200_H
#define CTOPPRM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTP_OK 0
#define CTP_ERR_INVALID (-1)       /* bad argument */
#define CTP_ERR_CUDA (-2)          /* CUDA runtime error (message has the cudaError string) */
#define CTP_ERR_NOMEM (-3)         /* device allocation failed */
#define CTP_ERR_NEED_CANDIDATES (-4) /* candidate stream exhausted before n nodes were accepted */
#define CTP_ERR_SAMPLE_PATH (-5)   /* samplePath cursor overrun: reference exit(1), src/base_map.cpp:127-133 */
#define CTP_ERR_CAPACITY (-6)      /* caller-provided output capacity too small */

/* device grid layouts (bit flags for ctp_map_create, single value for ctp_map_set_layout) */
#define CTP_LAYOUT_PLAIN 1 /* [x][y][z] f32 exactly as the .npy record (include/esdf_map.hpp:55-66) */
#define CTP_LAYOUT_CELL8 2 /* one 32-byte record per cell holding its 8 corners: 1 sector / lookup */
#define CTP_LAYOUT_TEX 4   /* 2-D gather texture (x-slab mosaic), 2x2 (y,z) footprints fetched with tld4; n <= 1024 */
#define CTP_LAYOUT_BRICK 8 /* 4x4x4-voxel bricks (256 B), keeps the plain footprint L2-friendly */

#define CTP_EDGE_NODES 0  /* BaseMap::isSimplePathFreeBetweenNodes  src/base_map.cpp:49-74 */
#define CTP_EDGE_GUARDS 1 /* BaseMap::isSimplePathFreeBetweenGuards src/base_map.cpp:77-97 */

typedef struct ctp_map ctp_map;

const char *ctp_last_error(void);
int ctp_version(void);
int ctp_device_count(int *count);

/* ---- ESDFMap::load + fill_data (src/esdf_map.cpp:6-72, include/esdf_map.hpp:55-66) ----
 * vox: n*n*n host floats in .npy order (x*n*n + y*n + z).  The f64 variant narrows to float
 * as fill_data<double> does.  layouts: OR of CTP_LAYOUT_* to materialise on the device
 * (PLAIN is always kept); the first set bit in the order TEX, CELL8, BRICK, PLAIN becomes the
 * active one unless changed with ctp_map_set_layout. */
int ctp_map_create(const float *vox, int n, const double centroid[3], const double extents[3],
                   int device, int layouts, ctp_map **out);
int ctp_map_create_f64(const double *vox, int n, const double centroid[3],
                       const double extents[3], int device, int layouts, ctp_map **out);
/* procedural map generated on the device (multi-map sweeps without a host grid): signed
 * distance to n_col vertical cylinders (centres cx, cy, radii r), constant along z, cubic extents.  Voxel i of
 * an axis sits at (centroid - E/2) + i * (E/N), the inverse of the index map of src/esdf_map.cpp:136-139. */
int ctp_map_create_columns(int n, const double centroid[3], const double extents[3], const double *cx,
                           const double *cy, const double *r, int n_col, int device, int layouts,
                           ctp_map **out);
/* map from a voxel occupancy grid (the voxel grid -> signed distance step of the map pipeline, map.py:49-73, as an
 * exact Euclidean distance transform on the device).  occ: n^3 bytes in .npy order, non-zero = occupied; cubic
 * extents, voxel pitch p = E/N.  Field: free voxel -> +p*sqrt(d2) to the nearest occupied voxel, occupied voxel ->
 * -p*sqrt(d2) to the nearest free voxel, d2 the exact integer squared voxel distance (3*n*n when the grid holds no
 * voxel of the other kind).  d2_signed (optional host buffer, n^3 int32): those squared distances, negative inside. */
int ctp_map_create_occupancy(const uint8_t *occ, int n, const double centroid[3], const double extents[3],
                             int device, int layouts, int32_t *d2_signed, ctp_map **out);
/* the f32 grid back on the host in .npy order (to write a map file in map.py's format) */
int ctp_map_download(ctp_map *map, float *vox);
int ctp_map_destroy(ctp_map *map);
int ctp_map_set_layout(ctp_map *map, int layout);
/* ESDFMap::getMinPos/getMaxPos (src/esdf_map.cpp:75-76) and bookkeeping.
 * info[0..2]=min, [3..5]=max, [6]=n, [7]=active layout, [8]=device, [9]=device bytes held,
 * [10]=lanes per edge used by ctp_build_prm, [11]=SM count */
int ctp_map_info(const ctp_map *map, double info[12]);
/* edge kernel used inside ctp_build_prm / ctp_sample_multiple: 1 (flat, default), 4, 8, 16 or 32 */
int ctp_set_prm_group(ctp_map *map, int group);
/* pre-size the scratch workspace so later _dev calls never allocate (needed before CUDA
 * graph capture).  q queries x n nodes x m candidates, k neighbours. */
int ctp_reserve(ctp_map *map, int64_t q, int64_t n, int64_t m, int k);
/* sum of ESDF lookups executed in reference order by edge-marching kernels since the last
 * reset (device counter; reading it synchronises the stream). */
int ctp_lookup_counter_read(ctp_map *map, void *stream, uint64_t *lookups, int reset);
/* out[0] = roadmap edges the whole-edge float32 filter decided with certainty (free, one lookup each), out[1] = roadmap
   edges it was shown, out[2] = the reference-order lookups (samples) of the decided edges, since the last reset; the other
   edges were marched by the exact FP64 kernel.  Same verdicts either way (base_map.cpp:49-74); a diagnostic for the
   bench line (lookups actually performed = out[1] + total lookups - out[2]). */
int ctp_filter_counter_read(ctp_map *map, void *stream, uint64_t out[3], int reset);

/* implementation knobs (defaults are the tuned ones; exposed for benchmarking) */
#define CTP_TUNE_KNN_PER_CELL 0 /* k-NN grid: expected points inside a ball of one cell size (default 22) */
#define CTP_TUNE_KNN_TILE 1     /* k-NN search: 1 = direct search with listed wider passes (default), 0 = per-thread ring search only */
#define CTP_TUNE_EDGE_ORDER 3   /* PRM edge kernel walks the rows in the k-NN grid's cell order (1), node order (0), automatic (-1, default) */
#define CTP_TUNE_SSSP_FRONTIER 4 /* ctp_sssp: 1 = explicit frontier, edge-parallel relaxation (default), 0 = flag scan */
#define CTP_TUNE_PIPE_HEAD 5     /* chunk pipelines: upload value*n (+256) candidates per query first, the tails only if a query
                                   of the chunk is still short (default 2.3; 0 = always upload whole streams) */
#define CTP_TUNE_SSSP_CLUSTER 6  /* ctp_sssp: one query per cluster of 8 or 16 CTAs, distances (and, with 16, the rows) in distributed shared memory: 1 whenever possible, 0 never, -1 automatic (default: few queries in flight, or a query too large for one CTA) */
#define CTP_TUNE_KNN_RADIUS 10   /* radius-neighbour candidates: value > 0 (metres) drops every neighbour farther than it from the k-NN table (slot = -1), so a node connects to its nearest nodes within the radius, at most k; 0 (default) = the reference's plain k nearest (:329-339) */
#define CTP_TUNE_SEG_BIN 9       /* ctp_edge_check: bin the segments of a batch by the 16^3-voxel block of their midpoint before marching them (measured: no gain, the kernel is not bound by memory on such batches): 1 always, 0 never and the layout as set, -1 (default) never, but batches of >= 65536 segments on a grid > 100 MB read the CELL8 records when they were materialised */
#define CTP_TUNE_CSR_PAIR_CAP 7  /* adjacency stage: capacity of the overflow list of rows with more than 16 incoming-only edges (0 = automatic: one pair per node; tests set it small to reach the full-scan fallback) */
#define CTP_TUNE_PERQ_CLUSTER 8  /* per-query setup / scan kernels on a cluster of 8 CTAs per query: 1 always, 0 never, -1 automatic (default: few queries of >= 16384 points) */
#define CTP_TUNE_EDGE_FILTER 11 /* roadmap edges (ctp_build_prm*, ctp_edge_check_knn_dev, ctp_prm_rows_dev): float32 whole-edge filter with rigorous error bounds in front of the exact FP64 kernel (a few lookups and a Lipschitz bound decide most free edges; the undecided ones are marched in FP64; verdicts identical): 1 (default) = where it pays (the first launch per threshold measures what the filter decides, one stream synchronisation), 2 = always, 0 = never */
#define CTP_TUNE_PIPE_CHUNK 2   /* queries per chunk of the ctp_sample_multiple copy/compute pipeline (0 = automatic) */
int ctp_set_tunable(ctp_map *map, int key, double value);
/* number of kernels this handle has enqueued since the last reset (host-side counter) */
int ctp_launch_counter_read(ctp_map *map, uint64_t *launches, int reset);
/* bytes the chunk pipelines (ctp_sample_multiple, ctp_query_paths) have copied host->device and
 * device->host since the last reset */
int ctp_io_counter_read(ctp_map *map, uint64_t *h2d_bytes, uint64_t *d2h_bytes, int reset);
/* stage timing of the _dev entry points, the library-side counterpart of the reference's
 * std::chrono brackets ("prm time clustering", include/topological_prm_clustering.hpp:2557-2561):
 * while enabled, every stage is bracketed by a CUDA event pair on the caller's stream.
 * ctp_profile_read waits for the recorded events and returns the summed device ms and the
 * number of spans per stage since the last reset.  Do not enable during stream capture. */
#define CTP_STAGE_REJECT 0 /* ctp_sample_reject_dev: clearance test + ordered compaction */
#define CTP_STAGE_KNN 1    /* neighbour stage of ctp_build_prm_dev */
#define CTP_STAGE_EDGES 2  /* the n*k directed segment marches (k_edge_march) */
#define CTP_STAGE_CSR 3    /* symmetric union + CSR emit */
#define CTP_STAGE_COUNT 4
int ctp_profile(ctp_map *map, int enable);
int ctp_profile_read(ctp_map *map, double ms[CTP_STAGE_COUNT], int64_t calls[CTP_STAGE_COUNT], int reset);
/* page-locked host memory for the host-pointer entry points (plain cudaHostAlloc/cudaFreeHost;
 * lets a C/C++ caller stage candidate streams and CSR outputs without linking the CUDA runtime) */
int ctp_host_alloc(void **ptr, size_t bytes);
int ctp_host_free(void *ptr);

/* ---- ESDFMap::getClearence (src/esdf_map.cpp:130-165) -> interpolateMapData (:78-121) ---- */
int ctp_clearance(ctp_map *map, const double *xyz, int64_t m, double *out);
int ctp_clearance_dev(ctp_map *map, const double *xyz, int64_t m, double *out, void *stream);
/* ---- BaseMap::isInCollision (src/base_map.cpp:7-12): out[i] = 1 if colliding ---- */
int ctp_in_collision(ctp_map *map, const double *xyz, int64_t m, double clr, uint8_t *out);
int ctp_in_collision_dev(ctp_map *map, const double *xyz, int64_t m, double clr, uint8_t *out,
                         void *stream);
/* ---- ESDFMap::gradientInVoxelCenter (src/esdf_map.cpp:167-227) ---- */
int ctp_gradient(ctp_map *map, const double *xyz, int64_t m, double *grad, double *centre);
int ctp_gradient_dev(ctp_map *map, const double *xyz, int64_t m, double *grad, double *centre,
                     void *stream);

/* ---- segment checks (src/base_map.cpp:49-74 / :77-97), one directed check per (a[i], b[i]).
 * free_out[i]=1 if free.  Optional (may be NULL): hit = first colliding sample or NaN^3,
 * hit_idx = its 1-based march index or -1, lookups = clearance evaluations the reference
 * would have made (up to and including the first hit).  group = lanes cooperating on one
 * edge (4, 8, 16, 32), or 1 = the flat lookup-parallel kernel (one lane per sample across 32
 * edges; best for short, mostly free edges such as the k-NN candidates); 0 = choose from cdc
 * and the segment lengths. */
int ctp_edge_check(ctp_map *map, const double *a, const double *b, int64_t m, double clr,
                   double cdc, int mode, int group, uint8_t *free_out, double *hit,
                   int32_t *hit_idx, int32_t *lookups);
int ctp_edge_check_dev(ctp_map *map, const double *a, const double *b, int64_t m, double clr,
                       double cdc, int mode, int group, uint8_t *free_out, double *hit,
                       int32_t *hit_idx, int32_t *lookups, void *stream);
/* same, endpoints given as indices into a node table (nodes: n_nodes x 3) */
int ctp_edge_check_indexed(ctp_map *map, const double *nodes, int64_t n_nodes,
                           const int32_t *from, const int32_t *to, int64_t m, double clr,
                           double cdc, int group, uint8_t *free_out, int32_t *hit_idx,
                           int32_t *lookups);
int ctp_edge_check_indexed_dev(ctp_map *map, const double *nodes, int64_t n_nodes,
                               const int32_t *from, const int32_t *to, int64_t m, double clr,
                               double cdc, int group, uint8_t *free_out, int32_t *hit_idx,
                               int32_t *lookups, void *stream);

/* ---- rejection sampling of TopologicalPRMClustering::sampleMultiple
 * (include/topological_prm_clustering.hpp:299-327), replayed over candidate streams.
 * q independent queries.  start/goal: q x 3.  cand: q x m x 3 (stream order = draw order of
 * fillRandomStateInEllipse).  nodes: q x n x 3 with rows 0,1 = start, goal (never tested,
 * :299-306) and rows 2.. = the first n-2 candidates with !isInCollision, in stream order.
 * n_filled[q] = rows written (n on success), n_used[q] = candidates the sequential do/while
 * would have consumed.  accept (optional): q x m bits, 1 = candidate not in collision.
 * Returns CTP_ERR_NEED_CANDIDATES (outputs still valid) if any query ran out. */
int ctp_sample_reject(ctp_map *map, const double *start, const double *goal, const double *cand,
                      int64_t q, int64_t m, double clr, int64_t n, double *nodes,
                      int32_t *n_filled, int32_t *n_used, uint8_t *accept);
int ctp_sample_reject_dev(ctp_map *map, const double *start, const double *goal,
                          const double *cand, int64_t q, int64_t m, double clr, int64_t n,
                          double *nodes, int32_t *n_filled, int32_t *n_used, uint8_t *accept,
                          void *stream);
/* candidate generation on the device: BaseMap::fillRandomStateInEllipse (src/base_map.cpp:29-43)
 * -> random_ellipsoid_point (src/common.cpp:256-312) with a counter-based generator
 * (Philox4x32-10 keyed by seed, counter = (query, candidate)); cand: q x m x 3 */
int ctp_sample_ellipsoid_dev(ctp_map *map, const double *start, const double *goal, int64_t q,
                             int64_t m, double ratio, int planar, uint64_t seed, double *cand,
                             void *stream);
int ctp_sample_ellipsoid(ctp_map *map, const double *start, const double *goal, int64_t q,
                         int64_t m, double ratio, int planar, uint64_t seed, double *cand);

/* the same candidates plus the draws behind them: draws q x m x 4 = (u, g0, g1, g2), the uniform and the three
 * normal variates random_ellipsoid_point consumes per point (src/common.cpp:295-303) */
int ctp_sample_ellipsoid_draws(ctp_map *map, const double *start, const double *goal, int64_t q, int64_t m,
                               double ratio, int planar, uint64_t seed, double *cand, double *draws);

/* ---- neighbour stage of sampleMultiple (:329-339): exact float32 k-NN.
 * nodes: q x n x 3 doubles, narrowed to float as :300-326.  idx: q x n x k (self excluded,
 * ascending (distance, index)), d2 (optional): q x n x k float32 squared distances. */
int ctp_knn(ctp_map *map, const double *nodes, int64_t q, int64_t n, int k, int32_t *idx,
            float *d2);
int ctp_knn_dev(ctp_map *map, const double *nodes, int64_t q, int64_t n, int k, int32_t *idx,
                float *d2, void *stream);

/* ---- PRM build = neighbour stage + the n*k directed checks + symmetric adjacency
 * (include/topological_prm_clustering.hpp:329-397), for q independent node sets.
 * Outputs: nbr q x n x k, directed_free q x n x k (both optional), and one CSR per query:
 * row_ptr q x (n+1) relative to that query's segment, col / w packed back to back with
 * nnz_off (q+1, int64) giving each query's segment; col ascending within a row;
 * w = ||to - from|| in FP64 from the double coordinates (:369-371).  cap = capacity of
 * col / w in entries (2*q*n*k always suffices). */
int ctp_build_prm(ctp_map *map, const double *nodes, int64_t q, int64_t n, int k, double clr,
                  double cdc, int32_t *nbr, uint8_t *directed_free, int32_t *row_ptr,
                  int32_t *col, double *w, int64_t cap, int64_t *nnz_off);
int ctp_build_prm_dev(ctp_map *map, const double *nodes, int64_t q, int64_t n, int k,
                      double clr, double cdc, int32_t *nbr, uint8_t *directed_free,
                      int32_t *row_ptr, int32_t *col, double *w, int64_t cap, int64_t *nnz_off,
                      void *stream);
/* Compact forms of the adjacency (the symmetric CSR stores every undirected edge twice, and w(i,j) = w(j,i) =
 * ||to - from|| can be recomputed from the node coordinates bit for bit): csr_flags is an OR of
 *   CTP_CSR_UPPER       only entries with col > row (each undirected edge once; row_ptr / nnz_off describe those)
 *   CTP_CSR_NO_WEIGHTS  w is neither computed nor written (w may be NULL)
 * include/ctopprm/csr_expand.hpp mirrors a compact roadmap into the full symmetric, weighted CSR on the host. */
#define CTP_CSR_UPPER 1
#define CTP_CSR_NO_WEIGHTS 2
/* wire formats of the host-pointer call ctp_sample_multiple_ex only (fewer bytes over the host link; with 8 GPUs on one
 * host the link, not the GPUs, bounds the end-to-end rate):
 *   CTP_CSR_COL16       n <= 65536: `col` receives uint16_t entries (the buffer is read as uint16_t *, cap counts entries)
 *   CTP_NODES_AS_INDEX  `nodes` receives, as int32_t q x n, the position of every accepted node in its query's candidate
 *                       stream (-2 = start, -3 = goal, -1 = row not filled) instead of coordinates the host already has */
#define CTP_CSR_COL16 4
#define CTP_NODES_AS_INDEX 8
int ctp_build_prm_ex_dev(ctp_map *map, const double *nodes, int64_t q, int64_t n, int k, double clr,
                         double cdc, int csr_flags, int32_t *nbr, uint8_t *directed_free, int32_t *row_ptr,
                         int32_t *col, double *w, int64_t cap, int64_t *nnz_off, void *stream);
/* the middle stage alone: the q*n*k directed checks (:355-373) of a given k-NN table nbr (q x n x k, -1 = no
 * neighbour) against this map -> directed_free (q x n x k).  ctp_knn_dev and ctp_csr_from_directed_dev do not look
 * at the map: a sweep over many maps runs them once for all queries and only the rejection and this stage per map. */
int ctp_edge_check_knn_dev(ctp_map *map, const double *nodes, int64_t q, int64_t n, int k, const int32_t *nbr,
                           double clr, double cdc, uint8_t *directed_free, void *stream);
/* the merge step when the n*k directed checks were sharded (several GPUs each checking a range of
 * source rows with ctp_edge_check_indexed): CSR from a given k-NN table nbr (q x n x k, -1 = no
 * neighbour) and its verdicts directed_free (q x n x k), same symmetric-union rule (:374-385). */
int ctp_csr_from_directed(ctp_map *map, const double *nodes, int64_t q, int64_t n, int k,
                          const int32_t *nbr, const uint8_t *directed_free, int32_t *row_ptr,
                          int32_t *col, double *w, int64_t cap, int64_t *nnz_off);
int ctp_csr_from_directed_dev(ctp_map *map, const double *nodes, int64_t q, int64_t n, int k,
                              const int32_t *nbr, const uint8_t *directed_free, int32_t *row_ptr,
                              int32_t *col, double *w, int64_t cap, int64_t *nnz_off, void *stream);
/* whole sampleMultiple from host candidate streams to host CSRs in one call (the
 * end-to-end path bench.py times): ctp_sample_reject + ctp_build_prm without the
 * intermediate host round trip.  nodes (q x n x 3) is also returned.  The queries are cut
 * into chunks whose host->device copies, kernels and device->host copies overlap on three
 * streams; pass page-locked buffers (ctp_host_alloc) to get the overlap.  When a candidate
 * stream runs out, the other queries' outputs are still complete and the call returns
 * CTP_ERR_NEED_CANDIDATES (n_filled tells which). */
int ctp_sample_multiple(ctp_map *map, const double *start, const double *goal,
                        const double *cand, int64_t q, int64_t m, int64_t n, int k, double clr,
                        double cdc, double *nodes, int32_t *n_filled, int32_t *row_ptr,
                        int32_t *col, double *w, int64_t cap, int64_t *nnz_off);

/* same with a compact adjacency (csr_flags as ctp_build_prm_ex_dev): two thirds to four fifths fewer bytes come back */
int ctp_sample_multiple_ex(ctp_map *map, const double *start, const double *goal, const double *cand, int64_t q,
                           int64_t m, int64_t n, int k, double clr, double cdc, int csr_flags, double *nodes,
                           int32_t *n_filled, int32_t *row_ptr, int32_t *col, double *w, int64_t cap,
                           int64_t *nnz_off);

/* One GPU's share of ONE large query split over several GPUs (the callers' side is
 * ctopprm_learn_b200/sharded.py: candidate blocks -> accept counts -> all-gather of the accepted nodes -> this call on
 * every GPU -> element-wise maximum of nbr and directed_free over the GPUs -> ctp_csr_from_directed_dev).
 * Neighbour search and directed checks (:329-397) for the nodes at positions [first, last) of the k-NN grid's cell
 * order, the bounds moved up to cell boundaries (all GPUs build the same grid from the same nodes and so agree on the
 * owner of every node).  nbr (n x k) / directed_free (n x k) rows of the slice are written; all other rows are set to
 * 0x80808080 / 0.  rows_done (host, optional): nodes in the slice after alignment.  Synchronises the stream once
 * (16-byte read back of the aligned bounds). */
int ctp_prm_rows_dev(ctp_map *map, const double *nodes, int64_t n, int k, double clr, double cdc, int64_t first,
                     int64_t last, int32_t *nbr, uint8_t *directed_free, int64_t *rows_done, void *stream);

/* ---- graph search over the CSR roadmaps (first "next" item): shortest paths from
 * a set of sources, one independent problem per query.  Replaces Dijkstra::findPath as called by
 * findShortestPath (include/topological_prm_clustering.hpp:2172, include/dijkstra.hpp:83-176) with
 * src = {0}, and the multi-source flood of wavefrontFill (:789-962) with src = the cluster seeds.
 * CSR as produced by ctp_build_prm (row_ptr q x (n+1) relative, col / w packed, nnz_off q+1).
 * src: q x ns node ids (-1 = unused slot).  dist: q x n (+inf = unreachable), the FP64 sum of the
 * edge weights along a shortest path in path order, as the reference accumulates it.  pred: q x n,
 * the smallest strictly nearer neighbour u with dist[u] + w(u,v) == dist[v]; -1 for sources and
 * unreachable nodes.  label (optional): q x n, index into the query's src list of the source at the
 * root of the predecessor chain (the cluster the flood assigns the node to), -1 if none. */
int ctp_sssp(ctp_map *map, int64_t q, int64_t n, const int32_t *row_ptr, const int32_t *col,
             const double *w, const int64_t *nnz_off, const int32_t *src, int ns, double *dist,
             int32_t *pred, int32_t *label);
int ctp_sssp_dev(ctp_map *map, int64_t q, int64_t n, const int32_t *row_ptr, const int32_t *col,
                 const double *w, const int64_t *nnz_off, const int32_t *src, int ns, double *dist,
                 int32_t *pred, int32_t *label, void *stream);

/* ---- the floods of the cluster stage (wavefrontFill + addCentroid on the GPU CSR).
 * One call = one flood for q roadmaps at once.
 *   mode 0: wavefrontFill's multi-source Dijkstra from seeds[0 .. n_seeds) with its inter-cluster connection
 *           records (include/topological_prm_clustering.hpp:826-962);
 *   mode 1: addCentroid's re-flood from the newest seed seeds[n_seeds-1] on top of the dist / prev / label state of
 *           the floods before it (:1276-1374) -- only the nodes the new cluster claims change.
 * seeds: q x max_seeds node ids, n_seeds[q] (0 = this query takes no part: nothing of it is touched).
 * dist / prev / label: q x n, written (mode 0) or updated (mode 1); dist = +inf, prev = label = -1 for nodes no seed
 * reaches.  prev[v] = the node whose relaxation first produced dist[v] in the sequential loop's pop order,
 * ascending (distance, node id).  Records: every (expanding node ex, neighbour nb) the sequential loop would have
 * appended to all_cluster_connections / new_connections, with its distance dist_then(nb) + dist(ex) + w where
 * dist_then is nb's possibly still tentative distance at that moment; rec_nodes q x rec_cap x 2 (ex, nb), rec_label
 * q x rec_cap (the cluster nb belonged to at that moment -- it may change later in the same flood), rec_dist
 * q x rec_cap, rec_count[q] (records beyond rec_cap are counted, not written).  The _dev form leaves them in no
 * particular order; ctp_cluster_flood sorts each query's records into the sequential loop's append order.
 * n_updates[q]: successful relaxations (the reference's cnt - 1, which caps addCentroid at n, :1319: a mode-1 flood
 * whose count reaches n - 1 is not reproduced -- callers test for it).
 * The neighbour order (ascending id) and the pop order among equal distances (ascending id) are this library's rule:
 * the reference iterates an unordered_map keyed by pointers, so it has no reproducible order of its own. */
int ctp_cluster_flood_dev(ctp_map *map, int64_t q, int64_t n, const int32_t *row_ptr, const int32_t *col,
                          const double *w, const int64_t *nnz_off, const int32_t *seeds, int max_seeds,
                          const int32_t *n_seeds, int mode, double *dist, int32_t *prev, int32_t *label,
                          int64_t rec_cap, int32_t *rec_nodes, int32_t *rec_label, double *rec_dist,
                          int32_t *rec_count, uint64_t *n_updates, void *stream);

/* q roadmaps kept on the device together with the state of their floods -- what clusterGraph (:1832-1952) keeps in
 * its HeapNodes (distance_from_start, previous_point, cluster_id) between wavefrontFill and the addCentroid calls.
 * CSR as ctp_build_prm returns it.  info = { q, n, nnz, record capacity per query (= its widest nnz) }. */
typedef struct ctp_roadmap ctp_roadmap;
int ctp_roadmap_create(ctp_map *map, int64_t q, int64_t n, const int32_t *row_ptr, const int32_t *col,
                       const double *w, const int64_t *nnz_off, int max_seeds, ctp_roadmap **out);
int ctp_roadmap_destroy(ctp_roadmap *r);
int ctp_roadmap_info(const ctp_roadmap *r, int64_t info[4]);
/* host form of ctp_cluster_flood_dev on a resident roadmap: seeds (q x max_seeds) and n_seeds in; the state after the
 * flood and the sorted records out (rec_nodes q x cap x 2, rec_label / rec_dist q x cap with cap from ctp_roadmap_info) */
int ctp_cluster_flood(ctp_roadmap *r, const int32_t *seeds, const int32_t *n_seeds, int mode, double *dist,
                      int32_t *prev, int32_t *label, int32_t *rec_nodes, int32_t *rec_label, double *rec_dist,
                      int32_t *rec_count, uint64_t *n_updates);

/* The planner's first two phases in one call (find_geometrical_paths :2558-2572: sampleMultiple, then
 * findShortestPath), for q independent queries: host candidate streams in, host start->goal paths
 * out; the roadmaps stay on the device.  path_ids: q x path_cap node ids in the query's own node
 * numbering (0 = start, 1 = goal, 2.. = accepted candidates in stream order), -1 padded; path_xyz:
 * q x path_cap x 3 coordinates; path_n[q] = nodes on the path (0 = goal unreachable or path longer
 * than path_cap); length[q] = its length (+inf if unreachable); n_filled as ctp_sample_multiple. */
int ctp_query_paths(ctp_map *map, const double *start, const double *goal, const double *cand,
                    int64_t q, int64_t m, int64_t n, int k, double clr, double cdc, int32_t path_cap,
                    int32_t *path_ids, double *path_xyz, int32_t *path_n, double *length,
                    int32_t *n_filled);

/* ---- polylines.  A batch of polylines is a point table pts (total x 3) plus offsets
 * off (count+1, int32): polyline i = pts[off[i] .. off[i+1]) ---- */

/* BaseMap::samplePath (src/base_map.cpp:105-157): out_off (count+1) gives where each
 * resampled polyline (int)num_samples[i] points) starts in out. status[i] = 0 or
 * CTP_ERR_SAMPLE_PATH. */
int ctp_sample_path(ctp_map *map, const double *pts, const int32_t *off, const double *length,
                    const double *num_samples, int64_t count, double *out,
                    const int64_t *out_off, int32_t *status);
/* BaseMap::isDeformablePath (src/base_map.cpp:160-192) for `pairs` pairs (ia[i], ib[i]) of
 * polylines; num_check[i] as the caller computes it (ceil(max len / cdc) + 1).
 * out[i] = 1 deformable, 0 not, CTP_ERR_SAMPLE_PATH (as uint8 0xFB) on overrun. */
int ctp_deformable(ctp_map *map, const double *pts, const int32_t *off, const double *length,
                   int64_t count, const int32_t *ia, const int32_t *ib, const double *num_check,
                   int64_t pairs, double clr, double cdc, uint8_t *out);
/* TopologicalPRMClustering::shorten_path (include/topological_prm_clustering.hpp:2313-2479)
 * for `count` polylines at once.  out: count x cap x 3 points, origin: count x cap (index of
 * the reused input node, -1 for a new push-off node), out_n / out_len / status per polyline
 * (status 0 ok, 1 = "no point added", original returned; <0 = CTP_ERR_*). */
int ctp_shorten_path(ctp_map *map, const double *pts, const int32_t *off, const double *length,
                     int64_t count, int forward, double clr, double cdc, int32_t cap,
                     double *out, int32_t *origin, int32_t *out_n, double *out_len,
                     int32_t *status);
/* the same with a direction per polyline (forward_each[i] != 0 = forward; NULL = `forward` for all): isNewHomotopyClass
 * (:1702-1707) shortens one path backward-then-forward and the other forward-then-backward -- both go in one launch */
int ctp_shorten_path_ex(ctp_map *map, const double *pts, const int32_t *off, const double *length,
                        int64_t count, int forward, const uint8_t *forward_each, double clr, double cdc, int32_t cap,
                        double *out, int32_t *origin, int32_t *out_n, double *out_len, int32_t *status);

#ifdef __cplusplus
}
#endif
#endif
