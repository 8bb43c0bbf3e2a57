// host_capi.cpp -- flat C entry points over the C++ class surface (include/ctopprm/*.hpp) so that
// the parity tests can drive ESDFMap / BaseMap / TopologicalPRMClustering exactly as a C++ caller
// of the reference would, from Python.  Not part of the drop-in boundary itself.
#include <algorithm>
#include <cstring>
#include <memory>

#include "ctopprm/csr_expand.hpp"
#include "ctopprm/esdf_map.hpp"
#include "ctopprm/topological_prm_clustering.hpp"

typedef Vector<3> V3;
typedef HeapNode<V3> Node;
typedef path_with_length<V3> Path;
typedef TopologicalPRMClustering<V3> Planner;

namespace {
V3 V(const double *p) { return V3(p[0], p[1], p[2]); }
void put(const V3 &v, double *o) { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
struct Poly {
  std::vector<std::unique_ptr<Node>> store;
  std::vector<Node *> ptrs;
  Poly(const double *pts, int n) {
    for (int i = 0; i < n; i++) {
      store.emplace_back(new Node(V(pts + 3 * i)));
      store.back()->id = i;
      ptrs.push_back(store.back().get());
    }
  }
};
struct MapBox { std::shared_ptr<ESDFMap> map; };
struct PlannerBox { std::unique_ptr<Planner> pl; };
}  // namespace

extern "C" {
// compact (upper, unweighted) adjacency of q roadmaps -> full symmetric weighted CSR, on `threads` host threads
void ctph_csr_expand_upper(const double *nodes, long long q, long long n, const int *rp_up, const int *col_up,
                           const long long *off_up, int *row_ptr, int *col, double *w, long long *nnz_off, int threads) {
  ctopprm::csr_expand_upper(nodes, q, n, rp_up, col_up, (const int64_t *)off_up, row_ptr, col, w, (int64_t *)nnz_off, threads);
}
void *ctph_map_load(const char *filename, int device, int layouts) {
  try {
    ESDFMap::setDevice(device);
    if (layouts) ESDFMap::setLayouts(layouts);
    auto *b = new MapBox{std::make_shared<ESDFMap>()};
    b->map->load(filename);
    return b;
  } catch (const std::exception &e) {
    fprintf(stderr, "ctph_map_load: %s\n", e.what());
    return nullptr;
  }
}
void *ctph_map_from_memory(const float *vox, int n, const double *c, const double *e, int device, int layouts) {
  ESDFMap::setDevice(device);
  if (layouts) ESDFMap::setLayouts(layouts);
  auto *b = new MapBox{std::make_shared<ESDFMap>()};
  b->map->loadFromMemory(vox, n, V(c), V(e));
  return b;
}
void *ctph_map_from_occupancy(const unsigned char *occ, int n, const double *c, const double *e, int device, int layouts) {
  ESDFMap::setDevice(device);
  if (layouts) ESDFMap::setLayouts(layouts);
  auto *b = new MapBox{std::make_shared<ESDFMap>()};
  b->map->loadFromOccupancy(occ, n, V(c), V(e));
  return b;
}
int ctph_map_save(void *h, const char *filename) {
  try {
    static_cast<MapBox *>(h)->map->save(filename);
    return 0;
  } catch (const std::exception &e) {
    fprintf(stderr, "ctph_map_save: %s\n", e.what());
    return -1;
  }
}
void ctph_map_free(void *h) { delete static_cast<MapBox *>(h); }
#define MAP(h) (static_cast<MapBox *>(h)->map)
double ctph_get_clearence(void *h, const double *p) { return MAP(h)->getClearence(V(p)); }
void ctph_min_max(void *h, double *out) {
  put(MAP(h)->getMinPos(), out);
  put(MAP(h)->getMaxPos(), out + 3);
}
void ctph_gradient(void *h, const double *p, double *g, double *c) {
  auto r = MAP(h)->gradientInVoxelCenter(V(p));
  put(r.first, g);
  put(r.second, c);
}
void ctph_real_pos_from_index(void *h, const double *q, double *out) { put(MAP(h)->getRealPosFromIndex(V(q)), out); }
int ctph_in_collision(void *h, const double *p, double clr) { return MAP(h)->isInCollision(V(p), clr); }
int ctph_edge_nodes(void *h, const double *a, const double *b, double clr, double cdc, double *hit) {
  auto r = MAP(h)->isSimplePathFreeBetweenNodes(V(a), V(b), clr, cdc);
  put(r.second, hit);
  return r.first;
}
int ctph_edge_guards(void *h, const double *a, const double *b, double clr, double cdc) {
  return MAP(h)->isSimplePathFreeBetweenGuards(V(a), V(b), clr, cdc);
}
int ctph_path_collision_free(void *h, const double *pts, int npts, double clr, double cdc, double *hit) {
  Poly p(pts, npts);
  Path pl;
  pl.plan = p.ptrs;
  auto r = MAP(h)->isPathCollisionFree(pl, clr, cdc);
  put(r.second, hit);
  return r.first;
}
int ctph_deformable(void *h, const double *p1, int n1, double l1, const double *p2, int n2, double l2, double num, double clr,
                    double cdc) {
  Poly a(p1, n1), b(p2, n2);
  return MAP(h)->isDeformablePath(a.ptrs, l1, b.ptrs, l2, num, clr, cdc);
}
int ctph_deformable_between(void *h, const double *s, const double *e, const double *b1, const double *b2, double clr,
                            double cdc) {
  Node ns(V(s)), ne(V(e)), n1(V(b1)), n2(V(b2));
  return MAP(h)->isDeformablePathBetween(&ns, &ne, &n1, &n2, clr, cdc);
}
void ctph_sample_path(void *h, const double *pts, int npts, double len, double num, double *out) {
  Poly p(pts, npts);
  auto s = MAP(h)->samplePath(p.ptrs, len, num);
  for (size_t i = 0; i < s.size(); i++) put(s[i], out + 3 * i);
}
void ctph_fill_random_state_in_ellipse(void *h, const double *s, const double *g, double ratio, int planar, uint64_t seed,
                                       int count, double *out) {
  Node ns(V(s)), ng(V(g)), nn;
  MAP(h)->setRandomSeed(seed);
  for (int i = 0; i < count; i++) {
    MAP(h)->fillRandomStateInEllipse(&nn, &ns, &ng, ratio, planar != 0);
    put(nn.data, out + 3 * i);
  }
}

// ---- planner ----
void *ctph_planner_new(void *map, const char *yaml_text, const double *start, const double *goal) {
  try {
    YAML::Node cfg = YAML::Load(std::string(yaml_text));
    auto *b = new PlannerBox{std::unique_ptr<Planner>(new Planner(cfg, MAP(map), "./", V(start), V(goal)))};
    b->pl->setBorders(MAP(map)->getMinPos(), MAP(map)->getMaxPos());
    return b;
  } catch (const std::exception &e) {
    fprintf(stderr, "ctph_planner_new: %s\n", e.what());
    return nullptr;
  }
}
void ctph_planner_free(void *p) { delete static_cast<PlannerBox *>(p); }
#define PL(p) (static_cast<PlannerBox *>(p)->pl)
void ctph_planner_set_candidates(void *p, const double *cand, int64_t m) {
  std::vector<V3> c((size_t)m);
  if (m) memcpy(c[0].data(), cand, (size_t)m * 24);
  PL(p)->setCandidateStream(c);
}
void ctph_planner_sample_multiple(void *p, int n) { PL(p)->sampleMultiple(n); }
// flatten the HeapNode graph: nodes n x 3, CSR with columns ascending; returns nnz (or -1 if cap is short)
int64_t ctph_planner_graph(void *p, double *nodes, int32_t *row_ptr, int32_t *col, double *w, int64_t cap) {
  const auto &ns = PL(p)->nodes();
  int64_t nnz = 0;
  row_ptr[0] = 0;
  for (size_t i = 0; i < ns.size(); i++) {
    put(ns[i]->data, nodes + 3 * i);
    std::vector<std::pair<int, double>> row;
    for (auto &nb : ns[i]->visibility_node_ids) row.push_back({nb.first->id, nb.second});
    std::sort(row.begin(), row.end());
    for (auto &r : row) {
      if (nnz >= cap) return -1;
      col[nnz] = r.first;
      w[nnz] = r.second;
      nnz++;
    }
    row_ptr[i + 1] = (int32_t)nnz;
  }
  return nnz;
}
int ctph_planner_shortest(void *p, int32_t *ids, int cap, double *length) {
  auto r = PL(p)->findShortestPath();
  if (r.empty()) return 0;
  *length = r[0].length;
  int c = 0;
  for (auto *n : r[0].plan)
    if (c < cap) ids[c++] = n->id;
  return (int)r[0].plan.size();
}
void ctph_planner_save_roadmaps(void *p, const char *dense_file, const char *csr_file) {
  PL(p)->saveRoadmapDense(dense_file);
  PL(p)->saveRoadmapFromCSR(csr_file);
}
int ctph_planner_flood(void *p, const int32_t *seeds, int ns, double *dist, int32_t *pred, int32_t *cluster) {
  std::vector<int> sd(seeds, seeds + ns), pr, cl;
  std::vector<double> d;
  PL(p)->floodFromSeeds(sd, d, pr, cl);
  for (size_t i = 0; i < d.size(); i++) {
    dist[i] = d[i];
    pred[i] = pr[i];
    cluster[i] = cl[i];
  }
  return (int)d.size();
}
// isNewHomotopyClassBatch on P (min, max) pairs given as packed polylines (2P polylines: mins then maxes)
int ctph_new_homotopy_batch(const double *pts, const int32_t *off, const double *lens, int P, uint8_t *out) {
  std::vector<std::unique_ptr<Poly>> polys;
  std::vector<Path> mn, mx;
  for (int i = 0; i < 2 * P; i++) {
    polys.emplace_back(new Poly(pts + 3 * off[i], off[i + 1] - off[i]));
    Path p;
    p.plan = polys.back()->ptrs;
    p.length = lens[i];
    (i < P ? mn : mx).push_back(p);
  }
  auto r = Planner::isNewHomotopyClassBatch(mn, mx);
  for (int i = 0; i < P; i++) out[i] = r[i];
  return P;
}
int ctph_planner_seed_accepted(void *p, int candidate_id, const int32_t *seed_ids, int ns) {
  const auto &ns_ = PL(p)->nodes();
  std::vector<Node *> seeds;
  for (int i = 0; i < ns; i++) seeds.push_back(ns_[seed_ids[i]]);
  return PL(p)->isNewSeedAccepted(ns_[candidate_id], seeds);
}
// static shorten_path on a polyline given as points; needs a planner to have been constructed (statics)
int ctph_shorten_path(const double *pts, int npts, double length, int forward, double *out, int cap, double *out_len) {
  Poly poly(pts, npts);
  Path in;
  in.plan = poly.ptrs;
  in.length = length;
  Path s = Planner::shorten_path(in, forward != 0);
  *out_len = s.length;
  for (size_t i = 0; i < s.plan.size() && (int)i < cap; i++) put(s.plan[i]->data, out + 3 * i);
  return (int)s.plan.size();
}
int ctph_remove_equivalent(const double *pts, const int32_t *off, const double *lens, int P, int32_t *keep_first_id) {
  std::vector<std::unique_ptr<Poly>> polys;
  std::vector<Path> paths;
  for (int i = 0; i < P; i++) {
    polys.emplace_back(new Poly(pts + 3 * off[i], off[i + 1] - off[i]));
    Path p;
    p.plan = polys.back()->ptrs;
    p.length = lens[i];
    p.from_id = i;  // tag: which input path this is
    paths.push_back(p);
  }
  auto out = Planner::removeEquivalentPaths(paths);
  for (size_t i = 0; i < out.size(); i++) keep_first_id[i] = out[i].from_id;
  return (int)out.size();
}
int ctph_remove_too_long(const double *lens, int P, int32_t *keep) {
  std::vector<Path> paths((size_t)P);
  for (int i = 0; i < P; i++) {
    paths[i].length = lens[i];
    paths[i].from_id = i;
  }
  auto out = Planner::removeTooLongPaths(paths);
  for (size_t i = 0; i < out.size(); i++) keep[i] = out[i].from_id;
  return (int)out.size();
}
// clusterGraph on a planner whose roadmap exists: stitched cluster paths (points back to back, offsets, DFS lengths), the
// final seeds and labels, stats = {seeds, floods, refloods, capped_refloods, homotopy_checks}.  Returns the number of
// paths, -1 when a capacity is too small.
int ctph_planner_cluster_graph(void *p, double max_path_length, const int32_t *seed_candidates, int n_cand, int cap_paths,
                               int cap_pts, double *out_pts, int32_t *out_off, double *out_len, int32_t *seeds_out,
                               int32_t *nseeds_out, int32_t *labels_out, int32_t *stats_out) {
  PL(p)->setMaxPathLength(max_path_length);
  PL(p)->setSeedCandidates(std::vector<int>(seed_candidates, seed_candidates + n_cand));
  std::vector<Path> r = PL(p)->clusterGraph();
  int tot = 0;
  out_off[0] = 0;
  for (size_t i = 0; i < r.size(); i++) {
    if ((int)i >= cap_paths || tot + (int)r[i].plan.size() > cap_pts) return -1;
    for (auto *nd : r[i].plan) put(nd->data, out_pts + 3 * (size_t)tot++);
    out_off[i + 1] = tot;
    out_len[i] = r[i].length;
  }
  const auto &sd = PL(p)->clusterSeeds();
  for (size_t i = 0; i < sd.size(); i++) seeds_out[i] = sd[i];
  *nseeds_out = (int32_t)sd.size();
  const auto &lb = PL(p)->clusterLabels();
  for (size_t i = 0; i < lb.size(); i++) labels_out[i] = lb[i];
  const auto &st = PL(p)->clusterStats();
  const int32_t sv[5] = {st.seeds, st.floods, st.refloods, st.capped_refloods, st.homotopy_checks};
  memcpy(stats_out, sv, sizeof sv);
  return (int)r.size();
}
// whole entry point for the single start-goal pair; cand (optional, m x 3): the candidate stream of sampleMultiple.
// Paths come back sorted by length: points back to back (out_pts, cap_pts points), offsets (cap_paths + 1), lengths.
// Returns the number of paths (-1: capacity).
int ctph_find_geometrical_paths(void *map, const char *yaml_text, const double *start, const double *goal,
                                const char *output_folder, const double *cand, int64_t m, int cap_paths, int cap_pts,
                                double *out_pts, int32_t *out_off, double *lengths) {
  YAML::Node cfg = YAML::Load(std::string(yaml_text));
  std::vector<V3> gates = {V(start), V(goal)};
  Planner::gate_candidate_streams.clear();
  if (cand && m > 0) {
    std::vector<V3> c((size_t)m);
    memcpy(c[0].data(), cand, (size_t)m * 24);
    Planner::gate_candidate_streams.push_back(std::move(c));
  }
  auto r = Planner::find_geometrical_paths(cfg, MAP(map), gates, output_folder);
  Planner::gate_candidate_streams.clear();
  if (r.empty()) return 0;
  int tot = 0;
  out_off[0] = 0;
  for (size_t i = 0; i < r[0].size(); i++) {
    if ((int)i >= cap_paths || tot + (int)r[0][i].plan.size() > cap_pts) return -1;
    for (auto *nd : r[0][i].plan) put(nd->data, out_pts + 3 * (size_t)tot++);
    out_off[i + 1] = tot;
    lengths[i] = r[0][i].length;
  }
  return (int)r[0].size();
}
}

// phases of the last find_geometrical_paths call: {prm, dijkstra, wavefront, min_tours, dfs_and_stitch, shorten_and_filter,
// total} in ms, then {cluster_paths, seeds, refloods, homotopy_checks, capped_refloods}
extern "C" void ctph_last_phase_times(double *ms7, int32_t *counts5) {
  const auto &p = Planner::last_phase_times;
  const double t[7] = {p.prm, p.dijkstra, p.wavefront, p.min_tours, p.dfs_and_stitch, p.shorten_and_filter, p.total};
  const int32_t c[5] = {p.cluster_paths, p.seeds, p.refloods, p.homotopy_checks, p.capped_refloods};
  memcpy(ms7, t, sizeof t);
  memcpy(counts5, c, sizeof c);
}

// YAML subset reader, testable without a device: value of "a" or "a.b" as text; returns 0 if undefined
extern "C" int ctph_yaml_get(const char *yaml_text, const char *dotted_key, char *out, int cap) {
  YAML::Node cfg = YAML::Load(std::string(yaml_text));
  std::string key(dotted_key);
  const YAML::Node *n = &cfg;
  size_t pos = 0;
  while (true) {
    const size_t dot = key.find('.', pos);
    n = &(*n)[key.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos)];
    if (dot == std::string::npos) break;
    pos = dot + 1;
  }
  if (!n->IsDefined()) return 0;
  snprintf(out, (size_t)cap, "%s", n->Scalar().c_str());
  return 1;
}
