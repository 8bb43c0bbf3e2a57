// topological_prm_clustering.cpp -- TopologicalPRMClustering<Vector<3>> on the C ABI.
#include <cstring>

#include "ctopprm/topological_prm_clustering.hpp"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <fstream>

typedef Vector<3> V3;
typedef HeapNode<V3> Node;
typedef path_with_length<V3> Path;
typedef TopologicalPRMClustering<V3> Planner;

template <> std::shared_ptr<BaseMap> Planner::map_ = nullptr;
template <> std::string Planner::output_folder_ = "./";
template <> double Planner::collision_distance_check_ = 0;
template <> double Planner::min_clearance_ = 0;
template <> double Planner::cutoof_distance_ratio_to_shortest_ = 0;
template <> double Planner::min_allowed = 0;
template <> uint64_t Planner::sampling_seed = 1;
template <> std::vector<std::vector<V3>> Planner::gate_candidate_streams = {};
template <> Planner::PhaseTimes Planner::last_phase_times = {};
template <>
std::function<std::vector<Path>(Planner &, const Path &)> Planner::cluster_stage =
    [](Planner &prm, const Path &) { return prm.clusterGraph(); };

namespace {
[[noreturn]] void die(const char *what) {
  ERROR(what << ": " << ctp_last_error());
  exit(1);
}
ctp_map *handleOf(const std::shared_ptr<BaseMap> &m, const char *who) {
  if (!m || !m->deviceMap()) throw std::runtime_error(std::string(who) + ": the map is not resident on a device");
  return m->deviceMap();
}
void packPaths(const std::vector<Path> &paths, std::vector<double> &pts, std::vector<int32_t> &off, std::vector<double> &len) {
  off.assign(1, 0);
  pts.clear();
  len.clear();
  for (auto &p : paths) {
    for (auto *n : p.plan) pts.insert(pts.end(), n->data.data(), n->data.data() + 3);
    off.push_back((int32_t)(pts.size() / 3));
    len.push_back(p.length);
  }
}
}  // namespace

template <>
Planner::TopologicalPRMClustering(const YAML::Node &planner_config, std::shared_ptr<BaseMap> map, std::string output_folder,
                                  V3 start, V3 end, double, double) {
  map_ = map;
  output_folder_ = output_folder;
  start_ = new Node(start);
  start_->city_node = true;
  start_->id = 0;
  end_ = new Node(end);
  end_->city_node = true;
  end_->id = 1;
  // the eleven keys of the reference constructor (:257-269)
  num_clusters_ = loadParam<int>(planner_config, "num_clusters");
  max_clusters_ = loadParam<int>(planner_config, "max_clusters");
  min_clusters_ = loadParam<int>(planner_config, "min_clusters");
  min_ratio_ = loadParam<double>(planner_config, "min_ratio");
  planar_ = loadParam<bool>(planner_config, "planar");
  max_path_length_ratio_ = loadParam<double>(planner_config, "max_path_length_ratio");
  cutoof_distance_ratio_to_shortest_ = loadParam<double>(planner_config, "cutoof_distance_ratio_to_shortest");
  min_clearance_ = loadParam<double>(planner_config, "min_clearance");
  collision_distance_check_ = loadParam<double>(planner_config, "collision_distance_check");
  ellipse_ratio_major_axis_focal_length_ = loadParam<double>(planner_config, "ellipse_ratio_major_axis_focal_length");
  // addition of the accelerated build (not a reference key): candidates limited to a radius around each node
  if (planner_config["neighbour_radius"].IsDefined()) neighbour_radius_ = planner_config["neighbour_radius"].as<double>();
  if (collision_distance_check_ == 0) {
    ERROR("you need to specify collision_distance_check for sampling-based motion planning");
    exit(1);
  }
}

template <> void Planner::setBorders(V3 min_position, V3 max_position) {
  min_position_ = min_position;
  max_position_ = max_position;
  position_range_ = max_position - min_position;
}

// PRM construction (:288-401): candidates -> rejection -> k-NN(13) -> directed checks -> symmetric
// adjacency, one ctp_sample_multiple call; the CSR is re-materialised into visibility_node_ids.
namespace {
// Page-locked buffers for the one call that moves megabytes per query (candidates in, roadmap out): through pageable
// std::vectors the copies are staged by the driver at a few GB/s and took most of sampleMultiple's 0.9 ms.
struct PlannerStaging {
  void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap[5] = {0, 0, 0, 0, 0};
  void *get(int i, size_t bytes) {
    if (bytes > cap[i]) {
      if (p[i]) ctp_host_free(p[i]);
      p[i] = nullptr;
      cap[i] = 0;
      const size_t want = bytes + bytes / 4 + 4096;
      if (ctp_host_alloc(&p[i], want) != CTP_OK) die("ctp_host_alloc");
      cap[i] = want;
    }
    return p[i];
  }
  ~PlannerStaging() {
    for (void *q : p)
      if (q) ctp_host_free(q);
  }
};
}  // namespace

template <> void Planner::sampleMultiple(const int num_samples) {
  ctp_map *h = handleOf(map_, "sampleMultiple");
  const int n = num_samples, k = num_nn - 1;
  if (n < 2) return;
  // one set per host thread, kept for the life of the thread: planners come and go per query, and page-locking a few
  // megabytes costs milliseconds
  static thread_local PlannerStaging stg;
  const size_t ecap = (size_t)2 * n * k;
  double *nodes = (double *)stg.get(0, (size_t)n * 24), *w = (double *)stg.get(1, ecap * 8);
  int32_t *row_ptr = (int32_t *)stg.get(2, ((size_t)n + 1) * 4), *col = (int32_t *)stg.get(3, ecap * 4);
  int32_t n_filled = 0;
  int64_t nnz_off[2] = {0, 0};
  if (ctp_set_tunable(h, CTP_TUNE_KNN_RADIUS, neighbour_radius_ > 0 ? neighbour_radius_ : 0.0) != CTP_OK) die("ctp_set_tunable");
  std::vector<V3> drawn;
  const bool generated = candidate_stream_.empty();
  int64_t m = generated ? 4 * (int64_t)n : (int64_t)candidate_stream_.size();
  for (int attempt = 0;; attempt++) {
    if (generated) {
      drawn.resize((size_t)m);
      if (ctp_sample_ellipsoid(h, start_->data.data(), end_->data.data(), 1, m, ellipse_ratio_major_axis_focal_length_,
                               planar_ ? 1 : 0, sampling_seed, drawn[0].data()) != CTP_OK)
        die("ctp_sample_ellipsoid");
    }
    const std::vector<V3> &cand = generated ? drawn : candidate_stream_;
    double *cand_pin = (double *)stg.get(4, (size_t)m * 24);
    memcpy(cand_pin, cand[0].data(), (size_t)m * 24);
    const int rc = ctp_sample_multiple(h, start_->data.data(), end_->data.data(), cand_pin, 1, m, n, k, min_clearance_,
                                       collision_distance_check_, nodes, &n_filled, row_ptr, col, w, (int64_t)ecap, nnz_off);
    if (rc == CTP_OK) break;
    if (rc == CTP_ERR_NEED_CANDIDATES && generated && attempt < 6) {
      m *= 4;  // the reference's do/while would simply keep drawing (:314-317)
      continue;
    }
    die("ctp_sample_multiple");
  }
  nodes_.assign((size_t)n, nullptr);
  nodes_[0] = start_;
  nodes_[1] = end_;
  node_xyz_.assign(nodes, nodes + (size_t)n * 3);
  sp_dist_.clear();
  sp_pred_.clear();
  cluster_labels_.clear();
  csr_row_ptr_.assign(row_ptr, row_ptr + n + 1);
  csr_col_.assign(col, col + nnz_off[1]);
  csr_w_.assign(w, w + nnz_off[1]);
  adjacency_ready_ = false;  // HeapNodes and their visibility_node_ids are created on first use (nodes(), nodeAt())
}

template <> Node *Planner::nodeAt(int i) {
  Node *&nd = nodes_[(size_t)i];
  if (!nd) {
    nd = new Node(V3(node_xyz_[3 * (size_t)i], node_xyz_[3 * (size_t)i + 1], node_xyz_[3 * (size_t)i + 2]));
    nd->id = i;
  }
  return nd;
}

// every HeapNode of the roadmap, with the fields the stages that ran so far would have left in them
template <> void Planner::materializeNodes() {
  const int n = (int)nodes_.size();
  for (int i = 0; i < n; i++) nodeAt(i);
  if ((int)sp_dist_.size() == n) {
    for (int i = 0; i < n; i++) {
      nodes_[i]->distance_from_start = std::isfinite(sp_dist_[i]) ? sp_dist_[i] : DBL_MAX;
      nodes_[i]->previous_point = sp_pred_[i] >= 0 ? nodes_[sp_pred_[i]] : nullptr;
    }
    nodes_[0]->previous_point = nodes_[0];
  }
  if ((int)cluster_labels_.size() == n)
    for (int i = 0; i < n; i++) nodes_[i]->cluster_id = cluster_labels_[i];
}

// CSR -> visibility_node_ids of every node (the pointer graph of the reference, :374-385)
template <> void Planner::materializeAdjacency() {
  materializeNodes();
  if (adjacency_ready_) return;
  adjacency_ready_ = true;
  const int n = (int)nodes_.size();
  for (int i = 0; i < n; i++) {
    auto &adj = nodes_[i]->visibility_node_ids;
    adj.reserve(adj.size() + (size_t)(csr_row_ptr_[i + 1] - csr_row_ptr_[i]));  // one bucket array per node, no rehash
    for (int32_t e = csr_row_ptr_[i]; e < csr_row_ptr_[i + 1]; e++) adj[nodes_[csr_col_[e]]] = csr_w_[e];
  }
}

// findShortestPath (:2172) = dijkstra.findPath(0, {1}, nodes_): on the device over the CSR the PRM build
// returned (ctp_sssp); distance_from_start / previous_point of every node are filled in as the
// reference's Dijkstra leaves them for the nodes it settles.
template <> std::vector<Path> Planner::findShortestPath() {
  const int n = (int)nodes_.size();
  if (n < 2 || (int)csr_row_ptr_.size() != n + 1) {
    materializeAdjacency();
    return dijkstra.findPath(0, {1}, nodes_);
  }
  ctp_map *h = handleOf(map_, "findShortestPath");
  sp_dist_.resize((size_t)n);
  sp_pred_.resize((size_t)n);
  const int64_t nnz_off[2] = {0, (int64_t)csr_col_.size()};
  const int32_t src = 0;
  if (ctp_sssp(h, 1, n, csr_row_ptr_.data(), csr_col_.data(), csr_w_.data(), nnz_off, &src, 1, sp_dist_.data(), sp_pred_.data(),
               nullptr) != CTP_OK)
    die("ctp_sssp");
  // the HeapNodes that exist already get their fields now, the others when they are created (materializeNodes)
  for (int i = 0; i < n; i++)
    if (nodes_[i]) {
      nodes_[i]->distance_from_start = std::isfinite(sp_dist_[i]) ? sp_dist_[i] : DBL_MAX;
      nodes_[i]->previous_point = sp_pred_[i] >= 0 ? nodeAt(sp_pred_[i]) : nullptr;
    }
  nodes_[0]->previous_point = nodes_[0];
  Path p;
  p.from_id = 0;
  p.to_id = 1;
  if (!std::isfinite(sp_dist_[1])) {  // no path: length "infinite", empty plan (include/dijkstra.hpp:165-171)
    p.length = DBL_MAX;
    return {p};
  }
  p.length = sp_dist_[1];
  std::vector<Node *> rev;
  for (int c = 1; c != 0; c = sp_pred_[c]) {
    Node *nd = nodeAt(c);
    nd->distance_from_start = sp_dist_[c];
    nd->previous_point = nodeAt(sp_pred_[c]);
    rev.push_back(nd);
  }
  rev.push_back(nodes_[0]);
  p.plan.assign(rev.rbegin(), rev.rend());
  return {p};
}

template <>
void Planner::floodFromSeeds(const std::vector<int> &seed_ids, std::vector<double> &dist, std::vector<int> &pred,
                             std::vector<int> &cluster) {
  const int n = (int)nodes_.size();
  if ((int)csr_row_ptr_.size() != n + 1) throw std::runtime_error("floodFromSeeds: call sampleMultiple first");
  ctp_map *h = handleOf(map_, "floodFromSeeds");
  dist.assign((size_t)n, 0.0);
  std::vector<int32_t> p32((size_t)n), l32((size_t)n), src(seed_ids.begin(), seed_ids.end());
  const int64_t nnz_off[2] = {0, (int64_t)csr_col_.size()};
  if (ctp_sssp(h, 1, n, csr_row_ptr_.data(), csr_col_.data(), csr_w_.data(), nnz_off, src.data(), (int)src.size(),
               dist.data(), p32.data(), l32.data()) != CTP_OK)
    die("ctp_sssp");
  pred.assign(p32.begin(), p32.end());
  cluster.assign(l32.begin(), l32.end());
  cluster_labels_ = cluster;
  for (int i = 0; i < n; i++)
    if (nodes_[i]) nodes_[i]->cluster_id = cluster[i];
}

template <> std::vector<V3> Planner::samplePath(std::vector<Node *> path, const double length_tot, const double num_samples) {
  return map_->samplePath(path, length_tot, num_samples);
}

// shorten_path (:2313-2479) for many polylines at once.  New push-off points become new HeapNodes;
// reused input nodes keep their identity; consecutive nodes of a shortened path are linked in
// visibility_node_ids as the reference does (:2414-2415, :2448-2449).
template <> std::vector<Path> Planner::shorten_paths(const std::vector<Path> &paths, bool forward) {
  if (paths.empty()) return {};
  ctp_map *h = handleOf(map_, "shorten_path");
  std::vector<double> pts, len;
  std::vector<int32_t> off;
  packPaths(paths, pts, off, len);
  const int64_t count = (int64_t)paths.size();
  int32_t cap = 0;
  for (auto &p : paths) cap = std::max<int32_t>(cap, (int32_t)p.plan.size());
  cap += 64;
  for (;;) {
    std::vector<double> out((size_t)count * cap * 3), out_len((size_t)count);
    std::vector<int32_t> origin((size_t)count * cap), out_n((size_t)count), status((size_t)count);
    if (ctp_shorten_path(h, pts.data(), off.data(), len.data(), count, forward ? 1 : 0, min_clearance_,
                         collision_distance_check_, cap, out.data(), origin.data(), out_n.data(), out_len.data(),
                         status.data()) != CTP_OK)
      die("ctp_shorten_path");
    bool grow = false;
    for (int64_t i = 0; i < count; i++) grow |= status[i] == CTP_ERR_CAPACITY;
    if (grow) {
      cap *= 2;
      continue;
    }
    std::vector<Path> res((size_t)count);
    for (int64_t i = 0; i < count; i++) {
      if (status[i] == 1) {  // "no point added to shortened path after collision": original path (:2423-2427)
        INFO("no point added to shortened path after collision");
        res[i] = paths[i];
        continue;
      }
      if (status[i] == CTP_ERR_SAMPLE_PATH) {
        ERROR("samplePath: sampling beyond the path end");
        exit(1);
      }
      if (status[i] != 0) {  // the reference's self-check (:2463-2473) ends the program
        ERROR("shortened path length does not equal lenght of its parts");
        exit(1);
      }
      Path s;
      s.from_id = paths[i].from_id;
      s.to_id = paths[i].to_id;
      s.length = out_len[i];
      for (int32_t a = 0; a < out_n[i]; a++) {
        const int32_t o = origin[(size_t)i * cap + a];
        Node *nd;
        if (o >= 0) nd = paths[i].plan[o];
        else {
          const double *q = &out[((size_t)i * cap + a) * 3];
          nd = new Node(V3(q[0], q[1], q[2]));
        }
        if (!s.plan.empty()) {
          Node *prev = s.plan.back();
          const double d = (nd->data - prev->data).norm();
          prev->visibility_node_ids[nd] = d;
          nd->visibility_node_ids[prev] = d;
        }
        s.plan.push_back(nd);
      }
      res[i] = s;
    }
    return res;
  }
}

template <> Path Planner::shorten_path(Path path, bool forward) { return shorten_paths({path}, forward)[0]; }

template <> std::vector<Path> Planner::removeTooLongPaths(std::vector<Path> paths) {
  std::vector<Path> proned = paths;
  double min_length = DBL_MAX;
  for (auto &p : proned)
    if (p.length < min_length && p.length > min_allowed) min_length = p.length;
  for (int i = (int)proned.size() - 1; i >= 0; i--)
    if (proned[i].length > cutoof_distance_ratio_to_shortest_ * min_length || proned[i].length < min_allowed)
      proned.erase(proned.begin() + i);
  return proned;
}

// removeEquivalentPaths (:2237-2298).  The reference only ever compares ORIGINAL paths (a slot that
// was replaced by a shorter equivalent is never compared again), so the whole upper-triangular
// deformability matrix is one batch and the keep-shortest bookkeeping is replayed on the host.
template <> std::vector<Path> Planner::removeEquivalentPaths(std::vector<Path> paths) {
  const int P = (int)paths.size();
  if (P <= 1) return paths;
  std::vector<std::vector<Node *>> plans;
  std::vector<double> lens;
  for (auto &p : paths) {
    plans.push_back(p.plan);
    lens.push_back(p.length);
  }
  std::vector<int> ia, ib;
  std::vector<double> num;
  for (int i = 0; i < P; i++)
    for (int j = i + 1; j < P; j++) {
      ia.push_back(i);
      ib.push_back(j);
      const double larger = lens[i] > lens[j] ? lens[i] : lens[j];
      num.push_back(std::ceil(larger / collision_distance_check_) + 1);
    }
  const std::vector<uint8_t> def = map_->isDeformablePathBatch(plans, lens, ia, ib, num, min_clearance_, collision_distance_check_);
  auto deformable = [&](int a, int b) {
    if (a > b) std::swap(a, b);
    // pair (a, b), a < b, sits at a*P - a(a+1)/2 + (b - a - 1)
    return def[(size_t)a * P - (size_t)a * (a + 1) / 2 + (size_t)(b - a - 1)] == 1;
  };
  std::vector<int> keep(P);
  for (int i = 0; i < P; i++) keep[i] = i;
  size_t i = 0;
  while (i < keep.size()) {
    size_t shortest = i;
    double shortest_len = lens[keep[i]];
    std::vector<size_t> rm;
    for (size_t j = i + 1; j < keep.size(); j++)
      if (deformable(keep[i], keep[j])) {
        rm.push_back(j);
        if (lens[keep[j]] < shortest_len) {
          shortest = j;
          shortest_len = lens[keep[j]];
        }
      }
    if (shortest != i) keep[i] = keep[shortest];
    for (size_t t = rm.size(); t-- > 0;) keep.erase(keep.begin() + (long)rm[t]);
    i++;
  }
  std::vector<Path> out;
  for (int id : keep) out.push_back(paths[id]);
  return out;
}

template <>
std::vector<uint8_t> Planner::isNewHomotopyClassBatch(const std::vector<Path> &min_paths, const std::vector<Path> &max_paths,
                                                      std::vector<Path> *shortened_min) {
  if (min_paths.size() != max_paths.size()) throw std::runtime_error("isNewHomotopyClassBatch: size mismatch");
  const size_t P = min_paths.size();
  if (P == 0) return {};
  std::vector<Path> mn = shorten_paths(shorten_paths(min_paths, false), true);   // :1702-1703
  std::vector<Path> mx = shorten_paths(shorten_paths(max_paths, true), false);   // :1706-1707
  std::vector<std::vector<Node *>> plans;
  std::vector<double> lens, num;
  std::vector<int> ia, ib;
  for (size_t i = 0; i < P; i++) {
    plans.push_back(mn[i].plan);
    lens.push_back(mn[i].length);
  }
  for (size_t i = 0; i < P; i++) {
    plans.push_back(mx[i].plan);
    lens.push_back(mx[i].length);
    ia.push_back((int)i);
    ib.push_back((int)(P + i));
    const double larger = mn[i].length > mx[i].length ? mn[i].length : mx[i].length;  // std::max (:1711-1712)
    num.push_back(std::ceil(larger / collision_distance_check_) + 1);
  }
  if (shortened_min) *shortened_min = mn;
  return map_->isDeformablePathBatch(plans, lens, ia, ib, num, min_clearance_, collision_distance_check_);
}

template <> bool Planner::isNewSeedAccepted(Node *candidate, const std::vector<Node *> &seeds) {
  // sequential form (:1861-1882): stop at the first seed that is a roadmap neighbour or visible; here all
  // segment checks go out in one batch and the same decision is read off the verdicts
  materializeAdjacency();
  std::vector<V3> a, b;
  for (Node *sd : seeds) {
    if (sd->visibility_node_ids.count(candidate) > 0) return false;
    a.push_back(sd->data);
    b.push_back(candidate->data);
  }
  if (seeds.empty()) return false;
  const std::vector<uint8_t> free = map_->isSimplePathFreeBetweenNodesBatch(a, b, min_clearance_, collision_distance_check_);
  int blocked = 0;
  for (uint8_t f : free) {
    if (f) return false;
    blocked++;
  }
  return blocked >= 2;
}

// ---- CSV writers: node rows "x,y,z", edge rows "x0,y0,z0,x1,y1,z1" (:2066-2150; consumer
// columns_2(...).py:20-33 tells them apart by column count).  Default ostream precision.
template <> void Planner::saveRoadmapDense(std::string filename) {
  materializeAdjacency();  // (creates every HeapNode first)
  std::ofstream f(filename);
  if (!f.is_open()) return;
  for (auto *n : nodes_) f << n->data(0) << "," << n->data(1) << "," << n->data(2) << std::endl;
  for (auto *n : nodes_)
    for (auto &nb : n->visibility_node_ids)
      f << n->data(0) << "," << n->data(1) << "," << n->data(2) << "," << nb.first->data(0) << "," << nb.first->data(1) << ","
        << nb.first->data(2) << std::endl;
}

template <>
void Planner::saveRoadmapCSR(std::string filename, const double *nodes, int n, const int32_t *row_ptr, const int32_t *col) {
  std::ofstream f(filename);
  if (!f.is_open()) return;
  for (int i = 0; i < n; i++) f << nodes[3 * i] << "," << nodes[3 * i + 1] << "," << nodes[3 * i + 2] << std::endl;
  for (int i = 0; i < n; i++)
    for (int32_t e = row_ptr[i]; e < row_ptr[i + 1]; e++) {
      const double *a = nodes + 3 * i, *b = nodes + 3 * col[e];
      f << a[0] << "," << a[1] << "," << a[2] << "," << b[0] << "," << b[1] << "," << b[2] << std::endl;
    }
}

template <> void Planner::savePath(std::string filename, std::vector<Node *> path) {
  std::ofstream f(filename);
  if (!f.is_open()) return;
  for (size_t i = 1; i < path.size(); i++)
    f << path[i - 1]->data(0) << "," << path[i - 1]->data(1) << "," << path[i - 1]->data(2) << "," << path[i]->data(0) << ","
      << path[i]->data(1) << "," << path[i]->data(2) << std::endl;
}

template <> void Planner::savePathSamples(std::string filename, std::vector<V3> path) {
  std::ofstream f(filename);
  if (!f.is_open()) return;
  for (size_t i = 1; i < path.size(); i++)
    f << path[i - 1](0) << "," << path[i - 1](1) << "," << path[i - 1](2) << "," << path[i](0) << "," << path[i](1) << ","
      << path[i](2) << std::endl;
}

// static entry point (:2489-2684): per consecutive gate pair, PRM build -> Dijkstra -> [cluster stage]
// -> shorten both directions -> prune long -> prune equivalent -> shorten again -> prune equivalent
// -> sort by length.  Phase timers keep the reference's names (:2561,2568,2605).
template <>
std::vector<std::vector<Path>> Planner::find_geometrical_paths(const YAML::Node &planner_config, std::shared_ptr<BaseMap> map,
                                                               std::vector<V3> &gates, std::string output_folder) {
  auto now = [] { return std::chrono::high_resolution_clock::now(); };
  auto ms = [](auto a, auto b) { return std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count() * 1e-6; };
  std::vector<std::vector<Path>> paths_between_gates;
  if (gates.size() < 2) return paths_between_gates;
  const int num_samples_between_gate = (int)loadParam<double>(planner_config, "num_samples_between_gate");
  paths_between_gates.resize(gates.size() - 1);
  for (size_t i = 0; i + 1 < gates.size(); i++) {
    Planner prm(planner_config, map, output_folder, gates[i], gates[i + 1]);
    prm.setBorders(map->getMinPos(), map->getMaxPos());
    if (i < gate_candidate_streams.size() && !gate_candidate_streams[i].empty())
      prm.setCandidateStream(std::move(gate_candidate_streams[i]));  // consumed: one stream serves one call
    auto t0 = now();
    prm.sampleMultiple(num_samples_between_gate);
    auto t1 = now();
    INFO_GREEN("prm time clustering " << ms(t0, t1) << " ms");
    std::vector<Path> shortest = prm.findShortestPath();
    auto t2 = now();
    INFO_GREEN("djikstra time clustering " << ms(t1, t2) << " ms");
    if (shortest.empty() || shortest[0].plan.empty()) return paths_between_gates;  // :2569-2572
    prm.max_path_length_ = shortest[0].length * prm.max_path_length_ratio_;
    last_phase_times = PhaseTimes();
    last_phase_times.prm = ms(t0, t1);
    last_phase_times.dijkstra = ms(t1, t2);
    std::vector<Path> found = cluster_stage(prm, shortest[0]);
    auto t3 = now();
    if (!output_folder.empty()) prm.saveRoadmapFromCSR(output_folder + "roadmap_all" + std::to_string(i) + ".csv");
    INFO_GREEN("algorithm time clustering " << ms(t0, now()) << " ms");
    auto t4 = now();
    auto both_ways = [](std::vector<Path> in) {  // shorten_path(false) then shorten_path(true) (:2612-2625)
      return shorten_paths(shorten_paths(in, false), true);
    };
    std::vector<Path> shortened = both_ways(found);
    shortened = removeTooLongPaths(shortened);
    shortened = removeEquivalentPaths(shortened);
    shortened = both_ways(shortened);
    shortened = removeEquivalentPaths(shortened);
    std::sort(shortened.begin(), shortened.end(), [](const Path &a, const Path &b) { return a.length < b.length; });
    {
      PhaseTimes &pt = last_phase_times;
      const PhaseTimes in_cluster = prm.cluster_phase_times_;
      pt.wavefront = in_cluster.wavefront;
      pt.min_tours = in_cluster.min_tours;
      pt.dfs_and_stitch = ms(t2, t3) - in_cluster.wavefront - in_cluster.min_tours;
      pt.shorten_and_filter = ms(t4, now());
      pt.total = ms(t0, t3) + pt.shorten_and_filter;
      pt.cluster_paths = (int)found.size();
      pt.seeds = prm.cluster_stats_.seeds;
      pt.refloods = prm.cluster_stats_.refloods;
      pt.homotopy_checks = prm.cluster_stats_.homotopy_checks;
      pt.capped_refloods = prm.cluster_stats_.capped_refloods;
    }
    paths_between_gates[i] = shortened;
  }
  return paths_between_gates;
}

template class TopologicalPRMClustering<Vector<3>>;
