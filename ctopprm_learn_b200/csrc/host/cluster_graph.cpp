// cluster_graph.cpp -- TopologicalPRMClustering<Vector<3>>::clusterGraph (reference
// include/topological_prm_clustering.hpp:1832-1952) on the device roadmap.
//
// What runs where:
//   * every flood -- wavefrontFill's multi-source Dijkstra (:826-962) and each addCentroid re-flood (:1276-1374),
//     together with the inter-cluster connection records the sequential loop would have appended -- is one
//     ctp_cluster_flood call on the roadmap kept on the device (ctp_roadmap);
//   * every path the stage shortens or compares (isNewHomotopyClass :1664-1726, findMinClustertours :1740-1793) goes
//     through the batched ctp_shorten_path / ctp_deformable kernels: the unchecked cluster pairs of one pass of the
//     ratio loop (:1030-1060) are evaluated together, their results taken up in the loop's own order;
//   * the bookkeeping between the floods -- min / max connection per cluster pair, the ratio queue, the stale
//     connection sweep of addCentroid (:1377-1427), the depth-first search over at most max_clusters cluster nodes
//     (:626-737) and the stitching (:1925-1952) -- is a few hundred records and stays here on the host.
//
// Where the reference has no reproducible order of its own (pointer-keyed unordered_map / unordered_set iteration,
// heap ties, random extra seeds) this layer follows the rules stated in include/ctopprm_b200.h (ctp_cluster_flood) and
// neighbours and connected clusters in ascending id, pops in ascending (distance, id), extra seeds from
// setSeedCandidates() or a seeded generator.  The node-count cap of addCentroid (:1319) is not applied; re-floods
// that would have hit it are counted in ClusterStats::capped_refloods.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <memory>
#include <random>

#include "ctopprm/topological_prm_clustering.hpp"

typedef Vector<3> V3;
typedef HeapNode<V3> Node;
typedef path_with_length<V3> Path;
typedef TopologicalPRMClustering<V3> Planner;

namespace {

struct Poly {  // a polyline by value: points back to back, and the length the reference carries along with it
  std::vector<double> pts;
  double len = 0;
  int n() const { return (int)(pts.size() / 3); }
};

struct Conn {  // ClusterConnection (:31-38); node -1 = NULL
  int node1 = -1, node2 = -1;
  double distance = 0;
  bool checked = false, deformation = false;
};
struct Link {
  int a, b;
  double d;
};

[[noreturn]] void dieCluster(const char *what) {
  ERROR(what << ": " << ctp_last_error());
  exit(1);
}

double polyLength(const std::vector<double> &p) {  // calc_path_length (include/dijkstra.hpp:27-38)
  double s = 0;
  for (size_t i = 3; i + 2 < p.size(); i += 3) s += (V3(p[i], p[i + 1], p[i + 2]) - V3(p[i - 3], p[i - 2], p[i - 1])).norm();
  return s;
}

// shorten_path (:2313-2479) for a batch of polylines held by value; "no point added" keeps the original (:2423-2427)
std::vector<Poly> shortenBatch(ctp_map *h, const std::vector<Poly> &in, bool forward, double clr, double cdc,
                               const std::vector<uint8_t> *forward_each = nullptr) {
  const int64_t count = (int64_t)in.size();
  if (count == 0) return {};
  std::vector<double> pts, len;
  std::vector<int32_t> off(1, 0);
  int32_t cap = 0;
  for (auto &p : in) {
    pts.insert(pts.end(), p.pts.begin(), p.pts.end());
    off.push_back((int32_t)(pts.size() / 3));
    len.push_back(p.len);
    cap = std::max<int32_t>(cap, p.n());
  }
  cap += 64;
  for (;;) {
    std::vector<double> out((size_t)count * cap * 3), out_len((size_t)count);
    std::vector<int32_t> origin((size_t)count * cap), out_n((size_t)count), status((size_t)count);
    if (ctp_shorten_path_ex(h, pts.data(), off.data(), len.data(), count, forward ? 1 : 0,
                            forward_each ? forward_each->data() : nullptr, clr, cdc, cap, out.data(), origin.data(),
                            out_n.data(), out_len.data(), status.data()) != CTP_OK)
      dieCluster("ctp_shorten_path");
    bool grow = false;
    for (int64_t i = 0; i < count; i++) grow |= status[i] == CTP_ERR_CAPACITY;
    if (grow) {
      cap *= 2;
      continue;
    }
    std::vector<Poly> res((size_t)count);
    for (int64_t i = 0; i < count; i++) {
      if (status[i] == 1) {
        res[i] = in[i];
        continue;
      }
      if (status[i] != 0) {  // samplePath overrun / the self-check of :2463-2473: the reference ends the program
        ERROR("shortened path length does not equal lenght of its parts");
        exit(1);
      }
      res[i].pts.assign(out.begin() + (size_t)i * cap * 3, out.begin() + ((size_t)i * cap + out_n[i]) * 3);
      res[i].len = out_len[i];
    }
    return res;
  }
}

// isDeformablePath (src/base_map.cpp:160-192) for pairs (a[i], b[i])
std::vector<uint8_t> deformableBatch(ctp_map *h, const std::vector<Poly> &a, const std::vector<Poly> &b, double clr, double cdc) {
  const int64_t P = (int64_t)a.size();
  if (P == 0) return {};
  std::vector<double> pts, len, num;
  std::vector<int32_t> off(1, 0), ia, ib;
  auto add = [&](const Poly &p) {
    pts.insert(pts.end(), p.pts.begin(), p.pts.end());
    off.push_back((int32_t)(pts.size() / 3));
    len.push_back(p.len);
  };
  for (auto &p : a) add(p);
  for (auto &p : b) add(p);
  for (int64_t i = 0; i < P; i++) {
    ia.push_back((int32_t)i);
    ib.push_back((int32_t)(P + i));
    const double larger = a[i].len > b[i].len ? a[i].len : b[i].len;  // std::max (:1711-1712)
    num.push_back(std::ceil(larger / cdc) + 1);
  }
  std::vector<uint8_t> out((size_t)P);
  if (ctp_deformable(h, pts.data(), off.data(), len.data(), 2 * P, ia.data(), ib.data(), num.data(), P, clr, cdc, out.data()) != CTP_OK)
    dieCluster("ctp_deformable");
  for (uint8_t v : out)
    if (v > 1) {
      ERROR("samplePath: sampling beyond the path end");
      exit(1);
    }
  return out;
}

// The state clusterGraph keeps between its phases, on flat arrays
struct Stage {
  ctp_map *h = nullptr;
  ctp_roadmap *rm = nullptr;
  int n = 0, C = 0;  // nodes, max_clusters (>= 2)
  const double *xyz = nullptr;
  const int32_t *rp = nullptr, *col = nullptr;
  double clr = 0, cdc = 0;
  std::vector<int32_t> seeds;
  std::vector<double> dist;
  std::vector<int32_t> prev, cl;
  std::vector<Conn> mn, mx;              // C x C, used for i < j
  std::vector<std::vector<Link>> all;    // all_cluster_connections
  std::vector<Poly> min_paths;           // min_cluster_paths_
  // record buffers of a flood, one record per directed edge at most (left uninitialised: tens of megabytes)
  std::unique_ptr<int32_t[]> rec_nodes, rec_label;
  std::unique_ptr<double[]> rec_dist;
  int64_t rec_cap = 0;
  Planner::ClusterStats stats;

  Conn &MN(int i, int j) { return mn[(size_t)i * C + j]; }
  Conn &MX(int i, int j) { return mx[(size_t)i * C + j]; }

  // one flood on the device; returns the number of records now in rec_*
  int flood(int mode) {
    std::vector<int32_t> sd((size_t)C, -1);
    std::copy(seeds.begin(), seeds.end(), sd.begin());
    const int32_t ns = (int32_t)seeds.size();
    int32_t count = 0;
    uint64_t upd = 0;
    const auto t0 = std::chrono::high_resolution_clock::now();
    if (ctp_cluster_flood(rm, sd.data(), &ns, mode, dist.data(), prev.data(), cl.data(), rec_nodes.get(), rec_label.get(),
                          rec_dist.get(), &count, &upd) != CTP_OK)
      dieCluster("ctp_cluster_flood");
    if (getenv("CTOPPRM_TRACE"))
      fprintf(stderr, "[ctopprm] flood mode %d: %.3f ms, %d records, %llu updates\n", mode,
              std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count(), count,
              (unsigned long long)upd);
    if (mode == 0) stats.floods++;
    else {
      stats.refloods++;
      if (upd + 1 >= (uint64_t)n) stats.capped_refloods++;  // :1319 would have stopped improving
    }
    return count;
  }

  // seed -> node1 followed by node2 -> seed (:1669-1676), length recomputed from the points
  Poly connectionPath(int node1, int node2) const {
    std::vector<int> a, b;
    for (int v = node1; v >= 0; v = prev[v]) a.push_back(v);
    for (int v = node2; v >= 0; v = prev[v]) b.push_back(v);
    Poly p;
    p.pts.reserve(3 * (a.size() + b.size()));
    for (size_t t = a.size(); t-- > 0;) p.pts.insert(p.pts.end(), xyz + 3 * (size_t)a[t], xyz + 3 * (size_t)a[t] + 3);
    for (int v : b) p.pts.insert(p.pts.end(), xyz + 3 * (size_t)v, xyz + 3 * (size_t)v + 3);
    p.len = polyLength(p.pts);
    return p;
  }

  // flood part of wavefrontFill (:826-962)
  void wavefrontFlood() {
    for (auto &c : mn) c = Conn{-1, -1, DBL_MAX, false, false};
    for (auto &c : mx) c = Conn{-1, -1, 0.0, false, false};
    for (auto &l : all) l.clear();
    const int m = flood(0);
    for (int t = 0; t < m; t++) {
      const int ex = rec_nodes[2 * t], nb = rec_nodes[2 * t + 1];
      const int cl_ex = cl[ex], cl_nb = rec_label[t];
      const bool nb_low = cl_nb < cl_ex;
      const int lo = nb_low ? cl_nb : cl_ex, hi = nb_low ? cl_ex : cl_nb;
      const int lo_node = nb_low ? nb : ex, hi_node = nb_low ? ex : nb;
      const double d = rec_dist[t];
      all[(size_t)lo * C + hi].push_back(Link{lo_node, hi_node, d});
      Conn &a = MN(lo, hi), &b = MX(lo, hi);
      if (d < a.distance) a.distance = d, a.node1 = lo_node, a.node2 = hi_node;
      if (d > b.distance) b.distance = d, b.node1 = lo_node, b.node2 = hi_node;
    }
  }

  // addCentroid (:1227-1441); the new seed is seeds.back()
  void addCentroid() {
    const int S = (int)seeds.size(), c = S - 1;
    std::vector<std::vector<Link>> fresh((size_t)S);
    std::vector<Conn> nmin((size_t)S, Conn{-1, -1, DBL_MAX, false, false}), nmax((size_t)S, Conn{-1, -1, 0.0, false, false});
    const int m = flood(1);
    for (int t = 0; t < m; t++) {  // :1338-1373
      const int ex = rec_nodes[2 * t], nb = rec_nodes[2 * t + 1], k = rec_label[t];
      const double d = rec_dist[t];
      fresh[(size_t)k].push_back(Link{nb, ex, d});
      if (d < nmin[k].distance) nmin[k].distance = d, nmin[k].node1 = nb, nmin[k].node2 = ex;
      if (d > nmax[k].distance) nmax[k].distance = d, nmax[k].node1 = nb, nmax[k].node2 = ex;
    }
    for (int i = 0; i < c; i++) {  // :1377-1381: the new column of every older row
      MN(i, c) = nmin[i];
      MX(i, c) = nmax[i];
      all[(size_t)i * C + c] = fresh[(size_t)i];
    }
    for (int i = 0; i < S - 1; i++)  // :1390-1427: pairs whose extreme connections lost a node to the new cluster
      for (int j = i + 1; j < S; j++) {
        Conn &a = MN(i, j), &b = MX(i, j);
        if (!(b.distance == 0 || cl[b.node1] != i || cl[b.node2] != j || cl[a.node1] != i || cl[a.node2] != j)) continue;
        a = Conn{-1, -1, DBL_MAX, a.checked, a.deformation};
        b = Conn{-1, -1, 0.0, false, b.deformation};
        std::vector<Link> &l = all[(size_t)i * C + j];
        for (size_t k = l.size(); k-- > 0;) {  // back to front, as the reference erases while it scans
          if (cl[l[k].a] != i || cl[l[k].b] != j) {
            l.erase(l.begin() + (long)k);
          } else {
            if (l[k].d < a.distance) a.distance = l[k].d, a.node1 = l[k].a, a.node2 = l[k].b;
            if (l[k].d > b.distance) b.distance = l[k].d, b.node1 = l[k].a, b.node2 = l[k].b;
          }
        }
      }
  }

  // isNewHomotopyClass (:1664-1726) for a list of cluster pairs in one batch: -> verdicts and the shortened min paths
  void homotopyBatch(const std::vector<std::pair<int, int>> &pairs, std::vector<uint8_t> &is_new, std::vector<Poly> &short_min) {
    std::vector<Poly> a, b;
    const auto t0 = std::chrono::high_resolution_clock::now();
    struct Tr {
      std::chrono::high_resolution_clock::time_point t0;
      size_t pairs;
      ~Tr() {
        if (getenv("CTOPPRM_TRACE"))
          fprintf(stderr, "[ctopprm] homotopy batch of %zu pairs: %.3f ms\n", pairs,
                  std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count());
      }
    } tr{t0, pairs.size()};
    for (auto &pr : pairs) {
      a.push_back(connectionPath(MN(pr.first, pr.second).node1, MN(pr.first, pr.second).node2));
      b.push_back(connectionPath(MX(pr.first, pr.second).node1, MX(pr.first, pr.second).node2));
    }
    // the min path backward then forward (:1702-1703), the max path forward then backward (:1706-1707): the two chains
    // are independent, so each pass is ONE launch over both sets with a direction per polyline
    {
      const size_t P = a.size();
      std::vector<Poly> both(a);
      both.insert(both.end(), b.begin(), b.end());
      std::vector<uint8_t> dir(2 * P);
      for (size_t i = 0; i < P; i++) dir[i] = 0, dir[P + i] = 1;
      both = shortenBatch(h, both, false, clr, cdc, &dir);
      for (size_t i = 0; i < P; i++) dir[i] = 1, dir[P + i] = 0;
      both = shortenBatch(h, both, false, clr, cdc, &dir);
      a.assign(both.begin(), both.begin() + (long)P);
      b.assign(both.begin() + (long)P, both.end());
    }
    is_new = deformableBatch(h, a, b, clr, cdc);
    short_min = a;
    stats.homotopy_checks += (int)pairs.size();
  }

  // wavefrontFill (:789-1170)
  void wavefrontFill(int min_clusters) {
    wavefrontFlood();
    bool ratio = true;
    while (ratio && (int)seeds.size() < C) {
      ratio = false;
      const int S = (int)seeds.size();
      struct Item {
        double r;
        int i, j;
      };
      std::vector<Item> order;
      for (int i = 0; i < S; i++)
        for (int j = i + 1; j < S; j++)
          if (MX(i, j).distance != 0) order.push_back(Item{MX(i, j).distance / MN(i, j).distance, i, j});
      // std::priority_queue<tuple<double,int,int>> hands out the largest tuple first (:1014-1032)
      std::sort(order.begin(), order.end(), [](const Item &x, const Item &y) {
        if (x.r != y.r) return x.r > y.r;
        if (x.i != y.i) return x.i > y.i;
        return x.j > y.j;
      });
      // the pairs this pass may have to test, evaluated together; results are taken up in queue order below
      std::vector<std::pair<int, int>> todo;
      for (auto &it : order)
        if (!MX(it.i, it.j).checked) todo.push_back({it.i, it.j});
      std::vector<uint8_t> verdict;
      std::vector<Poly> short_min;
      if (!todo.empty()) homotopyBatch(todo, verdict, short_min);
      double max_dist = 0;
      int cl_min = -1, cl_max = -1, best_min = -1, best_max = -1;
      for (auto &it : order) {
        cl_min = it.i;
        cl_max = it.j;
        Conn &b = MX(cl_min, cl_max);
        if (!b.checked) {
          const size_t at = (size_t)(std::find(todo.begin(), todo.end(), std::make_pair(cl_min, cl_max)) - todo.begin());
          b.deformation = verdict[at] == 1;
          b.checked = true;
          min_paths[(size_t)cl_min * C + cl_max] = short_min[at];
        }
        if (!b.deformation) {
          ratio = true;
          break;
        } else if (b.distance > max_dist) {
          max_dist = b.distance;
          best_min = cl_min;
          best_max = cl_max;
        }
      }
      if (!ratio) {
        if ((int)seeds.size() > min_clusters || best_min < 0) break;
        cl_min = best_min;
        cl_max = best_max;
        ratio = true;
      }
      const Conn &b = MX(cl_min, cl_max);
      seeds.push_back(dist[b.node1] > dist[b.node2] ? b.node2 : b.node1);  // :1107-1110
      addCentroid();
    }
  }

  // findMinClustertours (:1740-1793): all qualifying pairs in two shortening launches
  void findMinClusterTours() {
    const int S = (int)seeds.size();
    std::vector<std::pair<int, int>> pairs;
    std::vector<Poly> p;
    for (int i = 0; i < S; i++)
      for (int j = i + 1; j < S; j++)
        if (!MX(i, j).checked && MN(i, j).distance != DBL_MAX && MX(i, j).distance > 0) {
          pairs.push_back({i, j});
          p.push_back(connectionPath(MN(i, j).node1, MN(i, j).node2));
        }
    p = shortenBatch(h, shortenBatch(h, p, true, clr, cdc), false, clr, cdc);
    for (size_t t = 0; t < pairs.size(); t++) {
      min_paths[(size_t)pairs[t].first * C + pairs[t].second] = p[t];
      if (p[t].len < PRECISION) MN(pairs[t].first, pairs[t].second).node1 = -1;
    }
  }
};

// findPathsRecurseMaxLength (:682-737) over the cluster nodes, connected clusters in ascending id
struct ClusterDfs {
  const Stage &st;
  int S, end;
  double max_len;
  std::vector<uint8_t> adj;
  std::vector<std::vector<int>> paths;
  std::vector<double> lens;
  void walk(double cur, std::vector<int> &path, std::vector<uint8_t> &visited) {
    const int last = path.back();
    for (int node = 0; node < S; node++) {
      if (!adj[(size_t)last * S + node]) continue;
      const int lo = std::min(node, last), hi = std::max(node, last);
      const double seg = st.min_paths[(size_t)lo * st.C + hi].len;
      if (node == end) {
        path.push_back(node);
        paths.push_back(path);
        lens.push_back(cur + seg);
        path.pop_back();
      } else if (!visited[node]) {
        const double d = seg + cur;
        if (d > max_len) return;  // :722-725 returns from the whole frame
        visited[node] = 1;
        path.push_back(node);
        walk(d, path, visited);
        path.pop_back();
        visited[node] = 0;
      }
    }
  }
};

}  // namespace

template <> std::vector<Path> Planner::clusterGraph() {
  const int n = (int)nodes_.size();
  if (n < 2 || (int)csr_row_ptr_.size() != n + 1) throw std::runtime_error("clusterGraph: call sampleMultiple first");
  if (!map_ || !map_->deviceMap()) throw std::runtime_error("clusterGraph: the map is not resident on a device");
  auto now = [] { return std::chrono::high_resolution_clock::now(); };
  auto sec = [](auto a, auto b) { return std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count() * 1e-9; };
  Stage st;
  st.h = map_->deviceMap();
  st.n = n;
  st.C = std::max(2, max_clusters_);
  st.clr = min_clearance_;
  st.cdc = collision_distance_check_;
  st.xyz = node_xyz_.data();
  st.rp = csr_row_ptr_.data();
  st.col = csr_col_.data();
  const int64_t nnz_off[2] = {0, (int64_t)csr_col_.size()};
  if (ctp_roadmap_create(st.h, 1, n, csr_row_ptr_.data(), csr_col_.data(), csr_w_.data(), nnz_off, st.C, &st.rm) != CTP_OK)
    dieCluster("ctp_roadmap_create");
  int64_t info[4];
  ctp_roadmap_info(st.rm, info);
  st.rec_cap = info[3];
  st.dist.resize((size_t)n);
  st.prev.resize((size_t)n);
  st.cl.resize((size_t)n);
  st.rec_nodes.reset(new int32_t[(size_t)st.rec_cap * 2]);
  st.rec_label.reset(new int32_t[(size_t)st.rec_cap]);
  st.rec_dist.reset(new double[(size_t)st.rec_cap]);
  st.mn.resize((size_t)st.C * st.C);
  st.mx.resize((size_t)st.C * st.C);
  st.all.resize((size_t)st.C * st.C);
  st.min_paths.resize((size_t)st.C * st.C);

  // initial seeds: start, goal, then candidates that no seed sees and no seed neighbours (:1834-1891)
  st.seeds = {0, 1};
  std::mt19937_64 gen(sampling_seed);
  size_t next_candidate = 0;
  int draws = 0;
  while ((int)st.seeds.size() < num_clusters_ && (int)st.seeds.size() < st.C && n > 2) {
    int r;
    if (!seed_candidates_.empty()) {
      if (next_candidate >= seed_candidates_.size()) break;
      r = seed_candidates_[next_candidate++];
    } else {
      if (draws++ > 64 * n) break;  // the reference would loop forever when no node qualifies
      r = 2 + (int)(gen() % (uint64_t)(n - 2));
    }
    if (r < 2 || r >= n || std::find(st.seeds.begin(), st.seeds.end(), r) != st.seeds.end()) continue;
    bool adjacent = false;
    size_t upto = st.seeds.size();  // the sequential loop stops at the first seed that neighbours r (:1862-1865)
    for (size_t s = 0; s < st.seeds.size() && !adjacent; s++)
      for (int32_t e = st.rp[st.seeds[s]]; e < st.rp[st.seeds[s] + 1]; e++)
        if (st.col[e] == r) {
          adjacent = true;
          upto = s;
          break;
        }
    std::vector<V3> a, b;
    for (size_t s = 0; s < upto; s++) {
      a.push_back(V3(st.xyz[3 * (size_t)st.seeds[s]], st.xyz[3 * (size_t)st.seeds[s] + 1], st.xyz[3 * (size_t)st.seeds[s] + 2]));
      b.push_back(V3(st.xyz[3 * (size_t)r], st.xyz[3 * (size_t)r + 1], st.xyz[3 * (size_t)r + 2]));
    }
    const std::vector<uint8_t> free = a.empty() ? std::vector<uint8_t>() : map_->isSimplePathFreeBetweenNodesBatch(a, b, st.clr, st.cdc);
    bool visible = false;
    int blocked = 0;
    for (uint8_t f : free) {
      if (f) {
        visible = true;
        break;
      }
      blocked++;
    }
    if (!visible && !adjacent && blocked >= 2) st.seeds.push_back(r);
  }

  auto t0 = now();
  st.wavefrontFill(min_clusters_);
  cluster_phase_times_ = PhaseTimes();
  cluster_phase_times_.wavefront = sec(t0, now()) * 1e3;
  INFO_GREEN("computation time wavefront " << sec(t0, now()));
  auto t01 = now();
  st.findMinClusterTours();
  auto t1 = now();
  cluster_phase_times_.min_tours = sec(t01, t1) * 1e3;
  const int S = (int)st.seeds.size();
  ClusterDfs dfs{st, S, st.cl[1], max_path_length_, std::vector<uint8_t>((size_t)S * S, 0), {}, {}};
  for (int i = 0; i < S; i++)  // createClusterNodes (:589-606)
    for (int j = i + 1; j < S; j++) {
      const Conn &a = st.MN(i, j);
      if (a.node1 != -1 && a.node2 != -1 && a.distance != 0) dfs.adj[(size_t)i * S + j] = dfs.adj[(size_t)j * S + i] = 1;
    }
  {
    std::vector<int> path{st.cl[0]};
    std::vector<uint8_t> visited((size_t)S, 0);
    visited[(size_t)st.cl[0]] = 1;
    dfs.walk(0.0, path, visited);
  }
  INFO_GREEN("computation time DFS " << sec(t1, now()));

  // stitching (:1925-1952): the shortened min paths of consecutive clusters, joint node once; new nodes for every
  // point (the roadmap nodes' identity plays no part downstream, only coordinates do)
  std::vector<Path> distinct;
  for (size_t p = 0; p < dfs.paths.size(); p++) {
    std::vector<double> pts;
    for (size_t t = 1; t < dfs.paths[p].size(); t++) {
      const int from = dfs.paths[p][t - 1], to = dfs.paths[p][t];
      const Poly &seg = st.min_paths[(size_t)std::min(from, to) * st.C + std::max(from, to)];
      if (!pts.empty()) pts.resize(pts.size() - 3);
      if (from < to) pts.insert(pts.end(), seg.pts.begin(), seg.pts.end());
      else
        for (int u = seg.n() - 1; u >= 0; u--) pts.insert(pts.end(), seg.pts.begin() + 3 * u, seg.pts.begin() + 3 * u + 3);
    }
    const size_t m = pts.size() / 3;
    if (m == 0) continue;
    if (!(V3(pts[3 * m - 3], pts[3 * m - 2], pts[3 * m - 1]) == end_->data)) continue;  // :1946
    Path out;
    out.from_id = 0;
    out.to_id = 1;
    out.length = dfs.lens[p];
    for (size_t u = 0; u < m; u++) {
      Node *nd = new Node(V3(pts[3 * u], pts[3 * u + 1], pts[3 * u + 2]));
      nd->id = -1;
      out.plan.push_back(nd);
    }
    out.plan.front()->id = 0;
    out.plan.back()->id = 1;
    distinct.push_back(out);
  }
  st.stats.seeds = S;
  cluster_stats_ = st.stats;
  cluster_seeds_.assign(st.seeds.begin(), st.seeds.end());
  cluster_labels_.assign(st.cl.begin(), st.cl.end());
  for (int i = 0; i < n; i++)
    if (nodes_[i]) nodes_[i]->cluster_id = st.cl[i];  // the others take it from cluster_labels_ when they are created
  ctp_roadmap_destroy(st.rm);
  return distinct;
}
