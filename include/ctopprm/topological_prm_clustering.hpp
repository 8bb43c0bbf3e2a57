// topological_prm_clustering.hpp -- the accelerated stages of the reference's
// TopologicalPRMClustering<T> (include/topological_prm_clustering.hpp) behind the same method
// names: sampleMultiple (:288-401), findShortestPath (:2172), shorten_path (:2313-2479),
// removeTooLongPaths (:2193-2222), removeEquivalentPaths (:2237-2298), the roadmap CSV writers
// (:2066-2150), clusterGraph (:1832-1952: wavefrontFill, addCentroid, isNewHomotopyClass, findMinClustertours,
// the depth-first search over clusters and the stitching -- floods, shortening and mutual-visibility tests on
// the device, see csrc/host/cluster_graph.cpp) and the static entry point find_geometrical_paths (:2489-2684).
// find_geometrical_paths calls the hook `cluster_stage` where the reference calls clusterGraph() (:2591); by
// default the hook is clusterGraph().
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "base_map.hpp"
#include "common.hpp"
#include "dijkstra.hpp"

#define PRECISION (1E-4)  // reference :27

template <class T> class TopologicalPRMClustering {
 public:
  TopologicalPRMClustering(const YAML::Node &planner_config, std::shared_ptr<BaseMap> map, std::string output_folder,
                           T start, T end, double start_yaw_deg = NAN, double end_yaw_deg = NAN);
  void sampleMultiple(const int num_samples);
  void setBorders(Vector<3> min_position, Vector<3> max_position);
  void saveRoadmapDense(std::string filename);
  static void savePath(std::string filename, std::vector<HeapNode<T> *> path);
  // the same file as saveRoadmapDense, written straight from the CSR the device handed back (no pointer graph):
  // n node rows "x,y,z", then one 6-column edge row per CSR entry
  static void saveRoadmapCSR(std::string filename, const double *nodes, int n, const int32_t *row_ptr, const int32_t *col);
  void saveRoadmapFromCSR(std::string filename) {
    saveRoadmapCSR(filename, node_xyz_.data(), (int)nodes_.size(), csr_row_ptr_.data(), csr_col_.data());
  }
  static void savePathSamples(std::string filename, std::vector<T> path);
  static path_with_length<T> shorten_path(path_with_length<T> path, bool forward = true);
  // batched form: all paths in one launch (one CTA per polyline)
  static std::vector<path_with_length<T>> shorten_paths(const std::vector<path_with_length<T>> &paths, bool forward = true);
  static std::vector<path_with_length<T>> removeTooLongPaths(std::vector<path_with_length<T>> paths);
  static std::vector<path_with_length<T>> removeEquivalentPaths(std::vector<path_with_length<T>> paths);
  std::vector<path_with_length<T>> findShortestPath();
  // clusterGraph (:1832-1952): seeds -> wavefrontFill (floods on the device roadmap, centroids added while a max
  // connection is a new homotopy class) -> min tours -> depth-first search over clusters -> stitched paths
  std::vector<path_with_length<T>> clusterGraph();
  struct ClusterStats {
    int seeds = 0;            // clusters at the end
    int floods = 0;           // wavefrontFill floods (1)
    int refloods = 0;         // addCentroid calls
    int capped_refloods = 0;  // re-floods in which the reference's cnt < max_nodes cap (:1319) would have fired
    int homotopy_checks = 0;  // isNewHomotopyClass evaluations (batched; includes speculative ones)
  };
  const ClusterStats &clusterStats() const { return cluster_stats_; }
  // wall-clock milliseconds of the phases of the last find_geometrical_paths call (last gate pair): the reference prints
  // the same spans as "prm time clustering", "djikstra time clustering", "computation time wavefront", ... (:2561-2605)
  struct PhaseTimes {
    double prm = 0, dijkstra = 0, wavefront = 0, min_tours = 0, dfs_and_stitch = 0, shorten_and_filter = 0, total = 0;
    int cluster_paths = 0, seeds = 0, refloods = 0, homotopy_checks = 0, capped_refloods = 0;
  };
  static PhaseTimes last_phase_times;
  const std::vector<int> &clusterSeeds() const { return cluster_seeds_; }
  const std::vector<int> &clusterLabels() const { return cluster_labels_; }
  // node ids tried, in this order, as extra initial seeds when num_clusters > 2 (the reference draws them with
  // randIntMinMax, :1847); when unset they come from a generator seeded with sampling_seed
  void setSeedCandidates(const std::vector<int> &ids) { seed_candidates_ = ids; }
  // length bound of the search over clusters (find_geometrical_paths sets it to shortest * max_path_length_ratio, :2588)
  void setMaxPathLength(double len) { max_path_length_ = len; }
  // the decision at the end of isNewHomotopyClass (:1702-1726) for a batch of (min, max) connection path
  // pairs: shorten min backward then forward, max forward then backward, then isDeformablePath at
  // ceil(longer / cdc) + 1 connectors.  All pairs share four shorten launches and one deformability launch.
  // shortened_min (optional) receives the shortened min paths (the reference stores them in min_cluster_paths_).
  static std::vector<uint8_t> isNewHomotopyClassBatch(const std::vector<path_with_length<T>> &min_paths,
                                                      const std::vector<path_with_length<T>> &max_paths,
                                                      std::vector<path_with_length<T>> *shortened_min = nullptr);
  // the seed test of clusterGraph (:1856-1888): a candidate becomes a new cluster seed iff it is not a roadmap
  // neighbour of any seed, sees none of them along a straight segment, and there are at least two seeds
  bool isNewSeedAccepted(HeapNode<T> *candidate, const std::vector<HeapNode<T> *> &seeds);
  // multi-source flood over the roadmap (the core of wavefrontFill, :789-962): for every node the
  // distance to the nearest seed, the predecessor towards it and the index of that seed
  void floodFromSeeds(const std::vector<int> &seed_ids, std::vector<double> &dist, std::vector<int> &pred,
                      std::vector<int> &cluster);
  static std::vector<Vector<3>> samplePath(std::vector<HeapNode<Vector<3>> *> path, const double length_tot,
                                           const double num_samples);
  static std::vector<std::vector<path_with_length<Vector<3>>>> find_geometrical_paths(
      const YAML::Node &planner_config, std::shared_ptr<BaseMap> map, std::vector<Vector<3>> &gates_with_start_end_poses,
      std::string output_folder);
  double getEllipseRatioMajorAxis() { return ellipse_ratio_major_axis_focal_length_; }
  void setEllipseRatioMajorAxis(const double r) { ellipse_ratio_major_axis_focal_length_ = r; }

  // ---- additions of the accelerated build ----
  // the roadmap as HeapNodes; their visibility_node_ids maps are filled from the CSR on first use (the planner's own
  // stages work on the CSR and never need them)
  const std::vector<HeapNode<T> *> &nodes() {
    materializeAdjacency();
    return nodes_;
  }
  void materializeAdjacency();
  // the HeapNode of roadmap node i, created on first use (coordinates, id, and what findShortestPath / the floods
  // left for it); the planner's own stages work on the coordinate table and the CSR
  HeapNode<T> *nodeAt(int i);
  void materializeNodes();
  const std::vector<double> &nodeCoordinates() const { return node_xyz_; }  // n x 3
  // the CSR roadmap as the device handed it back
  const std::vector<int32_t> &csrRowPtr() const { return csr_row_ptr_; }
  const std::vector<int32_t> &csrCol() const { return csr_col_; }
  const std::vector<double> &csrWeights() const { return csr_w_; }
  // candidate stream for the next sampleMultiple (draw order of fillRandomStateInEllipse); when unset
  // the candidates are generated on the device from `sampling_seed`
  void setCandidateStream(const std::vector<Vector<3>> &cand) { candidate_stream_ = cand; }
  void setCandidateStream(std::vector<Vector<3>> &&cand) { candidate_stream_ = std::move(cand); }
  static uint64_t sampling_seed;
  // find_geometrical_paths: candidate stream of gate pair i (optional; pairs without one draw on the device)
  static std::vector<std::vector<Vector<3>>> gate_candidate_streams;
  // where the reference runs clusterGraph(); receives the planner after Dijkstra and returns the
  // candidate paths to shorten and filter
  static std::function<std::vector<path_with_length<T>>(TopologicalPRMClustering<T> &, const path_with_length<T> &)>
      cluster_stage;
  static constexpr int num_nn = 14;  // reference :291 (13 neighbours + self)

 private:
  std::vector<HeapNode<T> *> nodes_;     // entries 0 (start) and 1 (end) always exist, the others on demand
  std::vector<double> node_xyz_;         // n x 3, as the device returned them
  std::vector<double> sp_dist_;          // findShortestPath: distance from the start / predecessor of every node
  std::vector<int32_t> sp_pred_;
  std::vector<Vector<3>> candidate_stream_;
  // CSR of the roadmap as sampleMultiple received it from the device (input of the device graph search)
  std::vector<int32_t> csr_row_ptr_, csr_col_;
  std::vector<double> csr_w_;
  bool adjacency_ready_ = true;
  std::vector<int> seed_candidates_, cluster_seeds_, cluster_labels_;
  ClusterStats cluster_stats_;
  PhaseTimes cluster_phase_times_;
  int num_clusters_, max_clusters_, min_clusters_;
  double min_ratio_;
  double max_path_length_ratio_;
  double max_path_length_ = 0;
  static std::shared_ptr<BaseMap> map_;
  static std::string output_folder_;
  static double collision_distance_check_;
  static double min_clearance_;
  static double cutoof_distance_ratio_to_shortest_;
  static double min_allowed;
  double ellipse_ratio_major_axis_focal_length_;
  double neighbour_radius_ = 0;  // optional YAML key neighbour_radius (m): radius-neighbour candidates, 0 = plain k nearest
  HeapNode<T> *start_;
  HeapNode<T> *end_;
  bool planar_;
  Vector<3> min_position_, max_position_, position_range_;
  Dijkstra<T> dijkstra;
};
