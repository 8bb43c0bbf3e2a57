// base_map.hpp -- the reference's BaseMap class surface (include/base_map.hpp:20-119) on top of
// libctopprm_b200.so.  Same member names, argument meaning, return conventions (NaN = outside the
// map, (true, NaN^3) = "no hit") and exit(1) behaviour.  Every query is served by the CUDA path
// through the C ABI (include/ctopprm_b200.h); there is no host implementation of the arithmetic.
// The *Batch members are the batched forms the accelerated planner stages use; the scalar members
// of the reference are batches of one.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "../ctopprm_b200.h"
#include "common.hpp"
#include "dijkstra.hpp"

#define PRECISION_BASE_MAP (10e-6)  // reference include/base_map.hpp:18

class BaseMap {
 public:
  typedef HeapNode<Vector<3>> Node;
  BaseMap() {}
  virtual ~BaseMap();
  virtual void load(std::string filename) {}
  virtual double getClearence(const Vector<3> pos) = 0;
  virtual Vector<3> getMinPos() { return Vector<3>::Zero(); }
  virtual Vector<3> getMaxPos() { return Vector<3>::Zero(); }
  virtual std::pair<Vector<3>, Vector<3>> gradientInVoxelCenter(const Vector<3> pos) {
    return {Vector<3>::Zero(), Vector<3>::Zero()};
  }

  // ---- reference members (src/base_map.cpp) ----
  void fillRandomState(Node *positionToFill, Vector<3> min_position, Vector<3> position_range);      // :19-26
  void fillRandomStateInEllipse(Node *positionToFill, Node *start, Node *goal,
                                const double ellipse_ratio_major_axis_focal_length, bool planar);     // :29-43
  bool isInCollision(Vector<3> object_position, const double clearance);                             // :7-12
  std::pair<bool, Vector<3>> isSimplePathFreeBetweenNodes(Vector<3> actual, Vector<3> neigbour, const double clearance,
                                                          const double collision_distance_check);    // :49-74
  std::pair<bool, Vector<3>> isPathCollisionFree(path_with_length<Vector<3>> path, const double clearance,
                                                 const double collision_distance_check);             // :200-214
  bool isSimplePathFreeBetweenGuards(Vector<3> actual, Vector<3> neigbour, const double clearance,
                                     const double collision_distance_check);                         // :77-97
  bool isDeformablePath(std::vector<Node *> path1, const double length_tot1, std::vector<Node *> path2,
                        const double length_tot2, const double num_collision_check, const double clearance,
                        const double collision_distance_check);                                       // :160-192
  bool isDeformablePathBetween(Node *start, Node *end, Node *between1, Node *between2, const double clearance,
                               const double collision_distance_check);                               // :217-240
  std::vector<Vector<3>> samplePath(std::vector<Node *> path, const double length_tot, const double num_samples);  // :105-157

  // ---- batched forms ----
  std::vector<double> getClearenceBatch(const std::vector<Vector<3>> &pos);
  std::vector<uint8_t> isInCollisionBatch(const std::vector<Vector<3>> &pos, double clearance);
  // free[i] for (a[i] -> b[i]); hit (optional) receives the first colliding sample or NaN^3
  std::vector<uint8_t> isSimplePathFreeBetweenNodesBatch(const std::vector<Vector<3>> &a, const std::vector<Vector<3>> &b,
                                                         double clearance, double collision_distance_check,
                                                         std::vector<Vector<3>> *hit = nullptr, bool guards = false);
  // deformable[p] for pairs (ia[p], ib[p]) of paths; num_check as isDeformablePath's argument
  std::vector<uint8_t> isDeformablePathBatch(const std::vector<std::vector<Node *>> &paths, const std::vector<double> &lengths,
                                             const std::vector<int> &ia, const std::vector<int> &ib,
                                             const std::vector<double> &num_check, double clearance,
                                             double collision_distance_check);

  ctp_map *deviceMap() const { return handle_; }
  void setRandomSeed(uint64_t seed) { seed_ = seed; draws_ = 0; }

 protected:
  ctp_map *handle_ = nullptr;  // owned; created by the subclass's load()
  ctp_map *requireHandle(const char *who) const;
  uint64_t seed_ = 0x9E3779B97F4A7C15ull;
  uint64_t draws_ = 0;
};
