// esdf_map.hpp -- the reference's ESDFMap (include/esdf_map.hpp, src/esdf_map.cpp) with the grid
// resident in HBM behind a ctp_map handle.
#pragma once
#include <string>

#include "base_map.hpp"

class ESDFMap : public BaseMap {
 public:
  ESDFMap();
  // four concatenated NPY records: voxels (N,N,N) f4|f8, centroid (3,) f8, extents (3,) f8,
  // resolution () i8 (src/esdf_map.cpp:6-72, written by map.py:124-132).  Missing file ->
  // std::runtime_error as cnpy does (src/esdf_map.cpp:11-12); bad word size -> exit(1) (:66-69).
  void load(std::string filename) override;
  // the same from memory (vox in .npy order x*n*n + y*n + z)
  void loadFromMemory(const float *vox, int n, const Vector<3> &centroid, const Vector<3> &extents);
  // the voxel grid -> distance field step of the map pipeline (map.py:49-73) on the device: exact signed
  // Euclidean distance transform of an n^3 occupancy grid (non-zero = occupied), cubic extents
  void loadFromOccupancy(const unsigned char *occ, int n, const Vector<3> &centroid, const Vector<3> &extents);
  // write the map as the four NPY records of map.py:124-132 (f4 voxels, centroid, extents, resolution)
  void save(const std::string &filename);
  double getClearence(const Vector<3> pos) override;                                  // :130-165
  Vector<3> getMinPos() override;                                                     // :75
  Vector<3> getMaxPos() override;                                                     // :76
  std::pair<Vector<3>, Vector<3>> gradientInVoxelCenter(const Vector<3> pos) override;  // :167-227
  Vector<3> getRealPosFromIndex(Vector<3> pos_ijk);                                   // :123-128
  // device and layout selection for maps loaded afterwards (default: device 0, gather texture)
  static void setDevice(int device) { device_ = device; }
  static void setLayouts(int layouts) { layouts_ = layouts; }

 private:
  Vector<3> centroid, extents;
  double num_vexels_per_axis = 0;
  double max_extents_ = 0;
  static int device_;
  static int layouts_;
};
