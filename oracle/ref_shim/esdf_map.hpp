// Shim declaration of ESDFMap for compiling the reference's src/esdf_map.cpp unmodified
// (members named as that file uses them; storage is the nested vector of
// include/esdf_map.hpp:43 so the "faithful" CPU timing sees the same pointer chasing).
#pragma once
#include "base_map.hpp"
class ESDFMap : public BaseMap {
 public:
  ESDFMap();
  void load(std::string filename);
  double getClearence(const Vector<3> pos);
  Vector<3> getMinPos();
  Vector<3> getMaxPos();
  std::pair<Vector<3>, Vector<3>> gradientInVoxelCenter(const Vector<3> pos);
  Vector<3> getRealPosFromIndex(Vector<3> pos_ijk);
  double interpolateMapData(Vector<3> pos_ijk);
 private:
  std::vector<std::vector<std::vector<float>>> map_data;
  Vector<3> centroid, extents;
  double num_vexels_per_axis, max_extents_;
  template <typename T> void fill_data(T *data, cnpy::NpyArray &arr) {
    const size_t s0 = arr.shape[0], s1 = arr.shape[1], s2 = arr.shape[2];
    for (size_t x = 0; x < s0; x++)
      for (size_t y = 0; y < s1; y++)
        for (size_t z = 0; z < s2; z++) map_data[x][y][z] = data[x * s0 * s1 + y * s1 + z];
  }
};
