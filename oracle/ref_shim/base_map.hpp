// Shim declaration of BaseMap for compiling the reference's src/base_map.cpp unmodified.
// Member functions are DEFINED by the reference source; only their declarations (taken
// from the signatures at src/base_map.cpp:7,19,29,49,77,105,160,200,217) live here.
#pragma once
#include "cnpy.h"
#include "common.hpp"
#include "dijkstra.hpp"
#define PRECISION_BASE_MAP (10e-6)
class BaseMap {
 public:
  typedef HeapNode<Vector<3>> Node;
  virtual ~BaseMap() {}
  virtual void load(std::string) {}
  virtual double getClearence(const Vector<3> pos) = 0;
  virtual Vector<3> getMinPos() { return Vector<3>::Zero(); }
  virtual Vector<3> getMaxPos() { return Vector<3>::Zero(); }
  virtual std::pair<Vector<3>, Vector<3>> gradientInVoxelCenter(const Vector<3>) { return {Vector<3>::Zero(), Vector<3>::Zero()}; }
  static cnpy::NpyArray getArray(FILE *fp) {
    std::vector<size_t> shape; size_t ws; bool fo;
    cnpy::parse_npy_header(fp, ws, shape, fo);
    cnpy::NpyArray a(shape, ws, fo);
    if (fread(a.data<char>(), 1, a.num_bytes(), fp) != a.num_bytes()) throw std::runtime_error("getArray: short read");
    return a;
  }
  void fillRandomState(Node *, Vector<3>, Vector<3>);
  void fillRandomStateInEllipse(Node *, Node *, Node *, const double, bool);
  bool isInCollision(Vector<3>, const double);
  std::pair<bool, Vector<3>> isSimplePathFreeBetweenNodes(Vector<3>, Vector<3>, const double, const double);
  std::pair<bool, Vector<3>> isPathCollisionFree(path_with_length<Vector<3>>, const double, const double);
  bool isSimplePathFreeBetweenGuards(Vector<3>, Vector<3>, const double, const double);
  bool isDeformablePath(std::vector<Node *>, const double, std::vector<Node *>, const double, const double, const double, const double);
  bool isDeformablePathBetween(Node *, Node *, Node *, Node *, const double, const double);
  std::vector<Vector<3>> samplePath(std::vector<Node *>, const double, const double);
};
