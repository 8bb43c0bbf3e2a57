// Shim: path_with_length as used by BaseMap::isPathCollisionFree (include/dijkstra.hpp:23-47).
#pragma once
#include "heap.hpp"
template <typename T> struct path_with_length {
  std::vector<HeapNode<T> *> plan;
  double length;
  int from_id, to_id;
};
