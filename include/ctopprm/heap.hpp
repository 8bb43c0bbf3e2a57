// heap.hpp -- the graph node exchanged across the boundary (reference include/heap.hpp:17-50):
// same fields, same defaults.  The intrusive heap of the reference is not needed by the hot path.
#pragma once
#include <map>
#include <unordered_map>

template <typename T> struct HeapNode {
  HeapNode() : HeapNode(T()) {}
  HeapNode(T _data)
      : id(0), data(_data), distance_from_start(0), previous_point(nullptr), heap_position(0) {}
  bool city_node = false;
  int id;
  T data;
  int cluster_id = -1;
  bool is_border = false;
  std::unordered_map<HeapNode<T> *, double> visibility_node_ids;  // neighbour -> ||to - from||
  std::multimap<double, HeapNode<T> *> dist_sorted_visibility_node_ids;
  double distance_from_start;
  HeapNode<T> *previous_point;
  int heap_position;
};
