// Shim: the two data members of HeapNode that src/base_map.cpp touches (include/heap.hpp:17-50).
#pragma once
#include "common.hpp"
template <typename T> struct HeapNode {
  HeapNode() : id(0), data(T()) {}
  HeapNode(T d) : id(0), data(d) {}
  int id;
  T data;
  std::unordered_map<HeapNode<T> *, double> visibility_node_ids;
};
