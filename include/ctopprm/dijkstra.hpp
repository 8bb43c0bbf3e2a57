// dijkstra.hpp -- path_with_length (reference include/dijkstra.hpp:23-47) and a host Dijkstra with
// the reference's call shape findPath(start_index, goal_indexes, graph) (:83-176).  The search is
// host glue outside the accelerated path; it consumes the adjacency that
// sampleMultiple re-materialised from the device CSR.
#pragma once
#include <cfloat>
#include <iostream>
#include <queue>
#include <vector>

#include "heap.hpp"

template <typename T> struct path_with_length {
  std::vector<HeapNode<T> *> plan;
  double length = 0;
  int from_id = 0;
  int to_id = 0;
  double calc_path_length() const {
    double s = 0;
    for (size_t i = 1; i < plan.size(); i++) s += (plan[i]->data - plan[i - 1]->data).norm();
    return s;
  }
  void print() const {
    std::cout << "plan:" << std::endl;
    for (auto *p : plan) std::cout << p->id << " ";
    std::cout << std::endl;
  }
};

template <class T> class Dijkstra {
 public:
  // shortest paths from graph[start_index] to every graph[goal]; an unreachable goal yields an entry with
  // an empty plan and length DBL_MAX, as the reference does (:165-171)
  std::vector<path_with_length<T>> findPath(int start_index, std::vector<int> goal_indexes,
                                            std::vector<HeapNode<T> *> &graph) {
    for (auto *n : graph) {
      n->distance_from_start = DBL_MAX;
      n->previous_point = nullptr;
    }
    typedef std::pair<double, HeapNode<T> *> Item;
    std::priority_queue<Item, std::vector<Item>, std::greater<Item>> pq;
    graph[start_index]->distance_from_start = 0;
    graph[start_index]->previous_point = graph[start_index];
    pq.push({0.0, graph[start_index]});
    while (!pq.empty()) {
      const Item it = pq.top();
      pq.pop();
      HeapNode<T> *u = it.second;
      if (it.first > u->distance_from_start) continue;
      for (auto &nb : u->visibility_node_ids) {
        const double nd = u->distance_from_start + nb.second;
        if (nd < nb.first->distance_from_start) {
          nb.first->distance_from_start = nd;
          nb.first->previous_point = u;
          pq.push({nd, nb.first});
        }
      }
    }
    std::vector<path_with_length<T>> out;
    for (int g : goal_indexes) {
      HeapNode<T> *goal = graph[g];
      path_with_length<T> p;
      p.from_id = graph[start_index]->id;
      p.to_id = goal->id;
      if (goal->previous_point == nullptr) {
        p.length = DBL_MAX;
        out.push_back(p);
        continue;
      }
      p.length = goal->distance_from_start;
      std::vector<HeapNode<T> *> rev;
      for (HeapNode<T> *c = goal; c != graph[start_index]; c = c->previous_point) rev.push_back(c);
      rev.push_back(graph[start_index]);
      p.plan.assign(rev.rbegin(), rev.rend());
      out.push_back(p);
    }
    return out;
  }
};
