// yaml_lite.hpp -- the subset of yaml-cpp's YAML::Node the planner's configuration needs
// (src/main.cpp:209-243, include/topological_prm_clustering.hpp:257-269,2521-2522): flat scalars,
// one or two levels of nested maps by indentation, inline [a, b, c] float lists, '#' comments.
// yaml-cpp itself is not in this image.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace YAML {
class Node {
 public:
  Node() : defined_(false) {}
  bool IsDefined() const { return defined_; }
  explicit operator bool() const { return defined_; }
  const Node &operator[](const std::string &key) const {
    static const Node undefined;
    auto it = children_.find(key);
    return it == children_.end() ? undefined : *it->second;
  }
  template <typename T> T as() const;
  const std::string &Scalar() const { return scalar_; }
  // construction (used by the loader and by callers that build a config in code)
  Node &child(const std::string &key) {
    auto &p = children_[key];
    if (!p) p = std::make_shared<Node>();
    p->defined_ = true;
    defined_ = true;
    return *p;
  }
  void set(const std::string &scalar) { scalar_ = scalar; defined_ = true; }

 private:
  bool defined_;
  std::string scalar_;
  std::map<std::string, std::shared_ptr<Node>> children_;
};

namespace detail {
inline std::string trim(const std::string &s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
inline std::string unquote(const std::string &s) {
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\'')))
    return s.substr(1, s.size() - 2);
  return s;
}
}  // namespace detail

template <> inline std::string Node::as<std::string>() const {
  if (!defined_) throw std::runtime_error("yaml: undefined node");
  return detail::unquote(scalar_);
}
template <> inline double Node::as<double>() const {
  if (!defined_) throw std::runtime_error("yaml: undefined node");
  return std::strtod(scalar_.c_str(), nullptr);
}
template <> inline float Node::as<float>() const { return (float)as<double>(); }
template <> inline int Node::as<int>() const {
  if (!defined_) throw std::runtime_error("yaml: undefined node");
  return (int)std::strtol(scalar_.c_str(), nullptr, 10);
}
template <> inline bool Node::as<bool>() const {
  if (!defined_) throw std::runtime_error("yaml: undefined node");
  const std::string s = detail::unquote(scalar_);
  return s == "true" || s == "True" || s == "TRUE" || s == "yes" || s == "on" || s == "1";
}
template <> inline std::vector<double> Node::as<std::vector<double>>() const {
  if (!defined_) throw std::runtime_error("yaml: undefined node");
  std::vector<double> out;
  std::string s = scalar_;
  for (char &c : s)
    if (c == '[' || c == ']' || c == ',') c = ' ';
  std::istringstream is(s);
  double v;
  while (is >> v) out.push_back(v);
  return out;
}

inline Node Load(std::istream &in) {
  Node root;
  std::vector<std::pair<int, Node *>> stack;  // (indent, node)
  stack.push_back({-1, &root});
  std::string line;
  while (std::getline(in, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    if (detail::trim(line).empty()) continue;
    const int indent = (int)line.find_first_not_of(" ");
    const size_t colon = line.find(':');
    if (colon == std::string::npos) continue;
    const std::string key = detail::unquote(detail::trim(line.substr(0, colon)));
    const std::string val = detail::trim(line.substr(colon + 1));
    while (stack.size() > 1 && stack.back().first >= indent) stack.pop_back();
    Node &n = stack.back().second->child(key);
    if (val.empty()) stack.push_back({indent, &n});
    else n.set(val);
  }
  return root;
}
inline Node LoadFile(const std::string &filename) {
  std::ifstream f(filename);
  if (!f) throw std::runtime_error("yaml: cannot open " + filename);
  return Load(f);
}
inline Node Load(const std::string &text) {
  std::istringstream is(text);
  return Load(is);
}
}  // namespace YAML
