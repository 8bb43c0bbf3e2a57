// common.hpp -- dependency-free stand-ins for what the hot path uses from the reference's
// include/common.hpp (which needs Eigen, yaml-cpp, log4cxx, TCLAP): Scalar, Vector<3> with the
// handful of Eigen members the callers touch (include/common.hpp:651-671), the logging macro names
// (:53-107) on stderr, and loadParam (:636-637, declared but never defined in the reference).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "yaml_lite.hpp"

typedef double Scalar;

#ifndef CTOPPRM_QUIET
#define INFO(x) do { std::cerr << x << std::endl; } while (0)
#else
#define INFO(x) do { if (false) std::cerr << x; } while (0)
#endif
#define INFO_GREEN(x) INFO(x)
#define INFO_RED(x) INFO(x)
#define INFO_VAR(x) INFO(#x << " " << (x))
#define ERROR(x) do { std::cerr << "ERROR: " << x << std::endl; } while (0)

namespace ctopprm {
struct Vec3 {
  double v[3];
  Vec3() : v{0, 0, 0} {}
  Vec3(double x, double y, double z) : v{x, y, z} {}
  static Vec3 Zero() { return Vec3(0, 0, 0); }
  static Vec3 Ones() { return Vec3(1, 1, 1); }
  static Vec3 Constant(double c) { return Vec3(c, c, c); }
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  const double *data() const { return v; }
  double *data() { return v; }
  double squaredNorm() const { return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]; }
  double norm() const { return std::sqrt(squaredNorm()); }
  double maxCoeff() const { return std::fmax(std::fmax(v[0], v[1]), v[2]); }
  Vec3 normalized() const {
    const double z = squaredNorm();
    if (z > 0.0) { const double s = std::sqrt(z); return Vec3(v[0] / s, v[1] / s, v[2] / s); }
    return *this;
  }
  void normalize() { *this = normalized(); }
  Vec3 cross(const Vec3 &o) const {
    return Vec3(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
  }
  double dot(const Vec3 &o) const { return (v[0] * o.v[0] + v[1] * o.v[1]) + v[2] * o.v[2]; }
  Vec3 cwiseProduct(const Vec3 &o) const { return Vec3(v[0] * o.v[0], v[1] * o.v[1], v[2] * o.v[2]); }
  Vec3 &operator+=(const Vec3 &o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
  Vec3 &operator-=(const Vec3 &o) { v[0] -= o.v[0]; v[1] -= o.v[1]; v[2] -= o.v[2]; return *this; }
  Vec3 &operator*=(double s) { v[0] *= s; v[1] *= s; v[2] *= s; return *this; }
  bool operator==(const Vec3 &o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
  bool operator!=(const Vec3 &o) const { return !(*this == o); }
  bool allFinite() const { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }
};
inline Vec3 operator+(const Vec3 &a, const Vec3 &b) { return Vec3(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]); }
inline Vec3 operator-(const Vec3 &a, const Vec3 &b) { return Vec3(a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]); }
inline Vec3 operator-(const Vec3 &a) { return Vec3(-a.v[0], -a.v[1], -a.v[2]); }
inline Vec3 operator*(const Vec3 &a, double s) { return Vec3(a.v[0] * s, a.v[1] * s, a.v[2] * s); }
inline Vec3 operator*(double s, const Vec3 &a) { return Vec3(s * a.v[0], s * a.v[1], s * a.v[2]); }
inline Vec3 operator/(const Vec3 &a, double s) { return Vec3(a.v[0] / s, a.v[1] / s, a.v[2] / s); }
inline std::ostream &operator<<(std::ostream &o, const Vec3 &a) { return o << a.v[0] << " " << a.v[1] << " " << a.v[2]; }
template <int N> struct VectorSel;
template <> struct VectorSel<3> { typedef Vec3 type; };
}  // namespace ctopprm

template <int N = 3> using Vector = typename ctopprm::VectorSel<N>::type;

// include/common.hpp:636-637 (declaration only in the reference): YAML scalar -> T, or exit(1)
template <typename T> T loadParam(const YAML::Node &config, const std::string &key) {
  if (!config[key].IsDefined()) {
    ERROR("cannot load param " << key);
    exit(1);
  }
  return config[key].as<T>();
}
// src/common.cpp:16-98 (parseArrayParam): [x, y, z] list -> Vector<3>
inline bool parseArrayParam(const YAML::Node &config, const std::string &key, Vector<3> &out) {
  if (!config[key].IsDefined()) return false;
  const std::vector<double> a = config[key].as<std::vector<double>>();
  if (a.size() != 3) return false;
  out = Vector<3>(a[0], a[1], a[2]);
  return true;
}
