// Shim for the reference's include/common.hpp, used ONLY to compile the reference's own
// src/esdf_map.cpp and src/base_map.cpp into oracle/_ref/ (see oracle/Makefile).
// The real header cannot be used: it is syntactically broken in this fork (unbalanced
// braces, duplicated typedefs) and needs Eigen, yaml-cpp, log4cxx and TCLAP, none of which
// exist in this image.  Only what the two .cpp files touch is provided: a 3-vector with the
// handful of Eigen members they call, the logging macros as no-ops, and the
// random_ellipsoid_point declaration.  TEST INFRASTRUCTURE, never linked into the product.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <unordered_map>
#include <map>
#include <utility>
#include <vector>

#define INFO(x) do { } while (0)
#define ERROR(x) do { std::cerr << x << std::endl; } while (0)
#define INFO_VAR(x) do { } while (0)

typedef double Scalar;

namespace refshim {
struct Mask3 {
  bool b[3];
  bool any() const { return b[0] || b[1] || b[2]; }
};
struct Arr3 {
  double v[3];
  Mask3 operator<(const Arr3 &o) const { return {{v[0] < o.v[0], v[1] < o.v[1], v[2] < o.v[2]}}; }
  Mask3 operator>(const Arr3 &o) const { return {{v[0] > o.v[0], v[1] > o.v[1], v[2] > o.v[2]}}; }
};
struct Vec3 {
  double v[3];
  Vec3() : v{0, 0, 0} {}
  Vec3(double x, double y, double z) : v{x, y, z} {}
  static Vec3 Ones() { return Vec3(1, 1, 1); }
  static Vec3 Zero() { return Vec3(0, 0, 0); }
  static Vec3 Constant(double c) { return Vec3(c, c, c); }
  static Vec3 Random() {  // Eigen: uniform in [-1,1] from rand()
    auto r = [] { return 2.0 * ((double)std::rand() / (double)RAND_MAX) - 1.0; };
    return Vec3(r(), r(), r());
  }
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  Arr3 array() const { return {{v[0], v[1], v[2]}}; }
  double maxCoeff() const { return std::fmax(std::fmax(v[0], v[1]), v[2]); }
  // reduction order: see oracle/ctp_oracle.h ("parity unpinned" at Eigen)
  double squaredNorm() const { return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]; }
  double norm() const { return std::sqrt(squaredNorm()); }
  Vec3 normalized() const {
    double z = squaredNorm();
    if (z > 0.0) { double s = std::sqrt(z); return Vec3(v[0] / s, v[1] / s, v[2] / s); }
    return *this;
  }
  void normalize() { *this = normalized(); }
  Vec3 cwiseProduct(const Vec3 &o) const { return Vec3(v[0] * o.v[0], v[1] * o.v[1], v[2] * o.v[2]); }
  Vec3 cross(const Vec3 &o) const {
    return Vec3(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
  }
  Vec3 &operator*=(double s) { v[0] *= s; v[1] *= s; v[2] *= s; return *this; }
  Vec3 &operator+=(const Vec3 &o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
  bool operator!=(const Vec3 &o) const { return v[0] != o.v[0] || v[1] != o.v[1] || v[2] != o.v[2]; }
};
inline Vec3 operator+(const Vec3 &a, const Vec3 &b) { return Vec3(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]); }
inline Vec3 operator-(const Vec3 &a, const Vec3 &b) { return Vec3(a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]); }
inline Vec3 operator*(const Vec3 &a, double s) { return Vec3(a.v[0] * s, a.v[1] * s, a.v[2] * s); }
inline Vec3 operator*(double s, const Vec3 &a) { return Vec3(s * a.v[0], s * a.v[1], s * a.v[2]); }
inline Vec3 operator/(const Vec3 &a, double s) { return Vec3(a.v[0] / s, a.v[1] / s, a.v[2] / s); }
}  // namespace refshim

template <int N> struct VectorSel;
template <> struct VectorSel<3> { typedef refshim::Vec3 type; };
template <int N> using Vector = typename VectorSel<N>::type;

// src/common.cpp:256 -- defined in ref_capi.cpp on top of the oracle's restatement
Vector<3> random_ellipsoid_point(Vector<3> A, Vector<3> B, double two_major_length);
