// Shim for the absent cnpy library: NPY v1/v2 header parse + array holder, with exactly the
// members src/esdf_map.cpp and BaseMap::getArray use.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
namespace cnpy {
struct NpyArray {
  NpyArray(const std::vector<size_t> &s, size_t ws, bool fo) : shape(s), word_size(ws), fortran_order(fo) {
    num_vals = 1;
    for (size_t d : shape) num_vals *= d;
    holder = std::make_shared<std::vector<char>>(num_vals * word_size);
  }
  NpyArray() : word_size(0), fortran_order(false), num_vals(0) {}
  template <typename T> T *data() { return reinterpret_cast<T *>(holder->data()); }
  template <typename T> std::vector<T> as_vec() { T *p = data<T>(); return std::vector<T>(p, p + num_vals); }
  size_t num_bytes() const { return holder->size(); }
  std::shared_ptr<std::vector<char>> holder;
  std::vector<size_t> shape;
  size_t word_size;
  bool fortran_order;
  size_t num_vals;
};
inline void parse_npy_header(FILE *fp, size_t &word_size, std::vector<size_t> &shape, bool &fortran_order) {
  unsigned char pre[10];
  if (fread(pre, 1, 10, fp) != 10 || memcmp(pre, "\x93NUMPY", 6) != 0) throw std::runtime_error("parse_npy_header: bad magic");
  size_t hlen = pre[8] | (pre[9] << 8);
  if (pre[6] >= 2) {  // v2: 4-byte length
    unsigned char more[2];
    if (fread(more, 1, 2, fp) != 2) throw std::runtime_error("parse_npy_header: short");
    hlen |= (size_t)more[0] << 16 | (size_t)more[1] << 24;
  }
  std::string h(hlen, '\0');
  if (fread(&h[0], 1, hlen, fp) != hlen) throw std::runtime_error("parse_npy_header: short header");
  size_t p = h.find("fortran_order");
  fortran_order = h.compare(p + 16, 4, "True") == 0;
  p = h.find("(", h.find("shape"));
  size_t q = h.find(")", p);
  shape.clear();
  std::string s = h.substr(p + 1, q - p - 1);
  const char *c = s.c_str();
  while (*c) {
    while (*c && (*c < '0' || *c > '9')) c++;
    if (!*c) break;
    shape.push_back(strtoull(c, const_cast<char **>(&c), 10));
  }
  p = h.find("descr");
  p = h.find("'", p + 6);  // opening quote of the dtype string
  word_size = strtoull(h.c_str() + p + 3, nullptr, 10);  // skip quote, endian char, kind char
}
}  // namespace cnpy
