// esdf_map.cpp -- ESDFMap: .npy loader (format of map.py:124-132) + forwarding to the C ABI.
#include "ctopprm/esdf_map.hpp"

#include <cstdint>
#include <cstring>
#include <memory>

int ESDFMap::device_ = 0;
int ESDFMap::layouts_ = CTP_LAYOUT_TEX | CTP_LAYOUT_PLAIN;

namespace {
struct NpyRecord {
  std::vector<size_t> shape;
  size_t word_size = 0;
  char kind = 'f';
  char order = '<';  // '<' little-endian, '|' not applicable, '>' big-endian, '=' native
  std::vector<char> bytes;
  size_t count() const {
    size_t c = 1;
    for (size_t d : shape) c *= d;
    return c;
  }
};

// one NPY v1/v2 record from the current file position (replaces cnpy::parse_npy_header + fread,
// reference include/base_map.hpp:39-56)
NpyRecord readNpy(FILE *fp) {
  unsigned char pre[10];
  if (fread(pre, 1, 10, fp) != 10 || memcmp(pre, "\x93NUMPY", 6) != 0) throw std::runtime_error("npy: bad magic");
  size_t hlen = pre[8] | ((size_t)pre[9] << 8);
  if (pre[6] >= 2) {
    unsigned char more[2];
    if (fread(more, 1, 2, fp) != 2) throw std::runtime_error("npy: short header");
    hlen |= ((size_t)more[0] << 16) | ((size_t)more[1] << 24);
  }
  std::string h(hlen, '\0');
  if (fread(&h[0], 1, hlen, fp) != hlen) throw std::runtime_error("npy: short header");
  NpyRecord r;
  size_t p = h.find("fortran_order");
  if (p != std::string::npos && h.compare(p + 16, 4, "True") == 0) throw std::runtime_error("npy: fortran order unsupported");
  p = h.find("descr");
  p = h.find('\'', p + 6);
  r.order = h[p + 1];
  r.kind = h[p + 2];
  r.word_size = strtoull(h.c_str() + p + 3, nullptr, 10);
  p = h.find('(', h.find("shape"));
  const size_t e = h.find(')', p);
  const std::string s = h.substr(p + 1, e - p - 1);
  const char *c = s.c_str();
  while (*c) {
    while (*c && (*c < '0' || *c > '9')) c++;
    if (!*c) break;
    r.shape.push_back(strtoull(c, const_cast<char **>(&c), 10));
  }
  r.bytes.resize(r.count() * r.word_size);
  if (fread(r.bytes.data(), 1, r.bytes.size(), fp) != r.bytes.size()) {
    ERROR("npy: short read");  // reference getArray: exit(1)
    exit(1);
  }
  return r;
}
[[noreturn]] void die(const char *what) {
  ERROR(what << ": " << ctp_last_error());
  exit(1);
}
}  // namespace

ESDFMap::ESDFMap() {}

void ESDFMap::load(std::string filename) {
  FILE *fp = fopen(filename.c_str(), "rb");
  if (!fp) throw std::runtime_error("npy_load: Unable to open file " + filename);  // as cnpy (src/esdf_map.cpp:11-12)
  std::unique_ptr<FILE, int (*)(FILE *)> guard(fp, fclose);
  NpyRecord vox = readNpy(fp), cen = readNpy(fp), ext = readNpy(fp);
  readNpy(fp);  // voxel_resolution: read and discarded (src/esdf_map.cpp:25)
  if (vox.shape.size() != 3 || cen.count() != 3 || ext.count() != 3 || cen.word_size != 8 || ext.word_size != 8)
    throw std::runtime_error("ESDFMap::load: unexpected record shapes in " + filename);
  // the reference reads shape[0] for all three axes and reinterprets the bytes as native floats; a record that is
  // not a cube of little-endian floats would be read out of bounds / as garbage, so it is refused here
  if (vox.shape[0] != vox.shape[1] || vox.shape[0] != vox.shape[2])
    throw std::runtime_error("ESDFMap::load: voxel record is not a cube in " + filename);
  for (const NpyRecord *r : {&vox, &cen, &ext})
    if (r->kind != 'f' || r->order == '>')
      throw std::runtime_error("ESDFMap::load: records must be little-endian floating point in " + filename);
  const double *c = reinterpret_cast<const double *>(cen.bytes.data());
  const double *e = reinterpret_cast<const double *>(ext.bytes.data());
  centroid = Vector<3>(c[0], c[1], c[2]);
  extents = Vector<3>(e[0], e[1], e[2]);
  const int n = (int)vox.shape[0];  // the reference takes shape[0] for all axes (src/esdf_map.cpp:24)
  num_vexels_per_axis = n;
  max_extents_ = extents.maxCoeff();
  if (handle_) ctp_map_destroy(handle_);
  handle_ = nullptr;
  int rc;
  if (vox.word_size == 4)
    rc = ctp_map_create(reinterpret_cast<const float *>(vox.bytes.data()), n, c, e, device_, n <= 1024 ? layouts_ : (layouts_ & ~CTP_LAYOUT_TEX) | CTP_LAYOUT_PLAIN, &handle_);
  else if (vox.word_size == 8)
    rc = ctp_map_create_f64(reinterpret_cast<const double *>(vox.bytes.data()), n, c, e, device_, n <= 1024 ? layouts_ : (layouts_ & ~CTP_LAYOUT_TEX) | CTP_LAYOUT_PLAIN, &handle_);
  else {
    ERROR("bad word size " << vox.word_size);  // src/esdf_map.cpp:66-69
    exit(1);
  }
  if (rc != CTP_OK) die("ctp_map_create");
}

void ESDFMap::loadFromMemory(const float *vox, int n, const Vector<3> &c, const Vector<3> &e) {
  centroid = c;
  extents = e;
  num_vexels_per_axis = n;
  max_extents_ = extents.maxCoeff();
  if (handle_) ctp_map_destroy(handle_);
  handle_ = nullptr;
  if (ctp_map_create(vox, n, c.data(), e.data(), device_, n <= 1024 ? layouts_ : (layouts_ & ~CTP_LAYOUT_TEX) | CTP_LAYOUT_PLAIN, &handle_) != CTP_OK)
    die("ctp_map_create");
}

void ESDFMap::loadFromOccupancy(const unsigned char *occ, int n, const Vector<3> &c, const Vector<3> &e) {
  centroid = c;
  extents = e;
  num_vexels_per_axis = n;
  max_extents_ = extents.maxCoeff();
  if (handle_) ctp_map_destroy(handle_);
  handle_ = nullptr;
  if (ctp_map_create_occupancy(occ, n, c.data(), e.data(), device_, n <= 1024 ? layouts_ : (layouts_ & ~CTP_LAYOUT_TEX) | CTP_LAYOUT_PLAIN, nullptr, &handle_) != CTP_OK)
    die("ctp_map_create_occupancy");
}

namespace {
// one NPY v1 record: magic, version 1.0, little-endian header length, dict padded so that the data starts on a
// 64-byte boundary (what np.save writes and np.load / cnpy read back)
void writeNpy(FILE *fp, const char *descr, const std::vector<size_t> &shape, const void *data, size_t bytes) {
  std::string dict = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': (";
  for (size_t i = 0; i < shape.size(); i++) dict += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : "");
  dict += "), }";
  while ((10 + dict.size() + 1) % 64 != 0) dict += ' ';
  dict += '\n';
  const unsigned short hlen = (unsigned short)dict.size();
  const unsigned char pre[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(hlen & 0xff), (unsigned char)(hlen >> 8)};
  if (fwrite(pre, 1, 10, fp) != 10 || fwrite(dict.data(), 1, dict.size(), fp) != dict.size() ||
      (bytes && fwrite(data, 1, bytes, fp) != bytes))
    throw std::runtime_error("npy: short write");
}
}  // namespace

void ESDFMap::save(const std::string &filename) {
  const size_t n = (size_t)num_vexels_per_axis;
  std::vector<float> vox(n * n * n);
  if (ctp_map_download(requireHandle("save"), vox.data()) != CTP_OK) die("ctp_map_download");
  FILE *fp = fopen(filename.c_str(), "wb");
  if (!fp) throw std::runtime_error("ESDFMap::save: unable to open " + filename);
  std::unique_ptr<FILE, int (*)(FILE *)> guard(fp, fclose);
  const long long res = (long long)n;
  writeNpy(fp, "<f4", {n, n, n}, vox.data(), vox.size() * sizeof(float));
  writeNpy(fp, "<f8", {3}, centroid.data(), 24);
  writeNpy(fp, "<f8", {3}, extents.data(), 24);
  writeNpy(fp, "<i8", {}, &res, 8);
}

double ESDFMap::getClearence(const Vector<3> pos) {
  double out;
  if (ctp_clearance(requireHandle("getClearence"), pos.data(), 1, &out) != CTP_OK) die("ctp_clearance");
  return out;
}

Vector<3> ESDFMap::getMinPos() {
  double info[12];
  if (ctp_map_info(requireHandle("getMinPos"), info) != CTP_OK) die("ctp_map_info");
  return Vector<3>(info[0], info[1], info[2]);
}

Vector<3> ESDFMap::getMaxPos() {
  double info[12];
  if (ctp_map_info(requireHandle("getMaxPos"), info) != CTP_OK) die("ctp_map_info");
  return Vector<3>(info[3], info[4], info[5]);
}

std::pair<Vector<3>, Vector<3>> ESDFMap::gradientInVoxelCenter(const Vector<3> pos) {
  Vector<3> g, c;
  if (ctp_gradient(requireHandle("gradientInVoxelCenter"), pos.data(), 1, g.data(), c.data()) != CTP_OK) die("ctp_gradient");
  return {g, c};
}

Vector<3> ESDFMap::getRealPosFromIndex(Vector<3> pos_ijk) {
  // metadata-only affine map (src/esdf_map.cpp:123-128): min_ext + q * E / (N - 1)
  const double dp = max_extents_ / (num_vexels_per_axis - 1.0);
  const Vector<3> min_ext = centroid - Vector<3>::Constant(1.0) * max_extents_ / 2.0;
  return min_ext + pos_ijk * dp;
}
