// base_map.cpp -- BaseMap members forwarding to the C ABI (no arithmetic of the path on the host).
#include "ctopprm/base_map.hpp"

#include <cstring>

namespace {
[[noreturn]] void die(const char *what) {
  ERROR(what << ": " << ctp_last_error());
  exit(1);
}
inline const double *flat(const std::vector<Vector<3>> &v) { return v.empty() ? nullptr : v[0].data(); }
inline double *flat(std::vector<Vector<3>> &v) { return v.empty() ? nullptr : v[0].data(); }
static_assert(sizeof(Vector<3>) == 24, "Vector<3> must be three packed doubles");

struct Packed {
  std::vector<double> pts;
  std::vector<int32_t> off;
};
Packed pack(const std::vector<std::vector<BaseMap::Node *>> &paths) {
  Packed p;
  p.off.push_back(0);
  for (auto &path : paths) {
    for (auto *n : path) p.pts.insert(p.pts.end(), n->data.data(), n->data.data() + 3);
    p.off.push_back((int32_t)(p.pts.size() / 3));
  }
  return p;
}
}  // namespace

BaseMap::~BaseMap() {
  if (handle_) ctp_map_destroy(handle_);
}

ctp_map *BaseMap::requireHandle(const char *who) const {
  if (!handle_) throw std::runtime_error(std::string(who) + ": no map loaded on a device (there is no CPU fallback)");
  return handle_;
}

std::vector<double> BaseMap::getClearenceBatch(const std::vector<Vector<3>> &pos) {
  std::vector<double> out(pos.size());
  if (ctp_clearance(requireHandle("getClearenceBatch"), flat(pos), (int64_t)pos.size(), out.data()) != CTP_OK)
    die("ctp_clearance");
  return out;
}

std::vector<uint8_t> BaseMap::isInCollisionBatch(const std::vector<Vector<3>> &pos, double clearance) {
  std::vector<uint8_t> out(pos.size());
  if (ctp_in_collision(requireHandle("isInCollisionBatch"), flat(pos), (int64_t)pos.size(), clearance, out.data()) != CTP_OK)
    die("ctp_in_collision");
  return out;
}

bool BaseMap::isInCollision(Vector<3> object_position, const double clearance) {
  // virtual dispatch kept: a subclass may define clearance differently (src/base_map.cpp:8-10)
  const double c = getClearence(object_position);
  return !std::isfinite(c) || c < clearance;
}

std::vector<uint8_t> BaseMap::isSimplePathFreeBetweenNodesBatch(const std::vector<Vector<3>> &a,
                                                                const std::vector<Vector<3>> &b, double clearance,
                                                                double collision_distance_check,
                                                                std::vector<Vector<3>> *hit, bool guards) {
  if (a.size() != b.size()) throw std::runtime_error("isSimplePathFreeBetweenNodesBatch: size mismatch");
  std::vector<uint8_t> out(a.size());
  if (hit) hit->resize(a.size());
  if (ctp_edge_check(requireHandle("isSimplePathFreeBetweenNodesBatch"), flat(a), flat(b), (int64_t)a.size(), clearance,
                     collision_distance_check, guards ? CTP_EDGE_GUARDS : CTP_EDGE_NODES, 0, out.data(),
                     hit ? flat(*hit) : nullptr, nullptr, nullptr) != CTP_OK)
    die("ctp_edge_check");
  return out;
}

std::pair<bool, Vector<3>> BaseMap::isSimplePathFreeBetweenNodes(Vector<3> actual, Vector<3> neigbour, const double clearance,
                                                                 const double collision_distance_check) {
  std::vector<Vector<3>> hit;
  const auto f = isSimplePathFreeBetweenNodesBatch({actual}, {neigbour}, clearance, collision_distance_check, &hit);
  return {f[0] != 0, hit[0]};
}

bool BaseMap::isSimplePathFreeBetweenGuards(Vector<3> actual, Vector<3> neigbour, const double clearance,
                                            const double collision_distance_check) {
  return isSimplePathFreeBetweenNodesBatch({actual}, {neigbour}, clearance, collision_distance_check, nullptr, true)[0] != 0;
}

std::pair<bool, Vector<3>> BaseMap::isPathCollisionFree(path_with_length<Vector<3>> path, const double clearance,
                                                        const double collision_distance_check) {
  // all segments in one batch; the first blocked one in path order decides (src/base_map.cpp:203-212)
  std::vector<Vector<3>> a, b, hit;
  for (size_t i = 1; i < path.plan.size(); i++) {
    a.push_back(path.plan[i - 1]->data);
    b.push_back(path.plan[i]->data);
  }
  const double nan = std::nan("");
  if (a.empty()) return {true, Vector<3>(nan, nan, nan)};
  const auto f = isSimplePathFreeBetweenNodesBatch(a, b, clearance, collision_distance_check, &hit);
  for (size_t i = 0; i < f.size(); i++)
    if (!f[i]) return {false, hit[i]};
  return {true, Vector<3>(nan, nan, nan)};
}

std::vector<Vector<3>> BaseMap::samplePath(std::vector<Node *> path, const double length_tot, const double num_samples) {
  const Packed p = pack({path});
  const int64_t k = num_samples > 0 ? (int64_t)num_samples : 0;
  std::vector<Vector<3>> out((size_t)k);
  const int64_t out_off[2] = {0, k};
  int32_t status = 0;
  if (ctp_sample_path(requireHandle("samplePath"), p.pts.data(), p.off.data(), &length_tot, &num_samples, 1,
                      k ? out[0].data() : nullptr, out_off, &status) != CTP_OK)
    die("ctp_sample_path");
  if (status != 0) {  // reference: ERROR + exit(1) when the cursor overruns the path (src/base_map.cpp:127-133)
    ERROR("samplePath: sampling beyond the path end");
    exit(1);
  }
  return out;
}

std::vector<uint8_t> BaseMap::isDeformablePathBatch(const std::vector<std::vector<Node *>> &paths,
                                                    const std::vector<double> &lengths, const std::vector<int> &ia,
                                                    const std::vector<int> &ib, const std::vector<double> &num_check,
                                                    double clearance, double collision_distance_check) {
  const Packed p = pack(paths);
  std::vector<int32_t> a(ia.begin(), ia.end()), b(ib.begin(), ib.end());
  std::vector<uint8_t> out(ia.size());
  if (ia.empty()) return out;
  if (ctp_deformable(requireHandle("isDeformablePathBatch"), p.pts.data(), p.off.data(), lengths.data(), (int64_t)paths.size(),
                     a.data(), b.data(), num_check.data(), (int64_t)ia.size(), clearance, collision_distance_check,
                     out.data()) != CTP_OK)
    die("ctp_deformable");
  for (uint8_t v : out)
    if (v > 1) {  // samplePath overrun inside the pair: the reference exits
      ERROR("isDeformablePath: sampling beyond the path end");
      exit(1);
    }
  return out;
}

bool BaseMap::isDeformablePath(std::vector<Node *> path1, const double length_tot1, std::vector<Node *> path2,
                               const double length_tot2, const double num_collision_check, const double clearance,
                               const double collision_distance_check) {
  return isDeformablePathBatch({path1, path2}, {length_tot1, length_tot2}, {0}, {1}, {num_collision_check}, clearance,
                               collision_distance_check)[0] == 1;
}

bool BaseMap::isDeformablePathBetween(Node *start, Node *end, Node *between1, Node *between2, const double clearance,
                                      const double collision_distance_check) {
  // two 3-node paths start-between-end, compared at ceil(longer / cdc) + 1 connectors (src/base_map.cpp:217-240)
  std::vector<Node *> p1 = {start, between1, end}, p2 = {start, between2, end};
  const double l1 = (between1->data - start->data).norm() + (end->data - between1->data).norm();
  const double l2 = (between2->data - start->data).norm() + (end->data - between2->data).norm();
  const double num = std::ceil((l1 > l2 ? l1 : l2) / collision_distance_check) + 1;
  return isDeformablePath(p1, l1, p2, l2, num, clearance, collision_distance_check);
}

void BaseMap::fillRandomState(Node *positionToFill, Vector<3> min_position, Vector<3> position_range) {
  // uniform in the box (src/base_map.cpp:19-26); host-side draw, no map arithmetic involved
  auto next = [this]() {
    uint64_t z = (seed_ += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
  };
  positionToFill->data = min_position + Vector<3>(next(), next(), next()).cwiseProduct(position_range);
}

void BaseMap::fillRandomStateInEllipse(Node *positionToFill, Node *start, Node *goal,
                                       const double ellipse_ratio_major_axis_focal_length, bool planar) {
  // one draw of the device generator; candidate index = number of draws so far
  double out[3];
  // the generator is counter based: (seed, query 0, candidate draws_) -> skip to the next counter by
  // asking for one candidate with a per-draw seed
  if (ctp_sample_ellipsoid(requireHandle("fillRandomStateInEllipse"), start->data.data(), goal->data.data(), 1, 1,
                           ellipse_ratio_major_axis_focal_length, planar ? 1 : 0, seed_ + draws_, out) != CTP_OK)
    die("ctp_sample_ellipsoid");
  draws_++;
  positionToFill->data = Vector<3>(out[0], out[1], out[2]);
}
