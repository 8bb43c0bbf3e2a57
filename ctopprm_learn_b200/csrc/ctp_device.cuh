// ctp_device.cuh -- ESDF lookup and segment-march device code (sm_100a).
//
// Numerics contract: every FP64 operation below is evaluated in the reference's order with
// one rounding per operation.  The translation unit is compiled with -fmad=false, so the
// compiler never contracts a*b+c into an FMA; IEEE division, sqrt, floor/ceil are exact in
// CUDA.  That is what makes verdicts bit-identical to the CPU restatement (oracle/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CTP_L_PLAIN 1
#define CTP_L_CELL8 2
#define CTP_L_TEX 4
#define CTP_L_BRICK 8

struct MapDev {
  double cx, cy, cz;                   // centroid
  double lox, loy, loz, hix, hiy, hiz; // getMinPos / getMaxPos (src/esdf_map.cpp:75-76)
  double s1;                           // 2.0 / max_extents_      (src/esdf_map.cpp:137)
  double s2;                           // num_vexels_per_axis / 2 (src/esdf_map.cpp:139)
  double nm1;                          // N - 1
  double nd;                           // N
  double dp;                           // max_extents_ / (N - 1)  (src/esdf_map.cpp:124)
  double mex, mey, mez;                // centroid - max_extents_/2 (src/esdf_map.cpp:125)
  int n;
  int box_implied;                     // 1: outside the box => floor(q) >= N-1 or <= 0 on that axis (checked on the host)
  int nb;                              // bricks per axis
  int tex_cmask, tex_cshift;           // x-slab mosaic of the gather texture: tile (x & cmask, x >> cshift)
  int filt_ok;                         // 1: the float32 verdict filter of the flat edge kernel may be used on this map
  double filt_eps;                     // |float32 clearance - reference clearance| bound of the filter (see EdgeFilter)
  double filt_eta;                     // FP64 slack of the index coordinate (voxels) added to every edge's delta
  float filt_g;                        // largest |difference of adjacent voxels| (k_map_stats)
  const float *plain;
  const float *cell8;
  const float *brick;
  cudaTextureObject_t tex;
};

__device__ __forceinline__ double ctp_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// ---- corner fetch, v[dx*4 + dy*2 + dz] -------------------------------------------------
template <int L> struct Corners;

template <> struct Corners<CTP_L_PLAIN> {
  // [x][y][z] exactly as include/esdf_map.hpp:55-66 fills map_data
  static __device__ __forceinline__ void load(const MapDev &M, int x0, int y0, int z0, float v[8]) {
    const size_t n = (size_t)M.n;
    const float *p = M.plain + ((size_t)x0 * n + (size_t)y0) * n + (size_t)z0;
    const float *q = p + n * n;
    v[0] = __ldg(p);
    v[1] = __ldg(p + 1);
    v[2] = __ldg(p + n);
    v[3] = __ldg(p + n + 1);
    v[4] = __ldg(q);
    v[5] = __ldg(q + 1);
    v[6] = __ldg(q + n);
    v[7] = __ldg(q + n + 1);
  }
};

template <> struct Corners<CTP_L_CELL8> {
  // one 32-byte record per cell (x0,y0,z0), x0,y0,z0 in [0, n-2]: exactly one DRAM/L2 sector
  static __device__ __forceinline__ void load(const MapDev &M, int x0, int y0, int z0, float v[8]) {
    const size_t n1 = (size_t)(M.n - 1);
    const float4 *p =
        reinterpret_cast<const float4 *>(M.cell8 + (((size_t)x0 * n1 + (size_t)y0) * n1 + (size_t)z0) * 8);
    const float4 a = __ldg(p);
    const float4 b = __ldg(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

template <> struct Corners<CTP_L_TEX> {
  // One 2-D gather texture holding the N x-slabs as a mosaic of (y,z) tiles: slab x sits at tile
  // (x & cmask, x >> cshift), texel (u,v) = (z + tile_u*N, y + tile_v*N).  (Layered arrays reject
  // cudaArrayTextureGather on this driver.)  tld4 returns the 2x2 footprint around
  // (z0+1, y0+1): .x=(z0,y1) .y=(z1,y1) .z=(z1,y0) .w=(z0,y0); it never crosses a tile because
  // z0+1, y0+1 <= N-1 is established by the caller's range test.
  static __device__ __forceinline__ void gather(const MapDev &M, int x, int y0, int z0, float &v00, float &v01,
                                                float &v10, float &v11) {
    const float fu = (float)(z0 + 1 + (x & M.tex_cmask) * M.n);
    const float fv = (float)(y0 + 1 + (x >> M.tex_cshift) * M.n);
    float a0, a1, a2, a3;
    asm volatile("tld4.r.2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6}];"
                 : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3)
                 : "l"(M.tex), "f"(fu), "f"(fv));
    v00 = a3; v01 = a2; v10 = a0; v11 = a1;
  }
  static __device__ __forceinline__ void load(const MapDev &M, int x0, int y0, int z0, float v[8]) {
    gather(M, x0, y0, z0, v[0], v[1], v[2], v[3]);
    gather(M, x0 + 1, y0, z0, v[4], v[5], v[6], v[7]);
  }
};

template <> struct Corners<CTP_L_BRICK> {
  // 4x4x4 bricks of 256 B: offset = brick*64 + (x&3)*16 + (y&3)*4 + (z&3)
  static __device__ __forceinline__ size_t addr(const MapDev &M, int x, int y, int z) {
    const size_t nb = (size_t)M.nb;
    return ((((size_t)(x >> 2) * nb + (size_t)(y >> 2)) * nb + (size_t)(z >> 2)) << 6) +
           (size_t)(((x & 3) << 4) | ((y & 3) << 2) | (z & 3));
  }
  static __device__ __forceinline__ void load(const MapDev &M, int x0, int y0, int z0, float v[8]) {
    const float *b = M.brick;
    v[0] = __ldg(b + addr(M, x0, y0, z0));
    v[1] = __ldg(b + addr(M, x0, y0, z0 + 1));
    v[2] = __ldg(b + addr(M, x0, y0 + 1, z0));
    v[3] = __ldg(b + addr(M, x0, y0 + 1, z0 + 1));
    v[4] = __ldg(b + addr(M, x0 + 1, y0, z0));
    v[5] = __ldg(b + addr(M, x0 + 1, y0, z0 + 1));
    v[6] = __ldg(b + addr(M, x0 + 1, y0 + 1, z0));
    v[7] = __ldg(b + addr(M, x0 + 1, y0 + 1, z0 + 1));
  }
};

// ---- ESDFMap::interpolateMapData (src/esdf_map.cpp:78-121) ------------------------------
// Equivalent formulation of the reference's tests, proven in DESIGN.md section 4:
//   floor(q) < 0           <=>  q < 0            (NaN q fails q >= 0 as well)
//   ceil(q)  >= N          <=>  q > N-1
//   x1 == x0 (q integral)   =>  xd = 0/0 = NaN    => result NaN
//   otherwise x1 - x0 == 1  =>  xd = (q - x0)/1 = q - floor(q), exact
// BOX = true: the caller has not applied getClearence's position test (src/esdf_map.cpp:141-146); it is applied
// here, but only where the index tests below do not already imply it (see MapDev::box_implied).
template <int L, bool BOX = false>
__device__ __forceinline__ double esdf_interp(const MapDev &M, double qx, double qy, double qz, double px = 0.0,
                                              double py = 0.0, double pz = 0.0) {
  const int x0 = __double2int_rd(qx), y0 = __double2int_rd(qy), z0 = __double2int_rd(qz);
  const double xd = qx - (double)x0, yd = qy - (double)y0, zd = qz - (double)z0;
  // q in [0, N-1] and not integral  <=>  floor(q) in [0, N-2] and q - floor(q) > 0; a NaN q gives
  // a NaN fraction and fails the '>' (the conversion saturates for out-of-range q, which the
  // unsigned range test rejects)
  const unsigned nm2 = (unsigned)(M.n - 2);
  bool ok = ((unsigned)x0 <= nm2) & ((unsigned)y0 <= nm2) & ((unsigned)z0 <= nm2) & (xd > 0.0) & (yd > 0.0) &
            (zd > 0.0);
  if (BOX) {
    // a position above the box maps to floor(q) >= N-1 and one below it to floor(q) <= 0 (verified on the host for
    // this map, then true for every position because each step of the index map is monotone), so the six
    // comparisons are only needed when an index is 0
    if (ok && (((x0 == 0) | (y0 == 0) | (z0 == 0)) || !M.box_implied))
      ok = (px >= M.lox) & (px <= M.hix) & (py >= M.loy) & (py <= M.hiy) & (pz >= M.loz) & (pz <= M.hiz);
  }
  if (!ok) return ctp_nan();
  float v[8];
  Corners<L>::load(M, x0, y0, z0, v);
  const double omx = 1.0 - xd, omy = 1.0 - yd, omz = 1.0 - zd;
  const double c00 = omx * (double)v[0] + xd * (double)v[4];
  const double c01 = omx * (double)v[1] + xd * (double)v[5];
  const double c10 = omx * (double)v[2] + xd * (double)v[6];
  const double c11 = omx * (double)v[3] + xd * (double)v[7];
  const double c0 = omy * c00 + yd * c10;
  const double c1 = omy * c01 + yd * c11;
  return omz * c0 + zd * c1;
}

// world -> index coordinate along one axis (src/esdf_map.cpp:136-139)
__device__ __forceinline__ double ctp_w2i(double p, double c, double s1, double s2) {
  return ((p - c) * s1 + 1.0) * s2;
}

// ---- ESDFMap::getClearence (src/esdf_map.cpp:130-165) -----------------------------------
// The round()-based index test (:147-155) is implied by interpolateMapData's own range
// tests (q >= 0 => round(q) >= 0; q <= N-1 => round(q) < N), so it is not evaluated.
template <int L>
__device__ __forceinline__ double esdf_clearance(const MapDev &M, double px, double py, double pz) {
  return esdf_interp<L, true>(M, ctp_w2i(px, M.cx, M.s1, M.s2), ctp_w2i(py, M.cy, M.s1, M.s2),
                              ctp_w2i(pz, M.cz, M.s1, M.s2), px, py, pz);
}

// BaseMap::isInCollision (src/base_map.cpp:7-12): !isfinite(c) || c < clearance
__device__ __forceinline__ bool ctp_collides(double c, double clr) { return !(c >= clr) | isinf(c); }

// ---- ESDFMap::gradientInVoxelCenter (src/esdf_map.cpp:167-227) --------------------------
template <int L>
__device__ __forceinline__ void esdf_gradient(const MapDev &M, double px, double py, double pz,
                                              double g[3], double vc[3]) {
  const double q[3] = {ctp_w2i(px, M.cx, M.s1, M.s2), ctp_w2i(py, M.cy, M.s1, M.s2),
                       ctp_w2i(pz, M.cz, M.s1, M.s2)};
  vc[0] = M.mex + q[0] * M.dp;  // getRealPosFromIndex(pos_in_box) (:181, :123-128)
  vc[1] = M.mey + q[1] * M.dp;
  vc[2] = M.mez + q[2] * M.dp;
  double d[3];
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    const double rr = round(q[ax]);
    // (int)round(NaN or out of range) is INT_MIN on x86-64: both neighbours out of range
    const bool sane = (rr >= -2147483648.0) & (rr <= 2147483647.0);
    const double rp = rr + 1.0, rm = rr - 1.0;
    double dval = 0.0, dd = 0.0;
    if (sane & (rp >= 0.0) & (rp < M.nd)) {
      dval += esdf_interp<L>(M, ax == 0 ? q[0] + 1.0 : q[0], ax == 1 ? q[1] + 1.0 : q[1],
                             ax == 2 ? q[2] + 1.0 : q[2]);
      dd += M.nd;
    }
    if (sane & (rm >= 0.0) & (rm < M.nd)) {
      dval -= esdf_interp<L>(M, ax == 0 ? q[0] + -1.0 : q[0], ax == 1 ? q[1] + -1.0 : q[1],
                             ax == 2 ? q[2] + -1.0 : q[2]);
      dd += M.nd;
    }
    d[ax] = dval / dd;
  }
  const double z = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];  // .normalized() (:226)
  if (z > 0.0) {
    const double s = sqrt(z);
    d[0] /= s;
    d[1] /= s;
    d[2] /= s;
  }
  g[0] = d[0];
  g[1] = d[1];
  g[2] = d[2];
}

// ---- point-wise batch kernels ------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(256) k_clearance(MapDev M, const double *__restrict__ xyz,
                                                   int64_t m, double *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = esdf_clearance<L>(M, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

template <int L>
__global__ void __launch_bounds__(256) k_in_collision(MapDev M, const double *__restrict__ xyz,
                                                      int64_t m, double clr,
                                                      uint8_t *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ctp_collides(esdf_clearance<L>(M, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), clr);
}

template <int L>
__global__ void __launch_bounds__(256) k_gradient(MapDev M, const double *__restrict__ xyz,
                                                  int64_t m, double *__restrict__ grad,
                                                  double *__restrict__ centre) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m;
       i += (int64_t)gridDim.x * blockDim.x) {
    double g[3], c[3];
    esdf_gradient<L>(M, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], g, c);
    grad[3 * i] = g[0]; grad[3 * i + 1] = g[1]; grad[3 * i + 2] = g[2];
    centre[3 * i] = c[0]; centre[3 * i + 1] = c[1]; centre[3 * i + 2] = c[2];
  }
}

// Exact division of 0 <= x < 2^31 by a run-time constant 1 <= d < 2^31 as one widening multiply and a shift:
// with s = ceil(log2 d) and mul = floor(2^(31+s) / d) + 1, floor(x / d) = (x * mul) >> (31 + s)  (the excess
// x * e / (d * 2^(31+s)), e <= d <= 2^s, stays below 1/d).  The per-edge kernels (edge march, CSR) split a flat
// edge index into (query, row, slot) with two of these instead of two emulated 64-bit divisions; tests/csrc/div_identity.c
// checks the identity.  `small` = every index of the launch is below 2^31 (else the 64-bit path is taken).
struct CtpDiv {
  uint32_t d, mul;
  int sh;
};
__host__ __device__ __forceinline__ CtpDiv ctp_div_make(uint32_t d) {
  int s = 0;
  while ((1ull << s) < d) s++;
  CtpDiv r;
  r.d = d;
  r.mul = (uint32_t)(((1ull << (31 + s)) / d) + 1ull);
  r.sh = 31 + s;
  return r;
}
__host__ __device__ __forceinline__ uint32_t ctp_div_u31(uint32_t x, CtpDiv v) {
  return (uint32_t)(((unsigned long long)x * v.mul) >> v.sh);
}
// ---- segment march (src/base_map.cpp:49-74 and :77-97) ----------------------------------
// Where the endpoints of edge e come from.
struct EdgeSrc {
  const double *a;      // pairs mode: a[e], b[e];  indexed mode: node table
  const double *b;
  const int32_t *from;  // indexed mode (may be NULL => from = e / k within a query)
  const int32_t *to;    // indexed mode: to[e] (local index, <0 = no neighbour)
  int64_t n;            // nodes per query (indexed)
  int k;                // neighbours per node when from == NULL
  int indexed;
  int to_stride;        // from == NULL: to[row * to_stride + slot]
  const float4 *order;  // from == NULL, optional: work item r*k + slot takes the row whose id is order[r].w
                        // (the k-NN stage's cell-sorted points), so that neighbouring work items march
                        // through neighbouring voxels; outputs stay indexed by the original edge number
  CtpDiv dk, dn;        // from == NULL: division by k and by n (valid when small != 0)
  int small;            // every edge index of the launch is below 2^31
  const int32_t *perm;  // optional: work item i takes edge / segment perm[i] (segments binned by the voxel block of
                        // their midpoint; roadmap edges the whole-edge filter left undecided); outputs stay indexed
                        // by the edge number
  const int32_t *m_dev; // optional: the number of work items is read from here (a list filled by an earlier kernel)
};

// work item -> edge number (identity unless S.order is set)
__device__ __forceinline__ int64_t ctp_edge_of_item(const EdgeSrc &S, int64_t item) {
  if (S.perm) return S.perm[item];
  if (!S.order) return item;
  int64_t r, qbase;
  if (S.small) {
    const uint32_t r32 = ctp_div_u31((uint32_t)item, S.dk);
    r = r32;
    qbase = (int64_t)(ctp_div_u31(r32, S.dn) * (uint32_t)S.n);
  } else {
    r = item / S.k;
    qbase = (r / S.n) * S.n;
  }
  const int64_t row = qbase + __float_as_int(S.order[r].w);
  return row * S.k + (item - r * S.k);
}

struct EdgeOut {
  uint8_t *free_out;
  double *hit;
  int32_t *hit_idx;
  int32_t *lookups;
  unsigned long long *total_lookups;  // device counter, one atomic per warp
};

// endpoints of edge e as (start, vector), in march order (src/base_map.cpp:52-56, :82,90); false = no such edge
// (missing neighbour)
__device__ __forceinline__ bool ctp_edge_ends(const EdgeSrc &S, int64_t e, int guards, double &ax, double &ay, double &az,
                                              double &vx, double &vy, double &vz) {
  const double *pa, *pb;
  bool missing = false;
  if (S.indexed) {
    int64_t fi, ti;
    int32_t t;
    int64_t row = 0, qbase = 0;
    if (S.from) {
      t = S.to[e];
    } else {  // batched PRM: e = (query*n + node)*k + slot; neighbour ids are local to the query
      if (S.small) {
        const uint32_t r32 = ctp_div_u31((uint32_t)e, S.dk);
        row = r32;
        qbase = (int64_t)(ctp_div_u31(r32, S.dn) * (uint32_t)S.n);
      } else {
        row = e / S.k;
        qbase = (row / S.n) * S.n;
      }
      t = S.to[row * S.to_stride + (e - row * S.k)];
    }
    if (S.from) {
      fi = S.from[e];
      ti = t;
    } else {
      fi = row;
      ti = qbase + t;
    }
    missing = t < 0;
    pa = S.a + 3 * fi;
    pb = S.a + 3 * (missing ? fi : ti);
  } else {
    pa = S.a + 3 * e;
    pb = S.b + 3 * e;
  }
  double x0 = pa[0], y0 = pa[1], z0 = pa[2];
  double x1 = pb[0], y1 = pb[1], z1 = pb[2];
  if (guards) {  // marches from `neigbour` toward `actual` (src/base_map.cpp:82,90)
    double t0 = x0; x0 = x1; x1 = t0;
    t0 = y0; y0 = y1; y1 = t0;
    t0 = z0; z0 = z1; z1 = t0;
  }
  vx = x1 - x0;
  vy = y1 - y0;
  vz = z1 - z0;
  ax = x0;
  ay = y0;
  az = z0;
  return !missing;
}

// sample count of an edge vector (src/base_map.cpp:57-63).
// returns 0 = march needed, 1 = free without lookups, 2 = not free without lookups (absurd length)
__device__ __forceinline__ int ctp_edge_count(double vx, double vy, double vz, int guards, double cdc, double &npts,
                                              int &n_last) {
  const double d = sqrt((vx * vx + vy * vy) + vz * vz);
  npts = ceil((d / cdc) + 1.0);
  n_last = 0;
  // d < cdc shortcut exists only in ...BetweenNodes (src/base_map.cpp:57-59).
  // NaN npts: `index < NaN` is false, the reference loop body never runs => free.
  const bool trivially_free = (!guards && d < cdc) || !(npts == npts) || npts < 2.0;
  const bool absurd = npts > 2147483647.0;  // reference would spin on int overflow
  if (absurd) return 2;
  if (trivially_free) return 1;
  n_last = (int)npts - 1;  // indices 1 .. npts-1 (src/base_map.cpp:64)
  return 0;
}

// endpoints + sample count of edge e (src/base_map.cpp:52-63).
// returns 0 = march needed, 1 = free without lookups, 2 = not free without lookups (missing / absurd)
__device__ __forceinline__ int ctp_arm_edge(const EdgeSrc &S, int64_t e, int guards, double cdc, double &ax,
                                            double &ay, double &az, double &vx, double &vy, double &vz,
                                            double &npts, int &n_last) {
  const bool present = ctp_edge_ends(S, e, guards, ax, ay, az, vx, vy, vz);
  const int st = ctp_edge_count(vx, vy, vz, guards, cdc, npts, n_last);
  if (!present) {
    n_last = 0;
    return 2;
  }
  return st;
}

// One group of G lanes marches one segment, G consecutive samples per step, in the
// reference's sample order; a warp-wide ballot gives each group its first colliding lane, so
// the verdict, the first-hit index and the count of lookups "up to and including the first
// hit" are exactly the sequential ones.  Groups run a persistent strided loop over edges and
// re-arm independently, so lanes stay busy when neighbouring edges differ in length or
// exit early.
template <int L, int G>
__global__ void __launch_bounds__(256) k_edge_march(MapDev M, EdgeSrc S, int64_t m, double clr,
                                                    double cdc, int guards, EdgeOut O) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int sub = lane % G;
  const int gshift = lane - sub;
  const unsigned gmask = (G == 32) ? full : ((1u << G) - 1u);
  const int64_t n_groups = ((int64_t)gridDim.x * blockDim.x) / G;
  int64_t item = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G - n_groups;
  int64_t e = 0;  // the segment the group is marching (= item unless the launch is permuted)

  bool active = false;
  double ax = 0, ay = 0, az = 0, vx = 0, vy = 0, vz = 0, npts = 1.0;
  int idx0 = 1, n_last = 0;
  unsigned long long my_lookups = 0;  // only meaningful on sub == 0

  for (;;) {
    if (!active) {
      for (;;) {
        item += n_groups;
        if (item >= m) break;
        e = S.perm ? (int64_t)S.perm[item] : item;
        const int state = ctp_arm_edge(S, e, guards, cdc, ax, ay, az, vx, vy, vz, npts, n_last);
        if (state != 0) {
          if (sub == 0) {
            O.free_out[e] = state == 1 ? 1 : 0;
            if (O.hit) { O.hit[3 * e] = ctp_nan(); O.hit[3 * e + 1] = ctp_nan(); O.hit[3 * e + 2] = ctp_nan(); }
            if (O.hit_idx) O.hit_idx[e] = state == 1 ? -1 : -2;
            if (O.lookups) O.lookups[e] = 0;
          }
          continue;
        }
        idx0 = 1;
        active = true;
        break;
      }
    }
    if (!__any_sync(full, active)) break;

    const int idx = idx0 + sub;
    bool coll = false;
    double px = 0, py = 0, pz = 0;
    if (active && idx <= n_last) {
      const double t = (double)idx / npts;  // IEEE division, as (double)index / num_points_path
      px = ax + vx * t;
      py = ay + vy * t;
      pz = az + vz * t;
      coll = ctp_collides(esdf_clearance<L>(M, px, py, pz), clr);
    }
    const unsigned gb = (__ballot_sync(full, coll) >> gshift) & gmask;
    if (active) {
      if (gb) {
        const int first = __ffs(gb) - 1;
        if (sub == first && O.hit) { O.hit[3 * e] = px; O.hit[3 * e + 1] = py; O.hit[3 * e + 2] = pz; }
        if (sub == 0) {
          O.free_out[e] = 0;
          if (O.hit_idx) O.hit_idx[e] = idx0 + first;
          if (O.lookups) O.lookups[e] = idx0 + first;
          my_lookups += (unsigned)(idx0 + first);
        }
        active = false;
      } else {
        idx0 += G;
        if (idx0 > n_last) {
          if (sub == 0) {
            O.free_out[e] = 1;
            if (O.hit) { O.hit[3 * e] = ctp_nan(); O.hit[3 * e + 1] = ctp_nan(); O.hit[3 * e + 2] = ctp_nan(); }
            if (O.hit_idx) O.hit_idx[e] = -1;
            if (O.lookups) O.lookups[e] = n_last;
            my_lookups += (unsigned)n_last;
          }
          active = false;
        }
      }
    }
  }
  if (O.total_lookups) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_lookups += __shfl_xor_sync(full, my_lookups, o);
    if (lane == 0 && my_lookups) atomicAdd(O.total_lookups, my_lookups);
  }
}

// ---- flat (lookup-parallel) segment check -------------------------------------------------
// For the PRM build the n*k candidate edges are short (a handful of samples each) and almost all
// free, so lanes-per-edge marching wastes most issue slots on per-edge setup and idle lanes.  Here
// a warp takes 32 edges, one per lane for the setup (endpoints, length, sample count), then lays
// the samples of all 32 edges out back to back and works through them 32 at a time: every lane
// evaluates one ESDF lookup per step regardless of which edge it belongs to.  A colliding sample
// records its march index with a shared-memory atomicMin, so the verdict, the first-hit index and
// the reference-order lookup count are exactly the sequential ones.  Long edges are processed in
// rounds of CTP_FLAT_R samples per edge; an edge stops at the first round that found a hit.
// (double)idx / npts for the small integers of a segment march, without the division sequence:
// with r = 1/npts correctly rounded, q0 = idx*r is within an ulp or so of the quotient, the
// remainder idx - q0*npts is exact in one fma, and the corrected q0 + rem*r rounds to the correctly
// rounded quotient.  Verified exhaustively for every 1 <= idx < npts <= 65536 by
// tests/csrc/div_identity.c; longer marches take the true division.
#define CTP_DIV_SMALL_LIMIT 65536.0
__device__ __forceinline__ double ctp_div_small(double a, double b, double rb) {
  const double q0 = a * rb;  // caller guarantees b <= CTP_DIV_SMALL_LIMIT
  return __fma_rn(__fma_rn(-q0, b, a), rb, q0);
}

// ---- float32 whole-edge filter for roadmap edges ----------------------------------------------------
// Nearly all k-NN edges of a roadmap are free, and most of them run through space whose clearance is far above the
// threshold.  k_edge_mid decides those from ONE lookup: the trilinear interpolant f is continuous and, inside a cell,
// |df/dq_axis| <= G (G = the largest difference of adjacent voxels of the grid, k_map_stats), so along the march
// |f(q_j) - f(q_i)| <= G * |q_j - q_i|_1 = G * |qv|_1 * |j - i| / n  (q in voxel units, qv = the edge vector).  When the
// clearance of the MIDDLE sample exceeds the threshold by more than that bound for the farthest sample, every sample of
// the edge is free -- provided none of them sits exactly on a lattice plane (the reference's 0/0 -> NaN -> collision
// quirk on integral index coordinates) and all lie in cells 1 .. N-2 (no range / position test can fire).  Everything
// is evaluated in float32 with RIGOROUS bounds on the difference to the reference's FP64 values; an edge that cannot be
// decided with certainty goes, by its number, onto a list that the exact kernel (k_edge_flat) then marches.  Verdicts,
// and the reference-order lookup counts, are therefore bit-identical to the all-FP64 path (tests/test_gpu_filter.py
// runs both against the oracle on inputs built to sit on those bounds).  The same idea as floating-point filters for
// exact geometric predicates.
//
// Per node (k_node_qtable, FP64 once per node instead of once per edge): qa = w2i(a), cell Ia = floor(qa) (10 bits per
// axis) and fraction fa = float(qa - Ia) (|error| <= 2^-25).  Per edge: qv = float(Ib - Ia) + (fb - fa), off the true
// qb - qa by at most 2^-24 * (|qv| + 2) per axis.
//  * sample count: x = |qv| * (voxel / cdc) + 1 in float32 is within tol = 2^-24 * (6 x + 2 (voxel/cdc) (Qv + 2)) of
//    d / cdc + 1; npts = ceil(x) and the d < cdc shortcut are taken only when x is farther than tol from an integer.
//  * index coordinate of sample i relative to Ia: r = fma(qv, fl(i * rcp(n)), fa) with
//    |r - (q_ref(i) - Ia)| <= 2^-24 * (4.01 |qv| + 3.5) + eta < delta, eta = the FP64 roundings of the reference's own
//    transform against the affine form (host-computed from centroid, extent, N: MapDev::filt_eta).
//  * lattice planes: every |r - rint(r)| > delta  =>  no index coordinate of the sample is integral, its cell is certain.
//  * clearance at the middle sample: |f32 - f_ref| <= 3 G delta + 3 * 2^-24 (G + Vmax) < filt_eps (computed for
//    delta_cap = delta at CTP_FILT_QV_MAX voxels per axis -- longer edges are left to the exact kernel -- times 1.5).
#define CTP_FILT_QV_MAX 256.0f
#define CTP_QT_INVALID 0x80000000u

// cell and fraction of every node's index coordinate; .w = x | y << 10 | z << 20, or CTP_QT_INVALID when a coordinate
// is NaN or outside [0, 1024)
__global__ void __launch_bounds__(256) k_node_qtable(MapDev M, const double *__restrict__ nodes, int64_t total,
                                                     float4 *__restrict__ qt) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double q0 = ctp_w2i(nodes[3 * i], M.cx, M.s1, M.s2), q1 = ctp_w2i(nodes[3 * i + 1], M.cy, M.s1, M.s2),
               q2 = ctp_w2i(nodes[3 * i + 2], M.cz, M.s1, M.s2);
  const bool ok = (q0 >= 0.0) & (q0 < 1024.0) & (q1 >= 0.0) & (q1 < 1024.0) & (q2 >= 0.0) & (q2 < 1024.0);
  const double f0 = floor(q0), f1 = floor(q1), f2 = floor(q2);
  const unsigned pack = ok ? ((unsigned)(int)f0 | ((unsigned)(int)f1 << 10) | ((unsigned)(int)f2 << 20)) : CTP_QT_INVALID;
  qt[i] = make_float4((float)(q0 - f0), (float)(q1 - f1), (float)(q2 - f2), __uint_as_float(pack));
}

template <int L>
__device__ __forceinline__ float ctp_filter_value(const MapDev &M, float xd, float yd, float zd, int x0, int y0, int z0) {
  float v[8];
  Corners<L>::load(M, x0, y0, z0, v);
  const float c00 = __fmaf_rn(xd, __fsub_rn(v[4], v[0]), v[0]);
  const float c01 = __fmaf_rn(xd, __fsub_rn(v[5], v[1]), v[1]);
  const float c10 = __fmaf_rn(xd, __fsub_rn(v[6], v[2]), v[2]);
  const float c11 = __fmaf_rn(xd, __fsub_rn(v[7], v[3]), v[3]);
  const float c0 = __fmaf_rn(yd, __fsub_rn(c10, c00), c00);
  const float c1 = __fmaf_rn(yd, __fsub_rn(c11, c01), c01);
  return __fmaf_rn(zd, __fsub_rn(c1, c0), c0);
}

struct MidPar {
  float hc;      // voxel size / collision_distance_check, rounded to nearest
  float f_need;  // min_clearance + 2 * filt_eps, rounded up
  float lip_g;   // G * 1.001, rounded up
  float eta;     // filt_eta, rounded up
};

// Roadmap edges (batched PRM mode of EdgeSrc: from == NULL, neighbour table `to`; fewer than 2^31 edges): work item ->
// edge as in ctp_edge_of_item.  Decided edges get their verdict (free) here; the numbers of all others are appended
// to `list`.  A thread takes CTP_MID_IPT items a block apart and runs them phase by phase, so that the dependent
// loads of the items (row order -> neighbour id -> node table -> grid) are in flight together.
//  * lattice planes, all samples: index coordinates in 32-bit fixed point, R_i = FA + i * S mod 1 with
//    S = fl(qv * rcp(n)) truncated to 2^-32: |R_i - frac(q_ref(i))| <= 2^-24 (3.01 |qv| + 2.5) + i * 2^-32 + eta < delta
//    for i < 1000; a sample is off every lattice plane when (R + D) mod 1 >= 2 D on all three axes, D = delta.
// stats[1] += decided edges, stats[2] += edges seen, stats[3] += reference-order lookups of the decided edges (which
// O.total_lookups receives as well).
#ifndef CTP_MID_THREADS
#define CTP_MID_THREADS 256
#endif
#ifndef CTP_MID_IPT
#define CTP_MID_IPT 2
#endif
#ifndef CTP_MID_MINB
#define CTP_MID_MINB 5  // 48 registers: measured 1.74 ms per C2 step against 1.81 at 4 CTAs per SM (56 registers)
#endif
template <int L>
__global__ void __launch_bounds__(CTP_MID_THREADS, CTP_MID_MINB) k_edge_mid(MapDev M, EdgeSrc S, uint32_t m, MidPar P,
                                                              const float4 *__restrict__ qt, EdgeOut O,
                                                              int32_t *__restrict__ list, int32_t *__restrict__ list_cnt,
                                                              unsigned long long *__restrict__ stats) {
  __shared__ int s_push[CTP_MID_IPT][CTP_MID_THREADS / 32];
  __shared__ int s_base;
  __shared__ unsigned s_lk, s_dec;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t tile0 = blockIdx.x * (uint32_t)(CTP_MID_THREADS * CTP_MID_IPT);
  if (tile0 >= m) return;  // (uniform per CTA)
  if (threadIdx.x == 0) { s_lk = 0; s_dec = 0; }
  uint32_t e[CTP_MID_IPT], row[CTP_MID_IPT], nb[CTP_MID_IPT];
  bool valid[CTP_MID_IPT], sure[CTP_MID_IPT];
  // phase 1: work item -> (row, slot), neighbour id
#pragma unroll
  for (int h = 0; h < CTP_MID_IPT; h++) {
    const uint32_t item = tile0 + (uint32_t)h * CTP_MID_THREADS + threadIdx.x;
    valid[h] = item < m;
    const uint32_t it = valid[h] ? item : 0u;
    const uint32_t r32 = ctp_div_u31(it, S.dk);
    const uint32_t slot = it - r32 * (uint32_t)S.k;
    const uint32_t qbase = ctp_div_u31(r32, S.dn) * (uint32_t)S.n;
    row[h] = S.order ? qbase + (uint32_t)__float_as_int(S.order[r32].w) : r32;
    e[h] = row[h] * (uint32_t)S.k + slot;
    const int32_t t = S.to[(size_t)row[h] * S.to_stride + slot];
    sure[h] = valid[h] & (t >= 0);
    nb[h] = sure[h] ? qbase + (uint32_t)t : row[h];
  }
  // phase 2: node table
  float4 A[CTP_MID_IPT], B[CTP_MID_IPT];
#pragma unroll
  for (int h = 0; h < CTP_MID_IPT; h++) {
    A[h] = qt[row[h]];
    B[h] = qt[nb[h]];
  }
  // phase 3: edge vector, sample count, the middle sample's cell and fractions
  float vx[CTP_MID_IPT], vy[CTP_MID_IPT], vz[CTP_MID_IPT], inv_n[CTP_MID_IPT], delta[CTP_MID_IPT];
  float xd[CTP_MID_IPT], yd[CTP_MID_IPT], zd[CTP_MID_IPT], step[CTP_MID_IPT];
  int cx[CTP_MID_IPT], cy[CTP_MID_IPT], cz[CTP_MID_IPT], n_last[CTP_MID_IPT];
  bool trivial[CTP_MID_IPT];
#pragma unroll
  for (int h = 0; h < CTP_MID_IPT; h++) {
    const unsigned pa = __float_as_uint(A[h].w), pb = __float_as_uint(B[h].w);
    const int ax = pa & 1023, ay = (pa >> 10) & 1023, az = (pa >> 20) & 1023;
    const int bx = pb & 1023, by = (pb >> 10) & 1023, bz = (pb >> 20) & 1023;
    // both ends in cells 1 .. N-2 (exact cells): so is every sample between them
    const int lo = min(min(min(ax, ay), az), min(min(bx, by), bz)), hi = max(max(max(ax, ay), az), max(max(bx, by), bz));
    bool ok = sure[h] & (((pa | pb) & CTP_QT_INVALID) == 0) & (lo >= 1) & (hi <= M.n - 2);
    vx[h] = __fadd_rn((float)(bx - ax), __fsub_rn(B[h].x, A[h].x));
    vy[h] = __fadd_rn((float)(by - ay), __fsub_rn(B[h].y, A[h].y));
    vz[h] = __fadd_rn((float)(bz - az), __fsub_rn(B[h].z, A[h].z));
    const float Qv = fmaxf(fmaxf(fabsf(vx[h]), fabsf(vy[h])), fabsf(vz[h]));
    const float l1 = __fadd_ru(__fadd_ru(fabsf(vx[h]), fabsf(vy[h])), fabsf(vz[h]));
    const float d2 = __fmaf_rn(vz[h], vz[h], __fmaf_rn(vy[h], vy[h], __fmul_rn(vx[h], vx[h])));
    const float x = __fmaf_rn(sqrtf(d2), P.hc, 1.0f);
    const float tol = __fmul_rn(5.9604645e-8f, __fmaf_rn(6.0f, x, __fmul_rn(__fmul_rn(2.0f, P.hc), __fadd_rn(Qv, 2.0f))));
    const float up = ceilf(x);
    ok = ok & (Qv <= CTP_FILT_QV_MAX) & (x >= 1.0f) & (x < 1000.0f);
    ok = ok & (__fsub_rn(up, x) > tol) & (__fsub_rn(x, __fsub_rn(up, 1.0f)) > tol);  // ceil, and d < cdc at x = 2, certain
    trivial[h] = ok & (x < 2.0f);  // d < cdc: free without lookups (src/base_map.cpp:57-59)
    n_last[h] = (ok && !trivial[h]) ? (int)up - 1 : 0;
    inv_n[h] = __frcp_rn(fmaxf(up, 1.0f));
    delta[h] = __fadd_ru(__fmul_ru(__fmaf_rn(4.5f, Qv, 5.0f), 5.9610e-8f), __fadd_ru(P.eta, 3.0e-7f));
    const int mid = (n_last[h] + 1) >> 1;
    const float tm = __fmul_rn((float)mid, inv_n[h]);
    const float rx = __fmaf_rn(vx[h], tm, A[h].x), ry = __fmaf_rn(vy[h], tm, A[h].y), rz = __fmaf_rn(vz[h], tm, A[h].z);
    const float flx = floorf(rx), fly = floorf(ry), flz = floorf(rz);
    xd[h] = __fsub_rn(rx, flx); yd[h] = __fsub_rn(ry, fly); zd[h] = __fsub_rn(rz, flz);
    const float mx = fminf(xd[h], __fsub_rn(1.f, xd[h])), my = fminf(yd[h], __fsub_rn(1.f, yd[h])),
                mz = fminf(zd[h], __fsub_rn(1.f, zd[h]));
    ok = ok & (fminf(fminf(mx, my), mz) > delta[h]);
    // an undecidable item still fetches (its loads are in flight with the others'): keep its cell inside the grid
    const int hi_cell = M.n - 2;
    cx[h] = ok ? ax + (int)flx : 1; cy[h] = ok ? ay + (int)fly : 1; cz[h] = ok ? az + (int)flz : 1;
    cx[h] = min(max(cx[h], 0), hi_cell); cy[h] = min(max(cy[h], 0), hi_cell); cz[h] = min(max(cz[h], 0), hi_cell);
    // bound on the clearance lost per sample step: G * |qv|_1 / n, every rounding upward
    step[h] = fmaxf(__fmul_ru(__fmul_ru(P.lip_g, l1), __fmul_ru(inv_n[h], 1.000001f)), 1e-30f);
    sure[h] = ok;
  }
  // phase 4: the lookups
  float f[CTP_MID_IPT];
#pragma unroll
  for (int h = 0; h < CTP_MID_IPT; h++) f[h] = ctp_filter_value<L>(M, xd[h], yd[h], zd[h], cx[h], cy[h], cz[h]);
  // phase 5: verdicts; no sample of a free edge may sit on a lattice plane
  bool decided[CTP_MID_IPT];
  unsigned lk = 0, dec = 0;
#pragma unroll
  for (int h = 0; h < CTP_MID_IPT; h++) {
    // the middle probe proves the samples mid - s .. mid + s free, s = floor((f - f_need) / step)
    bool all_free = false;
    if (sure[h] & !trivial[h]) {
      const float a0 = __fsub_rd(f[h], P.f_need);
      const int mid = (n_last[h] + 1) >> 1;
      const int far = max(mid - 1, n_last[h] - mid);
      all_free = a0 >= __fmul_ru(step[h], (float)far);
    }
    if (all_free) {
      const float sc = 4294967296.0f;
      const unsigned D = __float2uint_ru(__fmul_ru(delta[h], sc)), W = 2u * D + 1u;
      const unsigned Sx = (unsigned)__float2ll_rz(__fmul_rn(__fmul_rn(vx[h], inv_n[h]), sc)),
                     Sy = (unsigned)__float2ll_rz(__fmul_rn(__fmul_rn(vy[h], inv_n[h]), sc)),
                     Sz = (unsigned)__float2ll_rz(__fmul_rn(__fmul_rn(vz[h], inv_n[h]), sc));
      unsigned Rx = (unsigned)__float2ll_rz(__fmul_rn(A[h].x, sc)) + D, Ry = (unsigned)__float2ll_rz(__fmul_rn(A[h].y, sc)) + D,
               Rz = (unsigned)__float2ll_rz(__fmul_rn(A[h].z, sc)) + D;
      // four samples per trip; the trips run past n_last (up to three positions beyond the far end): testing more
      // positions than needed can only leave an edge to the exact kernel
      bool near = false;
      for (int idx = 1; idx <= n_last[h]; idx += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          Rx += Sx; Ry += Sy; Rz += Sz;
          near |= (Rx < W) | (Ry < W) | (Rz < W);
        }
      }
      all_free = !near;
    }
    decided[h] = (sure[h] & trivial[h]) | all_free;
    if (decided[h]) {
      O.free_out[e[h]] = 1;
      if (O.hit) { O.hit[3 * (size_t)e[h]] = ctp_nan(); O.hit[3 * (size_t)e[h] + 1] = ctp_nan(); O.hit[3 * (size_t)e[h] + 2] = ctp_nan(); }
      if (O.hit_idx) O.hit_idx[e[h]] = -1;
      if (O.lookups) O.lookups[e[h]] = n_last[h];
      lk += (unsigned)n_last[h];
      dec++;
    }
  }
  // undecided edges -> list: one atomic per CTA; counters likewise
  unsigned pm[CTP_MID_IPT];
#pragma unroll
  for (int h = 0; h < CTP_MID_IPT; h++) {
    pm[h] = __ballot_sync(full, valid[h] && !decided[h]);
    if (lane == 0) s_push[h][wid] = __popc(pm[h]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lk += __shfl_xor_sync(full, lk, o);
    dec += __shfl_xor_sync(full, dec, o);
  }
  __syncthreads();
  if (lane == 0) {
    if (lk) atomicAdd(&s_lk, lk);
    if (dec) atomicAdd(&s_dec, dec);
  }
  if (threadIdx.x == 0) {
    int tot = 0;
#pragma unroll
    for (int h = 0; h < CTP_MID_IPT; h++)
      for (int w2 = 0; w2 < CTP_MID_THREADS / 32; w2++) {
        const int c = s_push[h][w2];
        s_push[h][w2] = tot;
        tot += c;
      }
    s_base = tot ? atomicAdd(list_cnt, tot) : 0;
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < CTP_MID_IPT; h++)
    if (valid[h] && !decided[h]) list[s_base + s_push[h][wid] + __popc(pm[h] & ((1u << lane) - 1u))] = (int32_t)e[h];
  if (threadIdx.x == 0) {
    if (O.total_lookups && s_lk) atomicAdd(O.total_lookups, (unsigned long long)s_lk);
    if (stats) {
      if (s_lk) atomicAdd(stats + 3, (unsigned long long)s_lk);
      if (s_dec) atomicAdd(stats + 1, (unsigned long long)s_dec);
      const uint32_t seen = min(m - min(tile0, m), (uint32_t)(CTP_MID_THREADS * CTP_MID_IPT));
      atomicAdd(stats + 2, (unsigned long long)seen);
    }
  }
}

#define CTP_FLAT_R 16
#ifndef CTP_FLAT_R0
#define CTP_FLAT_R0 2  // first round of the growing schedule (2^24 random segments, C3 / C2 map: 2 -> 2.94 / 2.11 ms, 4 -> 3.18 / 2.16, 8 -> 3.72 / 2.30, one round of 16 -> 4.60 / 2.52)
#endif
#define CTP_FLAT_WARPS 8

// (64 edges per warp, two per lane: the sample loop then runs ~11 full steps instead of ~5.4 with a
// partly empty last one)
#define CTP_FLAT_EPW 64
// GROW: the rounds of an edge take CTP_FLAT_R0 = 2, 4, 8, then CTP_FLAT_R samples.  The PRM's k-NN edges (a handful of samples, nearly
// all free) want everything in one round; arbitrary segments collide early and often, and every sample of a round
// past the first hit is a wasted lookup.
template <int L, bool GROW>
__global__ void __launch_bounds__(32 * CTP_FLAT_WARPS) k_edge_flat(MapDev M, EdgeSrc S, int64_t m, double clr,
                                                                   double cdc, int guards, EdgeOut O) {
  __shared__ double s_par[CTP_FLAT_WARPS][8][CTP_FLAT_EPW];
  __shared__ int s_first[CTP_FLAT_WARPS][CTP_FLAT_EPW];
  __shared__ int2 s_ob[CTP_FLAT_WARPS][CTP_FLAT_EPW];
  __shared__ uint8_t s_slot[CTP_FLAT_WARPS][CTP_FLAT_EPW * CTP_FLAT_R];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double(*par)[CTP_FLAT_EPW] = s_par[w];
  int *first = s_first[w];
  int2 *ob = s_ob[w];
  uint8_t *slot = s_slot[w];
  const int64_t n_warps = (int64_t)gridDim.x * CTP_FLAT_WARPS;
  if (S.m_dev) m = *S.m_dev;  // a list filled by an earlier kernel on the same stream
  const int64_t n_chunks = (m + CTP_FLAT_EPW - 1) / CTP_FLAT_EPW;
  unsigned long long my_lookups = 0;

  for (int64_t chunk = (int64_t)blockIdx.x * CTP_FLAT_WARPS + w; chunk < n_chunks; chunk += n_warps) {
    // lane owns edges `lane` and `lane + 32` of the chunk
    int64_t e[2];
    int n_last[2], state[2], base[2];
    bool alive[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int el = lane + 32 * h;
      const int64_t item = chunk * CTP_FLAT_EPW + el;
      e[h] = item < m ? ctp_edge_of_item(S, item) : item;
      double ax = 0, ay = 0, az = 0, vx = 0, vy = 0, vz = 0, npts = 1.0;
      n_last[h] = 0;
      state[h] = 3;  // 3 = no edge
      if (item < m) state[h] = ctp_arm_edge(S, e[h], guards, cdc, ax, ay, az, vx, vy, vz, npts, n_last[h]);
      par[0][el] = ax; par[1][el] = ay; par[2][el] = az;
      par[3][el] = vx; par[4][el] = vy; par[5][el] = vz;
      par[6][el] = npts;
      par[7][el] = 1.0 / npts;
      first[el] = 0x7fffffff;
      alive[h] = state[h] == 0;
      base[h] = 1;
      if (alive[h] && npts > CTP_DIV_SMALL_LIMIT) {
        // a march of more than 65536 samples (kilometres at the usual spacing) is outside the range the
        // division shortcut is verified for: the owning lane walks it alone with the true division
        alive[h] = false;
        for (int idx = 1; idx <= n_last[h]; idx++) {
          const double tt = (double)idx / npts;
          if (ctp_collides(esdf_clearance<L>(M, ax + vx * tt, ay + vy * tt, az + vz * tt), clr)) {
            first[el] = idx;
            break;
          }
        }
      }
    }
    int round_cap = GROW ? CTP_FLAT_R0 : CTP_FLAT_R;
    for (;;) {
      int c[2], incl[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        c[h] = alive[h] ? min(GROW ? round_cap : CTP_FLAT_R, n_last[h] - base[h] + 1) : 0;
        incl[h] = c[h];
      }
      if (GROW) round_cap = min(2 * round_cap, CTP_FLAT_R);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t0 = __shfl_up_sync(full, incl[0], o), t1 = __shfl_up_sync(full, incl[1], o);
        if (lane >= o) { incl[0] += t0; incl[1] += t1; }
      }
      const int tot0 = __shfl_sync(full, incl[0], 31);
      const int total = tot0 + __shfl_sync(full, incl[1], 31);
      if (total == 0) break;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int el = lane + 32 * h;
        const int off = (h ? tot0 : 0) + incl[h] - c[h];
        ob[el] = make_int2(off, base[h]);
        for (int j = 0; j < c[h]; j++) slot[off + j] = (uint8_t)el;
      }
      __syncwarp();
      for (int t = lane; t < total; t += 32) {
        const int el = slot[t];
        const int2 o2 = ob[el];
        const int idx = o2.y + (t - o2.x);
        const double tt = ctp_div_small((double)idx, par[6][el], par[7][el]);  // (double)index / num_points_path
        const double px = par[0][el] + par[3][el] * tt;
        const double py = par[1][el] + par[4][el] * tt;
        const double pz = par[2][el] + par[5][el] * tt;
        if (ctp_collides(esdf_clearance<L>(M, px, py, pz), clr)) atomicMin(&first[el], idx);
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; h++)
        if (alive[h]) {
          if (first[lane + 32 * h] != 0x7fffffff) alive[h] = false;
          else {
            base[h] += c[h];
            alive[h] = base[h] <= n_last[h];
          }
        }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      if (state[h] == 3) continue;
      const int el = lane + 32 * h;
      const int f = first[el];
      const bool blocked = f != 0x7fffffff;
      const bool fr = state[h] == 2 ? false : !blocked;
      O.free_out[e[h]] = fr ? 1 : 0;
      if (O.hit) {
        double hx = ctp_nan(), hy = ctp_nan(), hz = ctp_nan();
        if (blocked) {
          const double tt = (double)f / par[6][el];
          hx = par[0][el] + par[3][el] * tt;
          hy = par[1][el] + par[4][el] * tt;
          hz = par[2][el] + par[5][el] * tt;
        }
        O.hit[3 * e[h]] = hx; O.hit[3 * e[h] + 1] = hy; O.hit[3 * e[h] + 2] = hz;
      }
      if (O.hit_idx) O.hit_idx[e[h]] = blocked ? f : (state[h] == 2 ? -2 : -1);
      const int lk = blocked ? f : n_last[h];
      if (O.lookups) O.lookups[e[h]] = lk;
      my_lookups += (unsigned)lk;
    }
    __syncwarp();  // the next chunk reuses the warp's shared arrays
  }
  if (O.total_lookups) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_lookups += __shfl_xor_sync(full, my_lookups, o);
    if (lane == 0 && my_lookups) atomicAdd(O.total_lookups, my_lookups);
  }
}

// ---- statistics of the grid the filter's error bound needs ------------------------------------------------------
// out[0] = largest |difference of voxels adjacent along an axis| (float bits), out[1] = largest |voxel| (float bits),
// out[2] = number of non-finite voxels (saturating flag)
__global__ void __launch_bounds__(256) k_map_stats(const float *__restrict__ plain, int n, unsigned *__restrict__ out) {
  const size_t nn = (size_t)n * n, total = nn * n;
  float g = 0.f, vm = 0.f;
  bool bad = false;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int z = (int)(o % n), y = (int)((o / n) % n), x = (int)(o / nn);
    const float v = plain[o];
    bad |= !(fabsf(v) <= 3.4028234e38f);
    vm = fmaxf(vm, fabsf(v));
    if (z + 1 < n) g = fmaxf(g, fabsf(plain[o + 1] - v));
    if (y + 1 < n) g = fmaxf(g, fabsf(plain[o + n] - v));
    if (x + 1 < n) g = fmaxf(g, fabsf(plain[o + nn] - v));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, s));
    vm = fmaxf(vm, __shfl_xor_sync(0xffffffffu, vm, s));
  }
  const bool any_bad = __any_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, __float_as_uint(g));
    atomicMax(out + 1, __float_as_uint(vm));
    if (any_bad) atomicMax(out + 2, 1u);
  }
}

// ---- binning of generic segments ---------------------------------------------------------------
// Segments handed to ctp_edge_check come in no particular order; on a grid beyond the L2 (and the TLB reach) every
// lookup of every segment is then a DRAM access to a cold page.  Key = Morton code of the 16^3-voxel block that holds
// the segment's midpoint: after a counting sort on it the segments a warp (and its neighbours on the SM, and the SMs
// next to it) work on lie in one corner of the map.  Outputs are written by segment number, so the order of the
// segments inside a bin (decided by atomics) does not show.
#define CTP_BIN_SHIFT 4
__device__ __forceinline__ unsigned ctp_morton3(unsigned x, unsigned y, unsigned z) {
  unsigned key = 0;
#pragma unroll
  for (int b = 0; b < 8; b++) key |= ((x >> b) & 1u) << (3 * b) | ((y >> b) & 1u) << (3 * b + 1) | ((z >> b) & 1u) << (3 * b + 2);
  return key;
}
__device__ __forceinline__ unsigned ctp_segment_bin(const MapDev &M, const double *a, const double *b, unsigned nbins_axis) {
  unsigned c[3];
  const double cen[3] = {M.cx, M.cy, M.cz};
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    const double mid = 0.5 * (a[ax] + b[ax]);
    const double qd = ((mid - cen[ax]) * M.s1 + 1.0) * M.s2;  // the index transform of getClearence, any rounding will do
    int v = qd > 0.0 ? (qd < M.nd ? (int)qd : M.n - 1) : 0;   // NaN -> 0
    v >>= CTP_BIN_SHIFT;
    c[ax] = (unsigned)(v < (int)nbins_axis ? v : (int)nbins_axis - 1);
  }
  return ctp_morton3(c[0], c[1], c[2]);
}
__global__ void __launch_bounds__(256) k_seg_bin_count(MapDev M, const double *__restrict__ a, const double *__restrict__ b,
                                                       int64_t cnt, unsigned nbins_axis, unsigned *__restrict__ key_of,
                                                       int32_t *__restrict__ bin_cnt) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  const unsigned key = ctp_segment_bin(M, a + 3 * i, b + 3 * i, nbins_axis);
  key_of[i] = key;
  atomicAdd(bin_cnt + key, 1);
}
__global__ void __launch_bounds__(256) k_seg_bin_scatter(const unsigned *__restrict__ key_of, int64_t cnt,
                                                         const int32_t *__restrict__ bin_start, int32_t *__restrict__ cursor,
                                                         int32_t *__restrict__ perm) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  const unsigned key = key_of[i];
  perm[bin_start[key] + atomicAdd(cursor + key, 1)] = (int32_t)i;
}

// ---- layout builders ----------------------------------------------------------------------
__global__ void k_build_cell8(const float *__restrict__ plain, int n, float *__restrict__ cell8) {
  const size_t n1 = (size_t)(n - 1);
  const size_t cells = n1 * n1 * n1;
  for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < cells;
       c += (size_t)gridDim.x * blockDim.x) {
    const size_t z = c % n1, y = (c / n1) % n1, x = c / (n1 * n1);
    const float *p = plain + (x * n + y) * n + z;
    const size_t nn = (size_t)n * n;
    float4 a = make_float4(p[0], p[1], p[n], p[n + 1]);
    float4 b = make_float4(p[nn], p[nn + 1], p[nn + n], p[nn + n + 1]);
    reinterpret_cast<float4 *>(cell8)[2 * c] = a;
    reinterpret_cast<float4 *>(cell8)[2 * c + 1] = b;
  }
}

__global__ void k_build_brick(const float *__restrict__ plain, int n, int nb, float *__restrict__ brick) {
  const size_t total = (size_t)nb * nb * nb * 64;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total;
       o += (size_t)gridDim.x * blockDim.x) {
    const size_t b = o >> 6;
    const int in = (int)(o & 63);
    const int bz = (int)(b % nb), by = (int)((b / nb) % nb), bx = (int)(b / ((size_t)nb * nb));
    const int x = bx * 4 + (in >> 4), y = by * 4 + ((in >> 2) & 3), z = bz * 4 + (in & 3);
    brick[o] = (x < n && y < n && z < n) ? plain[((size_t)x * n + y) * n + z] : 0.0f;
  }
}

__global__ void k_narrow_f64(const double *__restrict__ in, size_t cnt, float *__restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cnt;
       i += (size_t)gridDim.x * blockDim.x)
    out[i] = (float)in[i];  // fill_data<double>: map_data is float (include/esdf_map.hpp:43)
}

// ---- procedural ESDF on the device ------------------------------------
// Signed distance to a set of vertical cylinders, min_i(sqrt(dx^2 + dy^2) - r_i), constant along z, written
// in the .npy order of the reference grid.  Voxel i of an axis sits at (c - E/2) + i * (E/N), the inverse of
// the world->index map of src/esdf_map.cpp:136-139.  FP64 with one rounding per operation, narrowed to f32
// at the end: bit-identical to ctopprm_learn_b200/synth.py:columns_field.
__global__ void __launch_bounds__(256) k_gen_columns(int n, double x_lo, double y_lo, double step,
                                                     const double *__restrict__ cx, const double *__restrict__ cy,
                                                     const double *__restrict__ r, int n_col, float *__restrict__ slab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * n) return;
  const int xi = t / n, yi = t - xi * n;
  const double x = x_lo + (double)xi * step, y = y_lo + (double)yi * step;
  double best = __longlong_as_double(0x7ff0000000000000LL);
  for (int i = 0; i < n_col; i++) {
    const double dx = x - cx[i], dy = y - cy[i];
    const double d = sqrt(dx * dx + dy * dy) - r[i];
    best = d < best ? d : best;  // np.minimum keeps the first argument on ties and propagates NaN the same way
  }
  slab[t] = (float)best;
}

__global__ void __launch_bounds__(256) k_broadcast_z(const float *__restrict__ slab, int n, float *__restrict__ vox) {
  const size_t total = (size_t)n * n * n;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x)
    vox[o] = slab[o / n];
}

