// ctp_paths.cuh -- candidate sampling, arc-length resampling, polyline deformability and
// path shortening (src/base_map.cpp:29-43,105-192; src/common.cpp:256-312;
// include/topological_prm_clustering.hpp:2313-2479).
#pragma once
#include "ctp_device.cuh"

__global__ void k_fill_i32(int32_t *p, int64_t cnt, int32_t v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

// ---------------- Philox4x32-10 counter-based generator (Salmon et al., SC'11) -------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// BaseMap::fillRandomStateInEllipse (src/base_map.cpp:29-43) -> random_ellipsoid_point
// (src/common.cpp:256-312): uniform point in the prolate spheroid with foci start/goal and
// major axis ratio*||goal-start||.  One uniform + three normal draws per point, like the
// reference; the draws come from Philox keyed by `seed`, counter = (candidate, query, block).
__global__ void __launch_bounds__(256)
k_sample_ellipsoid(const double *__restrict__ start, const double *__restrict__ goal, int64_t q, int64_t m,
                   double ratio, int planar, uint64_t seed, double *__restrict__ cand, double *__restrict__ draws = nullptr) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < q * m; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t qi = t / m, ci = t - qi * m;
    const double sx = start[3 * qi], sy = start[3 * qi + 1], sz = start[3 * qi + 2];
    const double gx = goal[3 * qi], gy = goal[3 * qi + 1], gz = goal[3 * qi + 2];
    double abx = gx - sx, aby = gy - sy, abz = gz - sz;
    const double two_focal = sqrt((abx * abx + aby * aby) + abz * abz);
    const double a1[3] = {abx / two_focal, aby / two_focal, abz / two_focal};
    int kk = 0;
    if (fabs(a1[1]) < fabs(a1[kk])) kk = 1;
    if (fabs(a1[2]) < fabs(a1[kk])) kk = 2;
    const double ek[3] = {kk == 0 ? 1.0 : 0.0, kk == 1 ? 1.0 : 0.0, kk == 2 ? 1.0 : 0.0};
    double a2[3] = {a1[1] * ek[2] - a1[2] * ek[1], a1[2] * ek[0] - a1[0] * ek[2], a1[0] * ek[1] - a1[1] * ek[0]};
    const double n2 = sqrt((a2[0] * a2[0] + a2[1] * a2[1]) + a2[2] * a2[2]);
    a2[0] /= n2; a2[1] /= n2; a2[2] /= n2;
    const double a3[3] = {a1[1] * a2[2] - a1[2] * a2[1], a1[2] * a2[0] - a1[0] * a2[2], a1[0] * a2[1] - a1[1] * a2[0]};
    const double two_major = two_focal * ratio;
    const double axis = sqrt(two_major * two_major - two_focal * two_focal);
    uint32_t c0[4] = {(uint32_t)ci, (uint32_t)(ci >> 32), (uint32_t)qi, 0u};
    uint32_t c1[4] = {(uint32_t)ci, (uint32_t)(ci >> 32), (uint32_t)qi, 1u};
    philox4x32_10(c0, (uint32_t)seed, (uint32_t)(seed >> 32));
    philox4x32_10(c1, (uint32_t)seed, (uint32_t)(seed >> 32));
    // u in (0,1) from 53 bits; Box-Muller normals from 32-bit uniforms in (0,1)
    const double u = ((double)(((uint64_t)c0[0] << 21) ^ (uint64_t)(c0[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
    const double u1 = ((double)c0[2] + 0.5) * (1.0 / 4294967296.0), u2 = ((double)c0[3] + 0.5) * (1.0 / 4294967296.0);
    const double u3 = ((double)c1[0] + 0.5) * (1.0 / 4294967296.0), u4 = ((double)c1[1] + 0.5) * (1.0 / 4294967296.0);
    const double r1 = sqrt(-2.0 * log(u1)), r2 = sqrt(-2.0 * log(u3));
    double s1, cs1, s2, cs2;
    sincospi(2.0 * u2, &s1, &cs1);
    sincospi(2.0 * u4, &s2, &cs2);
    const double g0 = r1 * cs1, g1 = r1 * s1, g2 = r2 * cs2;
    (void)s2;
    const double gn = sqrt((g0 * g0 + g1 * g1) + g2 * g2);
    const double radii = cbrt(u);  // pow(u, 1/3) (src/common.cpp:296)
    const double b0 = (two_major / 2.0) * (g0 * radii / gn);
    const double b1 = (axis / 2.0) * (g1 * radii / gn);
    const double b2 = (axis / 2.0) * (g2 * radii / gn);
    double *o = cand + 3 * t;
    o[0] = ((a1[0] * b0 + a2[0] * b1) + a3[0] * b2) + (sx + gx) / 2.0;
    o[1] = ((a1[1] * b0 + a2[1] * b1) + a3[1] * b2) + (sy + gy) / 2.0;
    o[2] = planar ? sz : ((a1[2] * b0 + a2[2] * b1) + a3[2] * b2) + (sz + gz) / 2.0;
    if (draws) {  // the uniform and the three normal draws of this point (src/common.cpp:295-303)
      draws[4 * t] = u;
      draws[4 * t + 1] = g0;
      draws[4 * t + 2] = g1;
      draws[4 * t + 3] = g2;
    }
  }
}

// ---------------- BaseMap::samplePath (src/base_map.cpp:105-157), one thread ---------------
// The cursor advances at most ONE segment per sample, exactly as the reference does.
// stride/offset let the caller write the samples reversed (shorten_path, forward == false).
__device__ inline int dev_sample_path(const double *__restrict__ pts, int npts, double length_tot,
                                      double num_samples, double *__restrict__ out, bool reversed) {
  int cur = 0;
  double sx = pts[3] - pts[0], sy = pts[4] - pts[1], sz = pts[5] - pts[2];
  double cur_len = sqrt((sx * sx + sy * sy) + sz * sz);
  double cur_start = 0.0;
  const int K = (int)num_samples;
  int k = 0;
  for (double i = 0; i < num_samples; ++i, ++k) {
    double at = length_tot * (i / (num_samples - 1));
    if (at > cur_start + cur_len) {
      cur += 1;
      if (cur + 1 >= npts) {
        if (at - length_tot > 10e-6) return -1;  // reference: ERROR + exit(1)
        cur -= 1;
        at = cur_start + cur_len;
      } else {
        cur_start += cur_len;
        sx = pts[3 * (cur + 1)] - pts[3 * cur];
        sy = pts[3 * (cur + 1) + 1] - pts[3 * cur + 1];
        sz = pts[3 * (cur + 1) + 2] - pts[3 * cur + 2];
        cur_len = sqrt((sx * sx + sy * sy) + sz * sz);
      }
    }
    const double *p0 = pts + 3 * cur, *p1 = pts + 3 * (cur + 1);
    const double f = (at - cur_start) / cur_len;
    double *o = out + 3 * (size_t)(reversed ? K - 1 - k : k);
    o[0] = p0[0] + (p1[0] - p0[0]) * f;
    o[1] = p0[1] + (p1[1] - p0[1]) * f;
    o[2] = p0[2] + (p1[2] - p0[2]) * f;
  }
  return 0;
}

__global__ void k_sample_path(const double *__restrict__ pts, const int32_t *__restrict__ off,
                              const double *__restrict__ length, const double *__restrict__ num_samples,
                              int64_t count, double *__restrict__ out, const int64_t *__restrict__ out_off,
                              int32_t *__restrict__ status) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int npts = off[i + 1] - off[i];
  int rc = -1;
  if (npts >= 2) rc = dev_sample_path(pts + 3 * (size_t)off[i], npts, length[i], num_samples[i], out + 3 * out_off[i], false);
  status[i] = rc == 0 ? 0 : -5;
}

// ---------------- warp-cooperative isSimplePathFreeBetweenNodes (src/base_map.cpp:49-74) ----
// All 32 lanes of the calling warp take part; 32 consecutive samples per step, ballot gives
// the first colliding one.  Returns the verdict on every lane.
template <int L>
__device__ __forceinline__ bool warp_edge_nodes(const MapDev &M, double ax, double ay, double az, double bx,
                                                double by, double bz, double clr, double cdc, int lane,
                                                double hit[3]) {
  const unsigned full = 0xffffffffu;
  const double vx = bx - ax, vy = by - ay, vz = bz - az;
  const double d = sqrt((vx * vx + vy * vy) + vz * vz);
  hit[0] = hit[1] = hit[2] = ctp_nan();
  if (d < cdc) return true;
  const double npts = ceil((d / cdc) + 1.0);
  if (!(npts == npts)) return true;
  if (npts > 2147483647.0) return false;
  const int n_last = (int)npts - 1;
  for (int idx0 = 1; idx0 <= n_last; idx0 += 32) {
    const int idx = idx0 + lane;
    bool coll = false;
    double px = 0, py = 0, pz = 0;
    if (idx <= n_last) {
      const double t = (double)idx / npts;
      px = ax + vx * t;
      py = ay + vy * t;
      pz = az + vz * t;
      coll = ctp_collides(esdf_clearance<L>(M, px, py, pz), clr);
    }
    const unsigned b = __ballot_sync(full, coll);
    if (b) {
      const int first = __ffs(b) - 1;
      hit[0] = __shfl_sync(full, px, first);
      hit[1] = __shfl_sync(full, py, first);
      hit[2] = __shfl_sync(full, pz, first);
      return false;
    }
  }
  return true;
}

// ---------------- BaseMap::isDeformablePath (src/base_map.cpp:160-192) ----------------------
// One CTA per pair: two threads resample the two polylines, then the warps share the K
// straight connectors between equal-parameter samples; first blocked connector ends the pair.
#ifndef CTP_PATH_BLOCK
#define CTP_PATH_BLOCK 1024  // largest CTA of the polyline kernels (shared arrays are sized for it); the launch picks
                             // ctp_path_threads(): a handful of polylines / pairs (the cluster stage of one query) is bound
                             // by the latency of the sequential steps and takes 32 warps per CTA -- 32 speculative anchor
                             // tests / connectors in flight per step instead of 8 -- while a large batch fills the SMs
                             // with 8-warp CTAs (2016 pairs: 0.51 ms, against 0.89 ms with 32-warp CTAs)
#endif
template <int L>
__global__ void __launch_bounds__(CTP_PATH_BLOCK)
k_deformable(MapDev M, const double *__restrict__ pts, const int32_t *__restrict__ off,
             const double *__restrict__ length, const int32_t *__restrict__ ia, const int32_t *__restrict__ ib,
             const double *__restrict__ num_check, const int64_t *__restrict__ scratch_off,
             double *__restrict__ scratch, double clr, double cdc, uint8_t *__restrict__ out) {
  const int pair = blockIdx.x;
  const int a = ia[pair], b = ib[pair];
  const double nc = num_check[pair];
  const int K = nc > 0 ? (int)nc : 0;
  double *s1 = scratch + 3 * scratch_off[pair];
  double *s2 = s1 + 3 * (size_t)K;
  __shared__ int s_fail, s_blocked;
  if (threadIdx.x == 0) { s_fail = 0; s_blocked = 0; }
  __syncthreads();
  if (threadIdx.x == 0 || threadIdx.x == 32) {
    const int p = threadIdx.x == 0 ? a : b;
    const int npts = off[p + 1] - off[p];
    int rc = -1;
    if (npts >= 2) rc = dev_sample_path(pts + 3 * (size_t)off[p], npts, length[p], nc, threadIdx.x == 0 ? s1 : s2, false);
    if (rc != 0) atomicExch(&s_fail, 1);
  }
  __syncthreads();
  if (s_fail) {
    if (threadIdx.x == 0) out[pair] = 0xFB;  // (uint8_t)CTP_ERR_SAMPLE_PATH
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (int)(blockDim.x >> 5);
  for (int i = warp; i < K; i += nwarp) {
    if (*(volatile int *)&s_blocked) break;
    double hit[3];
    const bool fr = warp_edge_nodes<L>(M, s1[3 * i], s1[3 * i + 1], s1[3 * i + 2], s2[3 * i], s2[3 * i + 1],
                                       s2[3 * i + 2], clr, cdc, lane, hit);
    if (!fr && lane == 0) atomicExch(&s_blocked, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[pair] = s_blocked ? 0 : 1;
}

// ---------------- TopologicalPRMClustering::shorten_path (:2313-2479) ------------------------
// One CTA per polyline.  The anchor loop is sequential in the reference; here the 8 warps test
// anchor -> sampled[i0 .. i0+7] speculatively, the smallest blocked i wins, and everything the
// sequential loop would have done before that i was side-effect free, so the result is the
// sequential one.  Thread 0 replays the gradient push-off (:2367-2421) in reference order.
template <int L>
__global__ void __launch_bounds__(CTP_PATH_BLOCK)
k_shorten(MapDev M, const double *__restrict__ pts, const int32_t *__restrict__ off,
          const double *__restrict__ length, int forward_all, const uint8_t *__restrict__ forward_each, double clr,
          double cdc, int32_t cap, const int64_t *__restrict__ scratch_off, double *__restrict__ scratch,
          double *__restrict__ out, int32_t *__restrict__ origin, int32_t *__restrict__ out_n,
          double *__restrict__ out_len, int32_t *__restrict__ status) {
  const int pi = blockIdx.x;
  const int forward = forward_each ? (int)forward_each[pi] : forward_all;  // per polyline, or one direction for the batch
  const int npts = off[pi + 1] - off[pi];
  const double *P = pts + 3 * (size_t)off[pi];
  const double len = length[pi];
  double *O = out + 3 * (size_t)pi * cap;
  int32_t *OR = origin + (size_t)pi * cap;
  double *sampled = scratch + 3 * scratch_off[pi];
  const double num_samples = ceil(len / cdc) + 1;  // :2326
  const int S = num_samples > 0 ? (int)num_samples : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (int)(blockDim.x >> 5);

  __shared__ int s_rc, s_nout, s_i0, s_state;  // s_state: 0 run, 1 "no point added", <0 error
  __shared__ double s_anchor[3], s_slen;
  __shared__ int s_blk[CTP_PATH_BLOCK / 32];
  __shared__ double s_hit[CTP_PATH_BLOCK / 32][3];

  if (threadIdx.x == 0) {
    s_rc = (npts >= 2 && cap >= 2) ? dev_sample_path(P, npts, len, num_samples, sampled, !forward) : -1;
    const int first = forward ? 0 : npts - 1;
    s_anchor[0] = P[3 * first]; s_anchor[1] = P[3 * first + 1]; s_anchor[2] = P[3 * first + 2];
    if (cap >= 1 && npts >= 1) { O[0] = s_anchor[0]; O[1] = s_anchor[1]; O[2] = s_anchor[2]; OR[0] = first; }
    s_nout = 1;
    s_slen = 0.0;
    s_i0 = 1;
    s_state = s_rc == 0 ? 0 : -5;
  }
  __syncthreads();

  while (s_state == 0 && s_i0 < S) {
    const int i = s_i0 + warp;
    bool blocked = false;
    double hit[3] = {0, 0, 0};
    if (i < S)
      blocked = !warp_edge_nodes<L>(M, s_anchor[0], s_anchor[1], s_anchor[2], sampled[3 * i], sampled[3 * i + 1],
                                    sampled[3 * i + 2], clr, cdc, lane, hit);
    if (lane == 0) {
      s_blk[warp] = blocked;
      s_hit[warp][0] = hit[0]; s_hit[warp][1] = hit[1]; s_hit[warp][2] = hit[2];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int wfirst = -1;
      for (int w = 0; w < nwarp; w++)
        if (s_blk[w]) { wfirst = w; break; }
      if (wfirst < 0) {
        s_i0 += nwarp;
      } else {
        const int ib = s_i0 + wfirst;
        double g[3], vc[3];
        esdf_gradient<L>(M, s_hit[wfirst][0], s_hit[wfirst][1], s_hit[wfirst][2], g, vc);
        const double nx = sampled[3 * ib] - s_anchor[0], ny = sampled[3 * ib + 1] - s_anchor[1],
                     nz = sampled[3 * ib + 2] - s_anchor[2];
        // normal_gradient_new_point = gradient.cross(new_point_vector) (:2383)
        const double ngx = g[1] * nz - g[2] * ny, ngy = g[2] * nx - g[0] * nz, ngz = g[0] * ny - g[1] * nx;
        // vec_from_collision = new_point_vector.cross(normal_gradient_new_point), normalised (:2390-2392)
        double vx = ny * ngz - nz * ngy, vy = nz * ngx - nx * ngz, vz = nx * ngy - ny * ngx;
        const double z = (vx * vx + vy * vy) + vz * vz;
        if (z > 0.0) { const double s = sqrt(z); vx /= s; vy /= s; vz /= s; }
        bool added = false;
        for (double tci = clr; tci <= clr * 4; tci += cdc) {  // :2405
          const double qx = vc[0] + vx * tci, qy = vc[1] + vy * tci, qz = vc[2] + vz * tci;
          if (!ctp_collides(esdf_clearance<L>(M, qx, qy, qz), clr)) {
            if (s_nout >= cap - 1) { s_state = -6; break; }
            const double dx = s_anchor[0] - qx, dy = s_anchor[1] - qy, dz = s_anchor[2] - qz;
            s_slen += sqrt((dx * dx + dy * dy) + dz * dz);
            O[3 * s_nout] = qx; O[3 * s_nout + 1] = qy; O[3 * s_nout + 2] = qz;
            OR[s_nout] = -1;
            s_nout++;
            s_anchor[0] = qx; s_anchor[1] = qy; s_anchor[2] = qz;
            added = true;
            break;
          }
        }
        if (!added && s_state == 0) s_state = 1;  // "no point added ..." -> original path (:2423-2427)
        s_i0 = ib + 1;
      }
    }
    __syncthreads();
  }

  if (threadIdx.x == 0) {
    int st = s_state;
    if (st == 0 && s_nout >= cap) st = -6;
    if (st == 0) {
      const int last = forward ? npts - 1 : 0;
      const double dx = s_anchor[0] - P[3 * last], dy = s_anchor[1] - P[3 * last + 1], dz = s_anchor[2] - P[3 * last + 2];
      s_slen += sqrt((dx * dx + dy * dy) + dz * dz);
      O[3 * s_nout] = P[3 * last]; O[3 * s_nout + 1] = P[3 * last + 1]; O[3 * s_nout + 2] = P[3 * last + 2];
      OR[s_nout] = last;
      s_nout++;
      if (!forward) {
        for (int a = 0, b = s_nout - 1; a < b; a++, b--) {
          for (int c = 0; c < 3; c++) { const double t = O[3 * a + c]; O[3 * a + c] = O[3 * b + c]; O[3 * b + c] = t; }
          const int32_t t = OR[a]; OR[a] = OR[b]; OR[b] = t;
        }
      }
      // self-check (:2463-2473): the reference exits when it fails; reported as status -2
      double calc = 0.0;
      for (int a = 1; a < s_nout; a++) {
        const double dx2 = O[3 * a] - O[3 * (a - 1)], dy2 = O[3 * a + 1] - O[3 * (a - 1) + 1], dz2 = O[3 * a + 2] - O[3 * (a - 1) + 2];
        calc += sqrt((dx2 * dx2 + dy2 * dy2) + dz2 * dz2);
      }
      const double *lo = O + 3 * (s_nout - 1), *li = P + 3 * (npts - 1);
      if (fabs(calc - s_slen) > 1E-4 || lo[0] != li[0] || lo[1] != li[1] || lo[2] != li[2]) st = -2;
      out_n[pi] = s_nout;
      out_len[pi] = s_slen;
    } else {  // original path back
      const int c = npts < cap ? npts : cap;
      for (int a = 0; a < c; a++) {
        O[3 * a] = P[3 * a]; O[3 * a + 1] = P[3 * a + 1]; O[3 * a + 2] = P[3 * a + 2];
        OR[a] = a;
      }
      out_n[pi] = c;
      out_len[pi] = len;
    }
    status[pi] = st;
  }
}
