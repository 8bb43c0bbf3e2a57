// ctp_edt.cuh -- exact Euclidean distance transform of a voxel occupancy grid on the device
// (the map pipeline's voxel grid -> signed distance field step,
// map.py:49-73, without mesh_to_sdf).  Integer arithmetic throughout: the squared distance of every
// voxel to the nearest feature voxel is exact, so the field is p * sqrt(d2) with one rounding each.
//
// Three separable passes (Meijster, Roerdink, Hesselink 2000): a binary scan along z, then the lower
// envelope of the parabolas (u - i)^2 + g[i] along y and along x.  In the envelope passes one thread
// owns one line and consecutive threads own consecutive z, so every access of the line data and of
// the envelope stacks (depth-major in scratch) is coalesced.
#pragma once
#include <cstdint>

#define CTP_EDT_INF (1 << 28)  // "no feature on this line / plane"; larger than any in-grid squared distance

// pass 1, one warp per (x, y) line, z contiguous: squared distance along z to the nearest feature voxel.
// feature = (occ != 0) == (want_occupied != 0)
__global__ void __launch_bounds__(256) k_edt_scan_z(const uint8_t *__restrict__ occ, int n, int want_occupied,
                                                    int32_t *__restrict__ g) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t line = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (line >= (int64_t)n * n) return;
  const uint8_t *src = occ + line * n;
  int32_t *dst = g + line * n;
  // forward: index of the nearest feature at or before z
  int carry = -CTP_EDT_INF;
  for (int z0 = 0; z0 < n; z0 += 32) {
    const int z = z0 + lane;
    int v = -CTP_EDT_INF;
    if (z < n && ((src[z] != 0) == (want_occupied != 0))) v = z;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(full, v, o);
      if (lane >= o) v = max(v, t);
    }
    v = max(v, carry);
    carry = __shfl_sync(full, v, 31);
    if (z < n) dst[z] = v;  // temporarily the index, not the distance
  }
  // backward: nearest feature at or after z; combine
  carry = CTP_EDT_INF;
  for (int z0 = ((n - 1) / 32) * 32; z0 >= 0; z0 -= 32) {
    const int z = z0 + lane;
    int v = CTP_EDT_INF;
    int before = -CTP_EDT_INF;
    if (z < n) {
      before = dst[z];
      if (before == z) v = z;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_down_sync(full, v, o);
      if (lane + o < 32) v = min(v, t);
    }
    v = min(v, carry);
    carry = __shfl_sync(full, v, 0);
    if (z < n) {
      const int64_t dl = before <= -CTP_EDT_INF ? (int64_t)CTP_EDT_INF : (int64_t)(z - before);
      const int64_t dr = v >= CTP_EDT_INF ? (int64_t)CTP_EDT_INF : (int64_t)(v - z);
      const int64_t d = dl < dr ? dl : dr;
      dst[z] = d >= CTP_EDT_INF ? CTP_EDT_INF : (int32_t)(d * d);
    }
  }
}

// passes 2 and 3: out[u] = min_i (u - i)^2 + g[i] along one axis (element stride `stride`), one thread per line.
// line -> base offset: (line / inner) * outer_stride + (line % inner), inner = number of contiguous lines.
// st_s / st_t: envelope stacks, [depth][line] (uint16: n <= 32768).
__global__ void __launch_bounds__(256) k_edt_envelope(const int32_t *__restrict__ g, int n, int64_t n_lines, int64_t inner,
                                                      int64_t outer_stride, int64_t stride, uint16_t *__restrict__ st_s,
                                                      uint16_t *__restrict__ st_t, int32_t *__restrict__ out) {
  const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines) return;
  const int64_t base = (line / inner) * outer_stride + (line % inner);
  const int32_t *gl = g + base;
  int32_t *ol = out + base;
  uint16_t *s = st_s + line, *t = st_t + line;
  int q = 0;
  s[0] = 0;
  t[0] = 0;
  // the parabola on top of the stack, kept in registers
  int sq = 0, tq = 0;
  int64_t gsq = gl[0];
  for (int u = 1; u < n; u++) {
    const int64_t gu = gl[(int64_t)u * stride];
    for (;;) {
      const int64_t a = (int64_t)(tq - sq) * (tq - sq) + gsq;
      const int64_t b = (int64_t)(tq - u) * (tq - u) + gu;
      if (a <= b) break;
      q--;
      if (q < 0) break;
      sq = s[(int64_t)q * n_lines];
      tq = t[(int64_t)q * n_lines];
      gsq = gl[(int64_t)sq * stride];
    }
    if (q < 0) {
      q = 0;
      sq = u; tq = 0; gsq = gu;
      s[0] = (uint16_t)u;
      t[0] = 0;
    } else {
      // largest x where parabola sq is still not above parabola u: floor(num / den), den > 0
      const int64_t num = (int64_t)u * u - (int64_t)sq * sq + gu - gsq;
      const int64_t den = 2 * (int64_t)(u - sq);
      int64_t sep = num / den;
      if (num % den != 0 && num < 0) sep--;
      const int64_t w = sep + 1;
      if (w < n) {
        q++;
        sq = u; tq = (int)w; gsq = gu;
        s[(int64_t)q * n_lines] = (uint16_t)u;
        t[(int64_t)q * n_lines] = (uint16_t)w;
      }
    }
  }
  for (int u = n - 1; u >= 0; u--) {
    const int64_t v = (int64_t)(u - sq) * (u - sq) + gsq;
    ol[(int64_t)u * stride] = v >= CTP_EDT_INF ? CTP_EDT_INF : (int32_t)v;
    if (u == tq && q > 0) {
      q--;
      sq = s[(int64_t)q * n_lines];
      tq = t[(int64_t)q * n_lines];
      gsq = gl[(int64_t)sq * stride];
    }
  }
}

// signed field: free voxel -> +p*sqrt(d2 to the nearest occupied voxel); occupied voxel -> -p*sqrt(d2 to the
// nearest free voxel).  Called once per feature kind; writes only the voxels of the other kind.  A grid with
// no feature voxel at all gets the sentinel d2 = 3 n^2 (farther than any in-grid distance).
__global__ void __launch_bounds__(256) k_edt_emit(const int32_t *__restrict__ d2, const uint8_t *__restrict__ occ, int n,
                                                  double pitch, int features_occupied, float *__restrict__ vox,
                                                  int32_t *__restrict__ d2_out) {
  const size_t total = (size_t)n * n * n;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const bool is_occ = occ[o] != 0;
    if (is_occ == (features_occupied != 0)) continue;  // a feature voxel of this run: the other run writes it
    int32_t v = d2[o];
    if (v >= CTP_EDT_INF) v = 3 * n * n;
    const double d = pitch * sqrt((double)v);
    vox[o] = (float)(features_occupied ? d : -d);
    if (d2_out) d2_out[o] = features_occupied ? v : -v;
  }
}
