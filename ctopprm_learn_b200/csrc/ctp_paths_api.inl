// ctp_paths_api.inl -- host-pointer entry points for the polyline operations
// (included at the end of ctp_api.cu).

// threads per CTA of k_shorten / k_deformable for a batch of `items` polylines / pairs
static int ctp_path_threads(const ctp_map *m, int64_t items) { return items <= 2 * (int64_t)m->sm_count ? CTP_PATH_BLOCK : 256; }

static int check_polylines(const int32_t *off, int64_t count) {
  REQUIRE(off[0] == 0, "offsets must start at 0");
  for (int64_t i = 0; i < count; i++) REQUIRE(off[i + 1] - off[i] >= 2, "every polyline needs at least 2 points");
  return CTP_OK;
}

extern "C" int ctp_sample_path(ctp_map *m, const double *pts, const int32_t *off, const double *length,
                               const double *num_samples, int64_t count, double *out, const int64_t *out_off,
                               int32_t *status) {
  REQUIRE(m && count >= 0, "bad argument");
  if (count == 0) return CTP_OK;
  REQUIRE(pts && off && length && num_samples && out && out_off && status, "null buffer");
  CKR(check_polylines(off, count));
  for (int64_t i = 0; i < count; i++)
    REQUIRE(num_samples[i] >= 0 && out_off[i + 1] - out_off[i] >= (int64_t)num_samples[i], "out_off too tight");
  CK(cudaSetDevice(m->device));
  const int64_t total = off[count], otot = out_off[count];
  CKR(arena_reserve(m->io, align_up(total * 24) + align_up((count + 1) * 4) + 2 * align_up(count * 8) +
                               align_up(otot * 24 + 24) + align_up((count + 1) * 8) + align_up(count * 4)));
  IoScope sc(m);
  double *d_p = arena_take<double>(m->io, total * 3);
  int32_t *d_off = arena_take<int32_t>(m->io, count + 1);
  double *d_len = arena_take<double>(m->io, count), *d_ns = arena_take<double>(m->io, count);
  double *d_out = arena_take<double>(m->io, otot * 3 + 3);
  int64_t *d_oo = arena_take<int64_t>(m->io, count + 1);
  int32_t *d_st = arena_take<int32_t>(m->io, count);
  CK(cudaMemcpyAsync(d_p, pts, total * 24, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_off, off, (count + 1) * 4, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_len, length, count * 8, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_ns, num_samples, count * 8, cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d_oo, out_off, (count + 1) * 8, cudaMemcpyHostToDevice, 0));
  k_sample_path<<<(int)((count + 63) / 64), 64>>>(d_p, d_off, d_len, d_ns, count, d_out, d_oo, d_st);
  m->launches++;
  CK(cudaGetLastError());
  if (otot) CK(cudaMemcpyAsync(out, d_out, otot * 24, cudaMemcpyDeviceToHost, 0));
  CK(cudaMemcpyAsync(status, d_st, count * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return CTP_OK;
}

extern "C" int ctp_deformable(ctp_map *m, const double *pts, const int32_t *off, const double *length,
                              int64_t count, const int32_t *ia, const int32_t *ib, const double *num_check,
                              int64_t pairs, double clr, double cdc, uint8_t *out) {
  REQUIRE(m && count >= 0 && pairs >= 0, "bad argument");
  if (pairs == 0) return CTP_OK;
  REQUIRE(pts && off && length && ia && ib && num_check && out, "null buffer");
  REQUIRE(cdc > 0, "collision_distance_check must be > 0");
  CKR(check_polylines(off, count));
  std::vector<int64_t> soff((size_t)pairs + 1);
  soff[0] = 0;
  for (int64_t i = 0; i < pairs; i++) {
    REQUIRE(ia[i] >= 0 && ia[i] < count && ib[i] >= 0 && ib[i] < count, "pair index out of range");
    REQUIRE(num_check[i] >= 0 && num_check[i] < 1e8, "num_check out of range");
    soff[i + 1] = soff[i] + 2 * (int64_t)num_check[i] + 2;
  }
  CK(cudaSetDevice(m->device));
  const int64_t total = off[count];
  const size_t in_bytes = align_up(total * 24) + align_up((count + 1) * 4) + align_up(count * 8) + 2 * align_up(pairs * 4) +
                          align_up(pairs * 8) + align_up((pairs + 1) * 8);
  CKR(arena_reserve(m->io, in_bytes + align_up(soff[pairs] * 24) + align_up(pairs)));
  CKR(stage_reserve(m, in_bytes + align_up(pairs)));
  IoScope sc(m);
  StagedIn in(m, 0);
  double *d_p = in.put(pts, total * 3);
  int32_t *d_off = in.put(off, count + 1);
  double *d_len = in.put(length, count);
  int32_t *d_ia = in.put(ia, pairs), *d_ib = in.put(ib, pairs);
  double *d_nc = in.put(num_check, pairs);
  int64_t *d_so = in.put(soff.data(), pairs + 1);
  double *d_scr = arena_take<double>(m->io, soff[pairs] * 3);
  StagedOut res(m, in_bytes);
  uint8_t *d_out = res.take<uint8_t>(pairs);
  CK(in.upload(0));
  DISPATCH_L(m, k_deformable<L><<<(unsigned)pairs, ctp_path_threads(m, pairs)>>>(m->dev, d_p, d_off, d_len, d_ia, d_ib, d_nc, d_so,
                                                                      d_scr, clr, cdc, d_out));
  m->launches++;
  CK(cudaGetLastError());
  CK(res.download(0));
  CK(cudaStreamSynchronize(0));
  res.get(d_out, out, pairs);
  return CTP_OK;
}

extern "C" int ctp_shorten_path(ctp_map *m, const double *pts, const int32_t *off, const double *length,
                                int64_t count, int forward, double clr, double cdc, int32_t cap, double *out,
                                int32_t *origin, int32_t *out_n, double *out_len, int32_t *status) {
  return ctp_shorten_path_ex(m, pts, off, length, count, forward, nullptr, clr, cdc, cap, out, origin, out_n, out_len, status);
}

extern "C" int ctp_shorten_path_ex(ctp_map *m, const double *pts, const int32_t *off, const double *length,
                                   int64_t count, int forward, const uint8_t *forward_each, double clr, double cdc,
                                   int32_t cap, double *out, int32_t *origin, int32_t *out_n, double *out_len,
                                   int32_t *status) {
  REQUIRE(m && count >= 0 && cap >= 2, "bad argument");
  if (count == 0) return CTP_OK;
  REQUIRE(pts && off && length && out && origin && out_n && out_len && status, "null buffer");
  REQUIRE(cdc > 0, "collision_distance_check must be > 0");
  CKR(check_polylines(off, count));
  std::vector<int64_t> soff((size_t)count + 1);
  soff[0] = 0;
  for (int64_t i = 0; i < count; i++) {
    REQUIRE(length[i] >= 0 && length[i] / cdc < 1e8, "path length out of range");
    REQUIRE(off[i + 1] - off[i] <= cap, "cap must hold the original polyline");
    soff[i + 1] = soff[i] + (int64_t)(ceil(length[i] / cdc) + 1) + 2;
  }
  CK(cudaSetDevice(m->device));
  const int64_t total = off[count];
  const size_t in_bytes = align_up(total * 24) + align_up((count + 1) * 4) + align_up(count * 8) + align_up((count + 1) * 8) +
                          align_up(count);
  const size_t out_bytes = align_up(count * cap * 24) + align_up(count * cap * 4) + 2 * align_up(count * 4) + align_up(count * 8);
  CKR(arena_reserve(m->io, in_bytes + align_up(soff[count] * 24) + out_bytes));
  CKR(stage_reserve(m, in_bytes + out_bytes));
  IoScope sc(m);
  StagedIn in(m, 0);
  double *d_p = in.put(pts, total * 3);
  int32_t *d_off = in.put(off, count + 1);
  double *d_len = in.put(length, count);
  int64_t *d_so = in.put(soff.data(), count + 1);
  uint8_t *d_fw = forward_each ? in.put(forward_each, count) : nullptr;
  double *d_scr = arena_take<double>(m->io, soff[count] * 3);
  StagedOut res(m, in_bytes);
  double *d_out = res.take<double>(count * cap * 3);
  int32_t *d_or = res.take<int32_t>(count * cap);
  int32_t *d_on = res.take<int32_t>(count), *d_st = res.take<int32_t>(count);
  double *d_ol = res.take<double>(count);
  CK(in.upload(0));
  CK(cudaMemsetAsync(d_or, 0xff, count * cap * 4, 0));
  CK(cudaMemsetAsync(d_out, 0, count * cap * 24, 0));
  DISPATCH_L(m, k_shorten<L><<<(unsigned)count, ctp_path_threads(m, count)>>>(m->dev, d_p, d_off, d_len, forward, d_fw, clr, cdc, cap,
                                                                   d_so, d_scr, d_out, d_or, d_on, d_ol, d_st));
  m->launches++;
  CK(cudaGetLastError());
  CK(res.download(0));
  CK(cudaStreamSynchronize(0));
  res.get(d_out, out, count * cap * 3);
  res.get(d_or, origin, count * cap);
  res.get(d_on, out_n, count);
  res.get(d_ol, out_len, count);
  res.get(d_st, status, count);
  return CTP_OK;
}
