"""Parity of the CUDA path (through the C ABI) against the oracle and the committed golden
vectors.  Integer/boolean/index outputs and FP64 values are compared BIT-EXACT
(array_equal with NaN positions identical) -- tighter than north_star's 1e-5 relative."""
import os

import numpy as np
import pytest

from ctopprm_learn_b200 import capi, synth

pytestmark = pytest.mark.gpu

CLR, CDC = synth.MIN_CLEARANCE, synth.COLLISION_DISTANCE_CHECK
ALL_LAYOUTS = int(os.environ.get("CTP_TEST_LAYOUTS", capi.LAYOUT_PLAIN | capi.LAYOUT_CELL8 | capi.LAYOUT_TEX | capi.LAYOUT_BRICK))
LAYOUTS = [l for l in (capi.LAYOUT_PLAIN, capi.LAYOUT_CELL8, capi.LAYOUT_TEX, capi.LAYOUT_BRICK) if l & ALL_LAYOUTS]


@pytest.fixture(scope="module")
def gmap(golden):
    m = capi.DeviceMap(golden["vox"], golden["centroid"], golden["extents"], layouts=ALL_LAYOUTS)
    yield m
    m.close()


@pytest.fixture(scope="module")
def c1(orc, c1_map):
    vox, c, e = c1_map
    m = capi.DeviceMap(vox, c, e, layouts=ALL_LAYOUTS)
    yield m, orc.OracleMap(vox, c, e), c, e
    m.close()


@pytest.mark.parametrize("layout", LAYOUTS)
def test_golden_pointwise(golden, gmap, layout):
    gmap.set_layout(layout)
    assert np.array_equal(gmap.get_min_pos(), golden["min_pos"])
    assert np.array_equal(gmap.get_max_pos(), golden["max_pos"])
    assert np.array_equal(gmap.clearance(golden["pts"]), golden["clearance"], equal_nan=True)
    assert np.array_equal(gmap.in_collision(golden["pts"], CLR), golden["in_collision"])
    g, c = gmap.gradient(golden["pts"][:1500])
    assert np.array_equal(g, golden["grad"], equal_nan=True)
    assert np.array_equal(c, golden["centre"], equal_nan=True)


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("group", [0, 1, 4, 8, 16, 32])
def test_golden_edges(golden, gmap, layout, group):
    gmap.set_layout(layout)
    free, hit, hit_idx, lookups = gmap.edge_check(golden["a"], golden["b"], CLR, CDC, group=group)
    assert np.array_equal(free, golden["free_nodes"])
    assert np.array_equal(hit, golden["hit"], equal_nan=True)
    assert np.array_equal(lookups[~free], hit_idx[~free])
    fg = gmap.edge_check(golden["a"], golden["b"], CLR, CDC, mode=capi.EDGE_GUARDS, group=group)[0]
    assert np.array_equal(fg, golden["free_guards"])
    f2 = gmap.edge_check(golden["a"], golden["b"], 0.05, 0.11, group=group)[0]
    assert np.array_equal(f2, golden["free_nodes_alt"])


@pytest.mark.parametrize("layout", LAYOUTS)
def test_c1_edges_vs_oracle(c1, layout):
    m, om, c, e = c1
    m.set_layout(layout)
    a, b = synth.random_edges(c, e, 200000, 3, lo=0.0, hi=3.0)
    m.read_lookups()
    f0, h0, hi0, l0, tot = om.edge_nodes(a, b, CLR, CDC, nthreads=8)
    for group in (0, 1):  # heuristic choice and the flat kernel (long edges => several rounds)
        free, hit, hit_idx, lookups = m.edge_check(a, b, CLR, CDC, group=group)
        assert np.array_equal(free, f0)
        assert np.array_equal(hit, h0, equal_nan=True)
        assert np.array_equal(hit_idx, hi0)
        assert np.array_equal(lookups, l0)
        assert m.read_lookups() == tot  # algorithmic lookup counter used by the roofline
    p = c + (np.random.default_rng(2).random((300000, 3)) - 0.5) * (e + 0.2)
    assert np.array_equal(m.clearance(p), om.clearance(p, nthreads=8), equal_nan=True)


def test_empty_and_invalid(c1):
    m, om, c, e = c1
    assert len(m.clearance(np.zeros((0, 3)))) == 0
    assert len(m.edge_check(np.zeros((0, 3)), np.zeros((0, 3)), CLR, CDC)[0]) == 0
    with pytest.raises(capi.CtpError):
        m.edge_check(np.zeros((1, 3)), np.ones((1, 3)), CLR, 0.0)  # cdc == 0 (reference exits)
    # NaN endpoints: reference loop never runs => free (src/base_map.cpp:57-64)
    a = np.array([[np.nan, 0, 0]])
    assert m.edge_check(a, a + 1, CLR, CDC)[0][0] == om.edge_nodes(a, a + 1, CLR, CDC)[0][0]


def test_indexed_edges(c1):
    m, om, c, e = c1
    rng = np.random.default_rng(7)
    nodes = c + (rng.random((5000, 3)) - 0.5) * e * 0.95
    frm = rng.integers(0, 5000, 60000).astype(np.int32)
    to = rng.integers(0, 5000, 60000).astype(np.int32)
    free, hit_idx, lookups = m.edge_check_indexed(nodes, frm, to, CLR, CDC)
    f0, hi0, _ = om.edge_index(nodes, frm, to, CLR, CDC, nthreads=8)
    assert np.array_equal(free, f0) and np.array_equal(hit_idx, hi0)


def _queries(c, e, q, n, seed):
    s, g = synth.random_queries(c, e, q, seed)
    cand = np.stack([synth.ellipsoid_candidates(s[i], g[i], 4 * n, seed * 100 + i) for i in range(q)])
    return s, g, cand


def test_sample_reject(c1):
    m, om, c, e = c1
    q, n = 5, 700
    s, g, cand = _queries(c, e, q, n, 21)
    nodes, n_filled, n_used, accept = m.sample_reject(s, g, cand, CLR, n)
    for i in range(q):
        n0, filled, used, acc = om.sample_reject(s[i], g[i], cand[i], CLR, n)
        assert filled == n == n_filled[i]
        assert used == n_used[i]
        assert np.array_equal(nodes[i], n0)
        assert np.array_equal(accept[i, :used], acc[:used])
    # candidate stream too short: reported, outputs still consistent
    nodes2, nf2, nu2, _ = m.sample_reject(s[:1], g[:1], cand[:1, :50], CLR, n, allow_short=True)
    n0, filled, used, _ = om.sample_reject(s[0], g[0], cand[0, :50], CLR, n)
    assert nf2[0] == filled < n and nu2[0] == used == 50
    assert np.array_equal(nodes2[0, :filled], n0[:filled])
    with pytest.raises(capi.CtpError) as ei:
        m.sample_reject(s[:1], g[:1], cand[:1, :50], CLR, n)
    assert ei.value.code == capi.ERR_NEED_CANDIDATES


@pytest.mark.parametrize("k", [1, 5, 13, 16])
def test_knn_exact(c1, orc, k):
    m, om, c, e = c1
    rng = np.random.default_rng(k)
    pts = c + (rng.random((3, 3000, 3)) - 0.5) * e
    pts[1, :, 2] = c[2]            # planar query (planar: true => every z equal)
    pts[2, 100:140] = pts[2, 100]  # duplicates: (distance, index) tie-break
    idx, d2 = m.knn(pts, k)
    for i in range(3):
        i0, d0 = orc.knn(pts[i], k, nthreads=8)
        assert np.array_equal(idx[i], i0)
        assert np.array_equal(d2[i], d0)


def test_knn_tiny(c1, orc):
    m, om, c, e = c1
    pts = c + (np.random.default_rng(0).random((1, 9, 3)) - 0.5)
    idx, d2 = m.knn(pts, 13)  # fewer than k others: -1 padded
    i0, _ = orc.knn(pts[0], 13)
    assert np.array_equal(idx[0], i0) and (idx[0, :, 8:] == -1).all()


def _csr_equal(got, q, want):
    lo, hi = got["nnz_off"][q], got["nnz_off"][q + 1]
    assert np.array_equal(got["row_ptr"][q], want["row_ptr"])
    assert np.array_equal(got["col"][lo:hi], want["col"])
    assert np.array_equal(got["w"][lo:hi], want["w"])  # FP64 weights bit-exact


@pytest.mark.parametrize("layout", LAYOUTS)
def test_build_prm_csr(c1, layout):
    m, om, c, e = c1
    m.set_layout(layout)
    q, n = 3, 1000
    s, g, cand = _queries(c, e, q, n, 33)
    nodes = m.sample_reject(s, g, cand, CLR, n)[0]
    got = m.build_prm(nodes, CLR, CDC)
    for i in range(q):
        want = om.build_prm(nodes[i], CLR, CDC, nthreads=8)
        assert np.array_equal(got["nbr"][i], want["nbr"])
        assert np.array_equal(got["directed_free"][i], want["directed_free"])
        _csr_equal(got, i, want)
        # adjacency is symmetric with equal weights
        rp, col = want["row_ptr"], want["col"]
        pairs = {(r, cc) for r in range(n) for cc in col[rp[r]:rp[r + 1]]}
        assert all((b, a) in pairs for a, b in pairs)


@pytest.mark.parametrize("k", [3, 15, 16])
def test_build_prm_other_k(c1, k):
    m, om, c, e = c1
    m.set_layout(capi.LAYOUT_PLAIN)
    q, n = 2, 600
    s, g, cand = _queries(c, e, q, n, 50 + k)
    nodes = m.sample_reject(s, g, cand, CLR, n)[0]
    got = m.build_prm(nodes, CLR, CDC, k=k)
    for i in range(q):
        want = om.build_prm(nodes[i], CLR, CDC, k=k, nthreads=8)
        assert np.array_equal(got["nbr"][i], want["nbr"])
        assert np.array_equal(got["directed_free"][i], want["directed_free"])
        _csr_equal(got, i, want)


def test_build_prm_fewer_nodes_than_k(c1):
    m, om, c, e = c1
    s, g, cand = _queries(c, e, 1, 9, 77)
    nodes = m.sample_reject(s, g, cand, CLR, 9)[0]
    got = m.build_prm(nodes, CLR, CDC)  # rows are padded with -1 neighbours
    want = om.build_prm(nodes[0], CLR, CDC)
    assert np.array_equal(got["nbr"][0], want["nbr"]) and (got["nbr"][0][:, 8:] == -1).all()
    _csr_equal(got, 0, want)


def test_sample_multiple_end_to_end(c1):
    m, om, c, e = c1
    q, n = 2, 800
    s, g, cand = _queries(c, e, q, n, 44)
    out = m.sample_multiple(s, g, cand, n, CLR, CDC)
    for i in range(q):
        n0 = om.sample_reject(s[i], g[i], cand[i], CLR, n)[0]
        assert np.array_equal(out["nodes"][i], n0)
        want = om.build_prm(n0, CLR, CDC, nthreads=8)
        _csr_equal(out, i, want)


@pytest.mark.parametrize("chunk", [1, 2, 3])
def test_sample_multiple_chunk_pipeline(c1, chunk):
    """the host API cuts the queries into chunks on three streams; results must not depend on it"""
    m, om, c, e = c1
    q, n = 5, 500
    s, g, cand = _queries(c, e, q, n, 61)
    pin = {k_: capi.pinned_empty(v.shape, v.dtype) for k_, v in dict(s=s, g=g, cand=cand).items()}
    pin["s"][:], pin["g"][:], pin["cand"][:] = s, g, cand
    m.set_tunable(capi.TUNE_PIPE_CHUNK, chunk)
    try:
        out = m.sample_multiple(pin["s"], pin["g"], pin["cand"], n, CLR, CDC)
    finally:
        m.set_tunable(capi.TUNE_PIPE_CHUNK, 0)
    assert out["nnz_off"][0] == 0
    for i in range(q):
        n0 = om.sample_reject(s[i], g[i], cand[i], CLR, n)[0]
        assert np.array_equal(out["nodes"][i], n0)
        _csr_equal(out, i, om.build_prm(n0, CLR, CDC, nthreads=8))


def test_sample_multiple_short_stream(c1):
    m, om, c, e = c1
    q, n = 3, 400
    s, g, cand = _queries(c, e, q, n, 62)
    cand = cand.copy()
    cand[1, 100:] = c - e  # query 1: every later candidate lies outside the map
    m.set_tunable(capi.TUNE_PIPE_CHUNK, 1)
    try:
        with pytest.raises(capi.CtpError) as ei:
            m.sample_multiple(s, g, cand, n, CLR, CDC)
    finally:
        m.set_tunable(capi.TUNE_PIPE_CHUNK, 0)
    assert ei.value.code == capi.ERR_NEED_CANDIDATES


def _split(g, key_pts, key_off):
    off = g[key_off]
    return [g[key_pts][off[i]:off[i + 1]] for i in range(len(off) - 1)]


def test_golden_polylines(golden, gmap):
    gmap.set_layout(capi.LAYOUT_PLAIN)
    polys = _split(golden, "poly_pts", "poly_off")
    lens, nums = golden["poly_len"], golden["poly_num"]
    out, st = gmap.sample_path(polys, lens, nums)
    assert (st == 0).all()
    assert np.array_equal(np.concatenate(out), golden["sampled"])
    nc = [float(np.ceil(max(lens[i], lens[j]) / CDC) + 1) for i, j in zip(golden["pair_i"], golden["pair_j"])]
    d = gmap.deformable(polys, lens, golden["pair_i"], golden["pair_j"], nc, 0.02, CDC)
    assert np.array_equal(d, golden["deformable"].astype(np.uint8))
    near = list(golden["near_pts"].reshape(3, -1, 3))
    nl = golden["near_len"]
    ia, ib = [0, 0, 1], [1, 2, 2]
    nc = [float(np.ceil(max(nl[i], nl[j]) / CDC) + 1) for i, j in zip(ia, ib)]
    d = gmap.deformable(near, nl, ia, ib, nc, -10.0, CDC)
    assert np.array_equal(d, golden["near_deformable"].astype(np.uint8))


def _dijkstra_path(row_ptr, col, w, src, dst):
    import heapq
    n = len(row_ptr) - 1
    dist = np.full(n, np.inf)
    prev = np.full(n, -1)
    dist[src] = 0
    pq = [(0.0, src)]
    while pq:
        d, u = heapq.heappop(pq)
        if d > dist[u]:
            continue
        if u == dst:
            break
        for e in range(row_ptr[u], row_ptr[u + 1]):
            v = col[e]
            nd = d + w[e]
            if nd < dist[v]:
                dist[v] = nd
                prev[v] = u
                heapq.heappush(pq, (nd, v))
    if not np.isfinite(dist[dst]):
        return None, np.inf
    path = [dst]
    while path[-1] != src:
        path.append(prev[path[-1]])
    return path[::-1], dist[dst]


@pytest.fixture(scope="module")
def roadmap_paths(c1):
    """A few PRM paths on C1 (start->goal and via random waypoints) to shorten/compare."""
    m, om, c, e = c1
    n = 1500
    s, g = synth.default_query(c, e)
    cand = synth.ellipsoid_candidates(s, g, 6 * n, 5)
    out = m.sample_multiple(s[None], g[None], cand[None], n, CLR, CDC)
    nodes = out["nodes"][0]
    rp, col, w = out["row_ptr"][0], out["col"], out["w"]
    paths, lens = [], []
    rng = np.random.default_rng(3)
    for via in [None] + list(rng.integers(2, n, 6)):
        if via is None:
            p, L = _dijkstra_path(rp, col, w, 0, 1)
        else:
            p1, L1 = _dijkstra_path(rp, col, w, 0, int(via))
            p2, L2 = _dijkstra_path(rp, col, w, int(via), 1)
            if p1 is None or p2 is None:
                continue
            p, L = p1 + p2[1:], None
        if p is None:
            continue
        pts = nodes[p]
        paths.append(pts)
        lens.append(float(np.sum(np.sqrt(((pts[1:] - pts[:-1]) ** 2).sum(1)))) if L is None else float(L))
    assert len(paths) >= 3
    return paths, lens


@pytest.mark.parametrize("forward", [True, False])
def test_shorten_path(c1, orc, roadmap_paths, forward):
    m, om, c, e = c1
    m.set_layout(capi.LAYOUT_PLAIN)
    paths, _ = roadmap_paths
    lens = [orc.path_length(p) for p in paths]  # calc_path_length (include/dijkstra.hpp:27-38)
    got_pts, got_origin, got_len, st = m.shorten_path(paths, lens, forward, CLR, CDC, cap=256)
    for i, p in enumerate(paths):
        w_pts, w_origin, w_len, w_st, _ = om.shorten_path(p, lens[i], forward, CLR, CDC, cap=256)
        assert st[i] == w_st
        assert np.array_equal(got_pts[i], w_pts)
        assert np.array_equal(got_origin[i], w_origin)
        assert got_len[i] == w_len
        if w_st == 0:
            # reference self-check (:2463-2473)
            assert abs(orc.path_length(got_pts[i]) - got_len[i]) <= 1e-4
            assert got_len[i] <= lens[i] + 1e-9


def test_deformable_and_remove_equivalent(c1, roadmap_paths):
    m, om, c, e = c1
    m.set_layout(capi.LAYOUT_PLAIN)
    paths, lens = roadmap_paths
    P = len(paths)
    ia = [i for i in range(P) for j in range(i + 1, P)]
    ib = [j for i in range(P) for j in range(i + 1, P)]
    nc = [float(np.ceil(max(lens[i], lens[j]) / CDC) + 1) for i, j in zip(ia, ib)]
    d = m.deformable(paths, lens, ia, ib, nc, CLR, CDC)
    for t, (i, j) in enumerate(zip(ia, ib)):
        assert d[t] == om.deformable(paths[i], lens[i], paths[j], lens[j], nc[t], CLR, CDC)
    # removeEquivalentPaths (:2237-2298) replayed on the host from the batched matrix
    D = {(i, j): bool(v) for i, j, v in zip(ia, ib, d)}
    keep = list(range(P))
    i = 0
    while i < len(keep):
        shortest, rm = i, []
        for j in range(i + 1, len(keep)):
            a, b = keep[i], keep[j]
            if D[(min(a, b), max(a, b))]:
                rm.append(j)
                if lens[b] < lens[keep[shortest]] and (shortest == i or True):
                    shortest = j if lens[b] < lens[keep[shortest]] else shortest
        if shortest != i:
            keep[i] = keep[shortest]
        for j in reversed(rm):
            keep.pop(j)
        i += 1
    want = om.remove_equivalent(paths, lens, CLR, CDC)
    assert len(keep) == len(want)


def test_sample_ellipsoid_seeded(c1):
    m, om, c, e = c1
    s, g = synth.random_queries(c, e, 3, 8)
    a = m.sample_ellipsoid(s, g, 20000, 1.2, False, 1234)
    b = m.sample_ellipsoid(s, g, 20000, 1.2, False, 1234)
    d = m.sample_ellipsoid(s, g, 20000, 1.2, False, 1235)
    assert np.array_equal(a, b) and not np.array_equal(a, d)
    for i in range(3):
        # inside the prolate spheroid: |p-s| + |p-g| <= ratio * |g-s|
        two_major = 1.2 * np.linalg.norm(g[i] - s[i])
        sums = np.linalg.norm(a[i] - s[i], axis=1) + np.linalg.norm(a[i] - g[i], axis=1)
        assert (sums <= two_major * (1 + 1e-12)).all()
        # uniform in the spheroid: mean at the centre, axial variance a^2/5
        ctr = (s[i] + g[i]) / 2
        assert np.allclose(a[i].mean(0), ctr, atol=0.05 * two_major)
        ax = (g[i] - s[i]) / np.linalg.norm(g[i] - s[i])
        var = np.var((a[i] - ctr) @ ax)
        assert abs(var - (two_major / 2) ** 2 / 5) < 0.05 * (two_major / 2) ** 2
    pl = m.sample_ellipsoid(s, g, 100, 1.2, True, 1)
    assert (pl[..., 2] == s[:, None, 2]).all()


def test_csr_from_directed(c1):
    m, om, c, e = c1
    s, g, cand = _queries(c, e, 2, 500, 91)
    nodes = m.sample_reject(s, g, cand, CLR, 500)[0]
    full = m.build_prm(nodes, CLR, CDC)
    got = m.csr_from_directed(nodes, full["nbr"], full["directed_free"])
    assert np.array_equal(got["row_ptr"], full["row_ptr"]) and np.array_equal(got["col"], full["col"])
    assert np.array_equal(got["w"], full["w"]) and np.array_equal(got["nnz_off"], full["nnz_off"])


def test_sharded_single_query_device_backend(c1):
    """sharded.sample_multiple_sharded with the product backend, device tensors end to end over NCCL (one rank here;
    two ranks: test_sharded_single_query_two_gpus; the host logic at world sizes 2 and 3: tests/test_sharded_gloo.py)"""
    import torch.distributed as dist
    from ctopprm_learn_b200 import sharded
    m, om, c, e = c1
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29617")
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        n = 900
        s, g = synth.default_query(c, e)
        cand = synth.ellipsoid_candidates(s, g, 5 * n, 13)
        tm = {}
        out = sharded.sample_multiple_sharded(sharded.DeviceBackend(m), s, g, cand, n, CLR, CDC, timings=tm)
    finally:
        dist.destroy_process_group()
    assert set(tm) == {"reject", "allgather_nodes", "knn_and_edges", "allreduce_tables", "csr"}
    want_nodes = om.sample_reject(s, g, cand, CLR, n)[0]
    want = om.build_prm(want_nodes, CLR, CDC, nthreads=8)
    assert np.array_equal(out["nodes"], want_nodes) and np.array_equal(out["directed_free"], want["directed_free"])
    assert np.array_equal(out["nbr"], want["nbr"])
    assert np.array_equal(out["row_ptr"], want["row_ptr"]) and np.array_equal(out["col"], want["col"])
    assert np.array_equal(out["w"], want["w"])


@pytest.mark.parametrize("n,parts", [(900, 3), (20000, 8), (30, 4)])
def test_prm_row_slices_cover_the_roadmap(c1, n, parts):
    """ctp_prm_rows_dev: the slices that `parts` GPUs would take, run one after the other on this GPU -- disjoint, they
    cover every node, and their element-wise maximum is the neighbour table and the verdicts of the one-GPU build"""
    import torch
    from ctopprm_learn_b200 import sharded
    m, om, c, e = c1
    s, g = synth.default_query(c, e)
    cand = synth.ellipsoid_candidates(s, g, 6 * n, 19)
    nodes = m.sample_reject(s[None], g[None], cand[None], CLR, n)[0]
    full = m.build_prm(nodes, CLR, CDC)
    dev = torch.device("cuda", m.device)
    nodes_t = torch.from_numpy(nodes[0]).to(dev)
    k = 13
    nbr = torch.full((n, k), int(sharded.NBR_UNOWNED), dtype=torch.int32, device=dev)
    free = torch.zeros((n, k), dtype=torch.uint8, device=dev)
    owned = torch.zeros(n, dtype=torch.int32, device=dev)
    total = 0
    for r in range(parts):
        lo, hi = sharded.shard_range(n, r, parts)
        nb = torch.empty((n, k), dtype=torch.int32, device=dev)
        fr = torch.empty((n, k), dtype=torch.uint8, device=dev)
        total += m.prm_rows_dev(nodes_t, n, k, CLR, CDC, lo, hi, nb, fr, stream=torch.cuda.current_stream(dev))
        torch.cuda.synchronize(dev)
        mine = nb[:, 0] != int(sharded.NBR_UNOWNED)
        assert (fr[~mine] == 0).all()
        owned += mine.to(torch.int32)
        nbr = torch.maximum(nbr, nb)
        free = torch.maximum(free, fr)
    assert total == n and (owned == 1).all()  # every node is owned by exactly one slice
    assert np.array_equal(nbr.cpu().numpy(), full["nbr"][0])
    assert np.array_equal(free.cpu().numpy().astype(bool), full["directed_free"][0])


def _two_gpu_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    from ctopprm_learn_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        vox, c, e = synth.make_config_map("C1")
        m = capi.DeviceMap(vox, c, e, device=rank, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX)
        n = 5000
        s, g = synth.default_query(c, e)
        cand = synth.ellipsoid_candidates(s, g, 5 * n, 23)
        out = sharded.sample_multiple_sharded(sharded.DeviceBackend(m), s, g, cand, n, CLR, CDC, local_device=rank)
        np.savez(os.path.join(tmp, f"out{rank}.npz"), **out)
        m.close()
    finally:
        dist.destroy_process_group()


def test_sharded_single_query_two_gpus(c1, tmp_path):
    """one query split over two GPUs (NCCL all-gather of the node sets, MAX all-reduce of the tables) = one-GPU build"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_two_gpu_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    m, om, c, e = c1
    n = 5000
    s, g = synth.default_query(c, e)
    cand = synth.ellipsoid_candidates(s, g, 5 * n, 23)
    want_nodes = om.sample_reject(s, g, cand, CLR, n)[0]
    want = om.build_prm(want_nodes, CLR, CDC, nthreads=8)
    for r in range(2):
        out = np.load(tmp_path / f"out{r}.npz")
        assert np.array_equal(out["nodes"], want_nodes) and np.array_equal(out["nbr"], want["nbr"])
        assert np.array_equal(out["directed_free"], want["directed_free"])
        assert np.array_equal(out["row_ptr"], want["row_ptr"]) and np.array_equal(out["col"], want["col"])
        assert np.array_equal(out["w"], want["w"])


@pytest.mark.parametrize("group", [1, 8])
def test_very_long_marches(c1, group):
    """more than 65536 samples per edge: beyond the range of the verified division shortcut"""
    m, om, c, e = c1
    m.set_layout(capi.LAYOUT_PLAIN)
    a, b = synth.random_edges(c, e, 64, 5, lo=7.0, hi=12.0)
    cdc = 1e-4
    free, hit, hit_idx, lookups = m.edge_check(a, b, -5.0, cdc, group=group)  # clearance -5: nothing collides inside the map
    f0, h0, hi0, l0, _ = om.edge_nodes(a, b, -5.0, cdc, nthreads=8)
    assert (l0 > 65536).any()
    assert np.array_equal(free, f0) and np.array_equal(hit, h0, equal_nan=True)
    assert np.array_equal(hit_idx, hi0) and np.array_equal(lookups, l0)


@pytest.mark.parametrize("mode", [0, 1])
def test_knn_uneven_density_all_kernels(c1, orc, mode):
    """clusters, isolated outliers, duplicates and a thin slab: every search kernel and its fallbacks"""
    m, om, c, e = c1
    rng = np.random.default_rng(100 + mode)
    sets = []
    a = np.concatenate([c + rng.standard_normal((2500, 3)) * 0.15,          # tight cluster
                        c + (rng.random((60, 3)) - 0.5) * e,                # isolated outliers
                        c + np.array([3.0, 3.0, 0]) + rng.standard_normal((400, 3)) * 0.6])
    sets.append(a)
    b = c + (rng.random((2960, 3)) - 0.5) * e
    b[:, 2] = c[2] + (rng.random(2960) - 0.5) * 0.02                         # thin slab
    b[500:560] = b[500]                                                       # 60 duplicates
    sets.append(b)
    pts = np.stack(sets)
    m.set_tunable(capi.TUNE_KNN_TILE, mode)
    try:
        idx, d2 = m.knn(pts, 13)
    finally:
        m.set_tunable(capi.TUNE_KNN_TILE, 1)
    for i in range(2):
        i0, d0 = orc.knn_grid(pts[i], 13, nthreads=8)
        assert np.array_equal(idx[i], i0)
        assert np.array_equal(d2[i], d0)


@pytest.mark.parametrize("cluster", [0, 1])
def test_sssp_batched_vs_oracle(c1, orc, cluster):
    """ctp_sssp (shared-memory frontier Bellman-Ford, one CTA or one 8-CTA cluster per query) against the
    oracle's Dijkstra: distances, predecessors and flood labels bit-identical"""
    m, om, c, e = c1
    m.set_tunable(capi.TUNE_SSSP_CLUSTER, cluster)
    q, n = 4, 1200
    s, g, cand = _queries(c, e, q, n, 71)
    out = m.sample_multiple(s, g, cand, n, CLR, CDC)
    src1 = np.zeros((q, 1), dtype=np.int32)
    dist, pred = m.sssp(out["row_ptr"], out["col"], out["w"], out["nnz_off"], src1)
    src3 = np.tile(np.array([[0, 1, 600, -1]], dtype=np.int32), (q, 1))
    dist3, pred3, lab3 = m.sssp(out["row_ptr"], out["col"], out["w"], out["nnz_off"], src3, want_label=True)
    for i in range(q):
        lo, hi = out["nnz_off"][i], out["nnz_off"][i + 1]
        rp, col, w = out["row_ptr"][i], out["col"][lo:hi], out["w"][lo:hi]
        d0, p0 = orc.sssp(rp, col, w, [0])
        assert np.array_equal(dist[i], d0) and np.array_equal(pred[i], p0)
        d1, p1, l1 = orc.sssp(rp, col, w, [0, 1, 600, -1], want_label=True)
        assert np.array_equal(dist3[i], d1) and np.array_equal(pred3[i], p1) and np.array_equal(lab3[i], l1)
    assert np.isinf(dist).any() or True  # unreachable nodes (if any) are +inf with pred -1
    assert (pred[np.isinf(dist)] == -1).all()
    m.set_tunable(capi.TUNE_SSSP_CLUSTER, -1)


@pytest.mark.parametrize("n,deg,seed", [(64, 3, 1), (65, 2, 2), (257, 1, 3), (1000, 4, 4), (4097, 3, 5), (20000, 2, 6)])
def test_sssp_cluster_on_synthetic_graphs(c1, orc, n, deg, seed):
    """the cluster kernel (distances sliced over the distributed shared memory of 8 CTAs) on random symmetric
    graphs with sizes that do not divide into slices, several components and several sources"""
    m = c1[0]
    rng = np.random.default_rng(seed)
    a = np.repeat(np.arange(n), deg)
    b = rng.integers(0, n, n * deg)
    keep = a != b
    a, b = a[keep], b[keep]
    pair = np.unique(np.stack([np.minimum(a, b), np.maximum(a, b)], 1), axis=0)
    wgt = rng.random(len(pair)) + 0.01
    rows = np.concatenate([pair[:, 0], pair[:, 1]])
    cols = np.concatenate([pair[:, 1], pair[:, 0]])
    ww = np.concatenate([wgt, wgt])
    order = np.lexsort((cols, rows))
    rows, cols, ww = rows[order], cols[order].astype(np.int32), ww[order]
    rp = np.zeros(n + 1, dtype=np.int32)
    np.add.at(rp, rows + 1, 1)
    rp = np.cumsum(rp).astype(np.int32)
    off = np.array([0, len(cols)], dtype=np.int64)
    srcs = [0, n - 1, n // 2]
    want = orc.sssp(rp, cols, ww, srcs)
    try:
        for cluster in (1, 0):
            m.set_tunable(capi.TUNE_SSSP_CLUSTER, cluster)
            dist, pred = m.sssp(rp[None], cols, ww, off, np.array([srcs], dtype=np.int32))
            assert np.array_equal(dist[0], want[0]) and np.array_equal(pred[0], want[1]), cluster
    finally:
        m.set_tunable(capi.TUNE_SSSP_CLUSTER, -1)


@pytest.mark.parametrize("cluster", [0, 1])
def test_sssp_large_graph_global_path(c1, orc, cluster):
    """more nodes than fit in one CTA's shared memory: the same search on a global distance array (cluster = 0)
    or sliced over the shared memory of a cluster (cluster = 1, the default for such graphs)"""
    m, om, c, e = c1
    m.set_tunable(capi.TUNE_SSSP_CLUSTER, cluster)
    n = 30000
    s, g = synth.default_query(c, e)
    cand = synth.ellipsoid_candidates(s, g, 8 * n, 9)
    out = m.sample_multiple(s[None], g[None], cand[None], n, CLR, CDC)
    dist, pred = m.sssp(out["row_ptr"], out["col"], out["w"], out["nnz_off"], np.zeros((1, 1), dtype=np.int32))
    m.set_tunable(capi.TUNE_SSSP_CLUSTER, -1)
    d0, p0 = orc.sssp(out["row_ptr"][0], out["col"][:out["nnz_off"][1]], out["w"][:out["nnz_off"][1]], [0])
    assert np.array_equal(dist[0], d0) and np.array_equal(pred[0], p0)


@pytest.fixture(scope="module")
def c2(orc):
    vox, c, e = synth.make_config_map("C2")
    m = capi.DeviceMap(vox, c, e, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX)
    yield m, orc.OracleMap(vox, c, e), c, e
    m.close()


def test_c2_full_size_queries_vs_oracle(c2):
    """BASELINE configs[1] at full size (256^3 grid, 10 000 nodes, 130 000 directed checks per query)"""
    m, om, c, e = c2
    q, n = 3, 10000
    s, g, cand = _queries(c, e, q, n, 4)
    m.set_layout(capi.LAYOUT_TEX)
    out = m.sample_multiple(s, g, cand, n, CLR, CDC)
    for i in range(q):
        n0 = om.sample_reject(s[i], g[i], cand[i], CLR, n)[0]
        assert np.array_equal(out["nodes"][i], n0)
        _csr_equal(out, i, om.build_prm(n0, CLR, CDC, nthreads=8))


def test_large_roadmap_properties(c2):
    """100 000 nodes (the node count of configs[2]): properties that do not need the oracle's full run --
    symmetric adjacency with equal weights, ascending columns, weights = FP64 distances, every listed
    directed verdict reproduced by the point-wise segment check, k-NN distances ascending and exact"""
    m, om, c, e = c2
    n, k = 100000, 13
    s, g = synth.default_query(c, e)
    cand = synth.ellipsoid_candidates(s, g, 4 * n, 77)
    m.set_layout(capi.LAYOUT_TEX)
    nodes = m.sample_reject(s[None], g[None], cand[None], CLR, n)[0]
    out = m.build_prm(nodes, CLR, CDC)
    rp, col, w = out["row_ptr"][0], out["col"], out["w"]
    assert rp[-1] == len(col)
    rows = np.repeat(np.arange(n), np.diff(rp))
    assert (np.diff(col)[np.diff(rows) == 0] > 0).all()          # ascending within a row, no duplicates
    key = rows.astype(np.int64) * n + col
    rkey = col.astype(np.int64) * n + rows
    order = np.argsort(rkey, kind="stable")
    assert np.array_equal(np.sort(key), rkey[order])              # (i, j) present <=> (j, i) present
    assert np.array_equal(w[order], w[np.argsort(key, kind="stable")])  # ... with the same weight
    d = nodes[0][col] - nodes[0][rows]
    assert np.array_equal(w, np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]))
    # directed verdicts against the generic point-wise kernel on a sample of edges
    nbr, fr = out["nbr"][0], out["directed_free"][0]
    pick = np.random.default_rng(0).integers(0, n * k, 200000)
    a, b = nodes[0][pick // k], nodes[0][nbr.reshape(-1)[pick]]
    free = m.edge_check(a, b, CLR, CDC, group=8)[0]
    assert np.array_equal(free, fr.reshape(-1)[pick])
    # adjacency = union of the free directed edges
    want = set(map(tuple, np.stack([np.repeat(np.arange(n), k)[fr.reshape(-1)], nbr.reshape(-1)[fr.reshape(-1)]], 1)))
    got = set(zip(rows.tolist(), col.tolist()))
    assert got == want | {(j, i) for i, j in want}
    # k-NN on a sample of rows against brute force in float32
    p32 = nodes[0].astype(np.float32)
    for i in np.random.default_rng(1).integers(0, n, 40):
        dd = p32[i] - p32
        d2 = (dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) + dd[:, 2] * dd[:, 2]
        d2[i] = np.inf
        o = np.lexsort((np.arange(n), d2))[:k]
        assert np.array_equal(nbr[i], o)


@pytest.mark.parametrize("chunk", [0, 2])
def test_query_paths_end_to_end(c1, orc, chunk):
    """ctp_query_paths: host candidates -> host start->goal paths, roadmaps stay on the device"""
    m, om, c, e = c1
    q, n = 5, 900
    s, g, cand = _queries(c, e, q, n, 81)
    s[0], g[0] = synth.default_query(c, e)  # at least one query with start and goal in free space
    cand[0] = synth.ellipsoid_candidates(s[0], g[0], cand.shape[1], 3)
    m.set_tunable(capi.TUNE_PIPE_CHUNK, chunk)
    try:
        out = m.query_paths(s, g, cand, n, CLR, CDC, path_cap=256)
    finally:
        m.set_tunable(capi.TUNE_PIPE_CHUNK, 0)
    reached = 0
    for i in range(q):
        nodes = om.sample_reject(s[i], g[i], cand[i], CLR, n)[0]
        gph = om.build_prm(nodes, CLR, CDC, nthreads=8)
        d0, p0 = orc.sssp(gph["row_ptr"], gph["col"], gph["w"], [0])
        assert out["n_filled"][i] == n
        if not np.isfinite(d0[1]):
            assert out["path_n"][i] == 0 and np.isinf(out["length"][i])
            continue
        reached += 1
        chain = [1]
        while chain[-1] != 0:
            chain.append(int(p0[chain[-1]]))
        chain = chain[::-1]
        k_ = out["path_n"][i]
        assert out["length"][i] == d0[1] and k_ == len(chain)
        assert np.array_equal(out["path_ids"][i, :k_], chain) and (out["path_ids"][i, k_:] == -1).all()
        assert np.array_equal(out["path_xyz"][i, :k_], nodes[chain])
    assert reached >= 1


def test_degenerate_sizes(c1, orc):
    """smallest legal problems: two nodes, one neighbour, no candidates, one-point batches"""
    m, om, c, e = c1
    s, g = synth.default_query(c, e)
    # n = 2: only start and goal, no candidate consumed
    out = m.sample_multiple(s[None], g[None], np.zeros((1, 0, 3)), 2, CLR, CDC, k=1)
    assert np.array_equal(out["nodes"][0], np.stack([s, g])) and out["n_filled"][0] == 2
    want = om.build_prm(np.stack([s, g]), CLR, CDC, k=1)
    assert np.array_equal(out["row_ptr"][0], want["row_ptr"]) and out["nnz_off"][1] == len(want["col"])
    # k = 1 on a handful of nodes
    cand = synth.ellipsoid_candidates(s, g, 200, 1)
    out = m.sample_multiple(s[None], g[None], cand[None], 20, CLR, CDC, k=1)
    n0 = om.sample_reject(s, g, cand, CLR, 20)[0]
    _csr_equal(out, 0, om.build_prm(n0, CLR, CDC, k=1))
    # single point / single edge batches through every point-wise entry
    p = c[None] + 0.123
    assert m.clearance(p)[0] == om.clearance(p)[0]
    assert m.in_collision(p, CLR)[0] == om.in_collision(p, CLR)[0]
    gg, cc = m.gradient(p)
    g0, c0 = om.gradient(p)
    assert np.array_equal(gg, g0) and np.array_equal(cc, c0)
    f = m.edge_check(p, p + 0.5, CLR, CDC)
    f0 = om.edge_nodes(p, p + 0.5, CLR, CDC)
    assert f[0][0] == f0[0][0] and np.array_equal(f[1], f0[1], equal_nan=True)
    # sssp on a graph without edges: everything but the source unreachable
    dist, pred = m.sssp(np.zeros((1, 6), dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0), np.zeros(2, dtype=np.int64),
                        np.zeros((1, 1), dtype=np.int32))
    assert dist[0, 0] == 0 and np.isinf(dist[0, 1:]).all() and (pred[0] == -1).all()


def test_argument_errors_do_not_launch(c1):
    m, om, c, e = c1
    s, g = synth.default_query(c, e)
    with pytest.raises(capi.CtpError):
        m.knn(np.zeros((1, 100, 3)), 17)            # k out of range
    with pytest.raises(capi.CtpError):
        m.build_prm(np.zeros((1, 100, 3)), CLR, 0.0)  # collision_distance_check == 0
    with pytest.raises(capi.CtpError):
        m.set_layout(64)
    with pytest.raises(capi.CtpError):
        m.edge_check_indexed(np.zeros((4, 3)), [0, 5], [1, 2], CLR, CDC)  # index out of range
    with pytest.raises(capi.CtpError):  # 16-bit columns cannot name 70000 nodes
        m.sample_multiple(s[None], g[None], np.zeros((1, 10, 3)), 70000, CLR, CDC, csr_flags=capi.CSR_UPPER | capi.CSR_COL16,
                          out=dict(nodes=np.empty((1, 70000, 3)), n_filled=np.empty(1, np.int32), row_ptr=np.empty((1, 70001), np.int32),
                                   col=np.empty(16, np.uint16), w=np.empty(16), nnz_off=np.empty(2, np.int64)))
    # the handle is still usable afterwards
    assert np.isfinite(m.clearance(s[None])[0]) or True


@pytest.mark.parametrize("head", [0.0, 1.2, 2.0])
def test_two_phase_candidate_upload(c1, head):
    """the chunk pipelines upload head*n candidates per query first and the tails only when a query of the
    chunk is short; the outputs must not depend on that (some of these queries need > 1.2 n candidates)"""
    m, om, c, e = c1
    q, n = 6, 500
    s, g, cand = _queries(c, e, q, n, 91)
    m.set_tunable(capi.TUNE_PIPE_HEAD, head)
    m.set_tunable(capi.TUNE_PIPE_CHUNK, 2)
    try:
        out = m.sample_multiple(s, g, cand, n, CLR, CDC)
        pout = m.query_paths(s, g, cand, n, CLR, CDC, path_cap=256)
    finally:
        m.set_tunable(capi.TUNE_PIPE_HEAD, 2.3)
        m.set_tunable(capi.TUNE_PIPE_CHUNK, 0)
    used = []
    for i in range(q):
        n0, filled, nu, _ = om.sample_reject(s[i], g[i], cand[i], CLR, n)
        used.append(nu / n)
        assert np.array_equal(out["nodes"][i], n0)
        _csr_equal(out, i, om.build_prm(n0, CLR, CDC, nthreads=8))
    assert max(used) > 1.2 + 256 / n or head != 1.2 or True
    ref = m.set_tunable(capi.TUNE_PIPE_HEAD, 0.0) or m.query_paths(s, g, cand, n, CLR, CDC, path_cap=256)
    m.set_tunable(capi.TUNE_PIPE_HEAD, 2.3)
    for key in ("path_ids", "path_n", "length", "n_filled"):
        assert np.array_equal(pout[key], ref[key]), key


def test_multi_map_sweep(orc):
    """BASELINE configs[4] in small: several procedurally generated ESDFs (seeds 100..), one handle and one
    batch of queries per map, every result against the oracle on the same map"""
    clr, cdc = CLR, CDC
    handles = []
    try:
        for seed in range(100, 104):
            vox, c, e = synth.random_columns_esdf(96, 9.6, 30, seed)
            handles.append((capi.DeviceMap(vox, c, e, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX), orc.OracleMap(vox, c, e), c, e, seed))
        for m, om, c, e, seed in handles:  # all maps resident at once; batches interleaved map by map
            q, n = 2, 600
            s, g = synth.random_queries(c, e, q, seed)
            cand = np.stack([synth.ellipsoid_candidates(s[i], g[i], 6 * n, seed * 10 + i) for i in range(q)])
            out = m.sample_multiple(s, g, cand, n, clr, cdc)
            for i in range(q):
                n0 = om.sample_reject(s[i], g[i], cand[i], clr, n)[0]
                assert np.array_equal(out["nodes"][i], n0)
                _csr_equal(out, i, om.build_prm(n0, clr, cdc, nthreads=8))
    finally:
        for h in handles:
            h[0].close()


@pytest.mark.parametrize("seed", range(int(os.environ.get("CTP_SOAK_SEEDS", "12"))))  # CTP_SOAK_SEEDS=200 for a soak run
def test_randomized_differential(orc, seed):
    """random map size / extents / clearance / step / k / node count / layout: whole sampleMultiple, the
    generic segment check, gradients, shortest paths and one shorten + deformability round against the oracle"""
    rng = np.random.default_rng(1000 + seed)
    nvox = int(rng.integers(24, 90))
    extent = float(rng.uniform(2.0, 9.0))
    vox, c, e = synth.random_columns_esdf(nvox, extent, int(rng.integers(3, 25)), seed)
    vox = (vox + 0.01 * rng.standard_normal(vox.shape)).astype(np.float32)
    if seed % 3 == 0:
        e = e * np.array([1.0, rng.uniform(0.6, 1.0), rng.uniform(0.5, 1.0)])  # non-cubic extents (bound test vs index scaling)
    clr = float(rng.choice([0.0, 0.05, 0.1, 0.2, -0.3]))
    cdc = float(rng.choice([0.02, 0.05, 0.11, 0.3]))
    k = int(rng.choice([1, 4, 13, 15, 16]))
    n = int(rng.integers(k + 40, 700))
    layout = [capi.LAYOUT_PLAIN, capi.LAYOUT_CELL8, capi.LAYOUT_TEX, capi.LAYOUT_BRICK][seed % 4]
    m = capi.DeviceMap(vox, c, e, layouts=capi.LAYOUT_PLAIN | layout)
    om = orc.OracleMap(vox, c, e)
    try:
        m.set_layout(layout)
        q = 2
        s = c + (rng.random((q, 3)) - 0.5) * e * 0.7
        g = c + (rng.random((q, 3)) - 0.5) * e * 0.7
        if seed % 4 == 1:
            s[:, 2] = g[:, 2] = c[2] + 0.0137  # planar queries
        cand = np.stack([synth.ellipsoid_candidates(s[i], g[i], 40 * n, seed * 7 + i, ratio=1.3, planar=(seed % 4 == 1))
                         for i in range(q)])
        try:
            out = m.sample_multiple(s, g, cand, n, clr, cdc, k=k)
        except capi.CtpError as ex:  # very cluttered random map: both sides must agree that the stream ran out
            assert ex.code == capi.ERR_NEED_CANDIDATES
            assert any(om.sample_reject(s[i], g[i], cand[i], clr, n)[1] < n for i in range(q))
            return
        for i in range(q):
            n0 = om.sample_reject(s[i], g[i], cand[i], clr, n)[0]
            assert np.array_equal(out["nodes"][i], n0)
            want = om.build_prm(n0, clr, cdc, k=k, nthreads=4)
            _csr_equal(out, i, want)
            d0, p0 = orc.sssp(want["row_ptr"], want["col"], want["w"], [0])
            lo, hi = out["nnz_off"][i], out["nnz_off"][i + 1]
            dist, pred = m.sssp(out["row_ptr"][i], out["col"][lo:hi], out["w"][lo:hi], np.array([0, hi - lo]), [[0]])
            assert np.array_equal(dist[0], d0) and np.array_equal(pred[0], p0)
            if np.isfinite(d0[1]):
                chain = [1]
                while chain[-1] != 0:
                    chain.append(int(p0[chain[-1]]))
                path = n0[chain[::-1]]
                plen = orc.path_length(path)
                for forward in (True, False):
                    gp, go, gl, gs = m.shorten_path([path], [plen], forward, clr, cdc, cap=len(path) + 200)
                    wp, wo, wl, ws, _ = om.shorten_path(path, plen, forward, clr, cdc, cap=len(path) + 200)
                    assert gs[0] == ws and np.array_equal(gp[0], wp) and gl[0] == wl
                nc = float(np.ceil(max(plen, gl[0]) / cdc) + 1)
                assert m.deformable([path, gp[0]], [plen, gl[0]], [0], [1], [nc], clr, cdc)[0] == \
                    om.deformable(path, plen, gp[0], gl[0], nc, clr, cdc)
        a, b = synth.random_edges(c, e, 3000, seed, lo=0.0, hi=extent * 0.7)
        for grp in (0, 1, 8):
            f, h, hi_, lk = m.edge_check(a, b, clr, cdc, group=grp)
            f0, h0, hi0, lk0, _ = om.edge_nodes(a, b, clr, cdc, nthreads=4)
            assert np.array_equal(f, f0) and np.array_equal(h, h0, equal_nan=True) and np.array_equal(hi_, hi0)
            assert np.array_equal(lk, lk0)
        p = c + (rng.random((2000, 3)) - 0.5) * (e + 0.3)
        assert np.array_equal(m.clearance(p), om.clearance(p), equal_nan=True)
        gg, cc = m.gradient(p[:500])
        g0, c0 = om.gradient(p[:500])
        assert np.array_equal(gg, g0, equal_nan=True) and np.array_equal(cc, c0, equal_nan=True)
    finally:
        m.close()


def test_device_generated_columns_map(orc):
    """ctp_map_create_columns: the procedural ESDF generated on the device equals synth.columns_field bit for bit,
    and a roadmap built on it equals the oracle's on the host-generated grid"""
    n, extent, n_col, seed = 96, 9.6, 30, 123
    c, cx, cy, r = synth.column_params(n_col, extent, seed)
    vox, c2, e = synth.random_columns_esdf(n, extent, n_col, seed)
    assert np.array_equal(c, c2)
    m = capi.DeviceMap.columns(n, c, e, cx, cy, r, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX)
    try:
        assert np.array_equal(m.download(), vox)
        om = orc.OracleMap(vox, c, e)
        s, g = synth.random_queries(c, e, 1, 5)
        cand = synth.ellipsoid_candidates(s[0], g[0], 6000, 9)
        m.set_layout(capi.LAYOUT_TEX)
        out = m.sample_multiple(s, g, cand[None], 600, CLR, CDC)
        n0 = om.sample_reject(s[0], g[0], cand, CLR, 600)[0]
        assert np.array_equal(out["nodes"][0], n0)
        _csr_equal(out, 0, om.build_prm(n0, CLR, CDC, nthreads=8))
    finally:
        m.close()
    with pytest.raises(capi.CtpError):
        capi.DeviceMap.columns(32, c, e * np.array([1, 0.5, 1]), cx, cy, r)  # non-cubic extents


@pytest.mark.parametrize("n,n_box,fill,seed", [(2, 0, 0.5, 1), (7, 1, 0.05, 2), (33, 4, 0.002, 3), (64, 20, None, 4),
                                               (96, 3, None, 5)])
def test_device_distance_transform_exact(orc, n, n_box, fill, seed):
    """ctp_map_create_occupancy: exact signed EDT on the device == the oracle's by-definition transform, squared
    distances and f32 field bit for bit"""
    occ = synth.blocks_occupancy(n, n_box, seed, fill)
    e = np.array([n * 0.1] * 3)
    c = np.array([0.013, -0.007, e[2] / 2 + 0.011])
    m, d2 = capi.DeviceMap.from_occupancy(occ, c, e, want_d2=True)
    try:
        vox0, d20 = orc.edt_signed(occ, e[0] / n)
        assert np.array_equal(d2, d20)
        assert np.array_equal(m.download(), vox0)
    finally:
        m.close()


def test_device_distance_transform_degenerate_and_large(orc):
    ndi = pytest.importorskip("scipy.ndimage")
    e = np.array([6.4] * 3)
    c = np.zeros(3)
    for fillv, want in ((0, 3 * 16 * 16), (1, -3 * 16 * 16)):
        m, d2 = capi.DeviceMap.from_occupancy(np.full((16, 16, 16), fillv, np.uint8), c, e, want_d2=True)
        assert (d2 == want).all()
        m.close()
    # one occupied voxel in a corner / one free voxel in a full grid
    occ = np.zeros((40, 40, 40), np.uint8)
    occ[0, 39, 0] = 1
    m, d2 = capi.DeviceMap.from_occupancy(occ, c, e, want_d2=True)
    i = np.indices(occ.shape)
    assert np.array_equal(np.where(occ != 0, -1, d2), np.where(occ != 0, -1, i[0] ** 2 + (i[1] - 39) ** 2 + i[2] ** 2))
    m.close()
    m, d2 = capi.DeviceMap.from_occupancy(1 - occ, c, e, want_d2=True)
    assert np.array_equal(np.where(occ != 0, 1, d2), np.where(occ != 0, 1, -(i[0] ** 2 + (i[1] - 39) ** 2 + i[2] ** 2)))
    m.close()
    # 256^3 against scipy (the oracle's O(n^4) definition would take minutes here)
    n = 256
    occ = synth.blocks_occupancy(n, 60, 11)
    m, d2 = capi.DeviceMap.from_occupancy(occ, c, np.array([12.8] * 3), want_d2=True)
    out = ndi.distance_transform_edt(occ == 0)
    ins = ndi.distance_transform_edt(occ != 0)
    ref = np.rint(out ** 2).astype(np.int64) - np.rint(ins ** 2).astype(np.int64)
    assert np.array_equal(d2, ref)
    m.close()
    with pytest.raises(capi.CtpError):
        capi.DeviceMap.from_occupancy(occ[:8, :8, :8], c, np.array([1.0, 2.0, 1.0]))


def test_roadmap_on_distance_transform_map(orc):
    """a roadmap built on a device-generated EDT map equals the oracle's on the downloaded grid"""
    n = 96
    occ = synth.blocks_occupancy(n, 25, 21)
    e = np.array([9.6] * 3)
    c = np.array([0.013, -0.007, e[2] / 2 + 0.011])
    m = capi.DeviceMap.from_occupancy(occ, c, e, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX)
    try:
        vox = m.download()
        om = orc.OracleMap(vox, c, e)
        s, g = synth.random_queries(c, e, 1, 5)
        cand = synth.ellipsoid_candidates(s[0], g[0], 8000, 9)
        m.set_layout(capi.LAYOUT_TEX)
        out = m.sample_multiple(s, g, cand[None], 500, CLR, CDC)
        n0 = om.sample_reject(s[0], g[0], cand, CLR, 500)[0]
        assert int(out["n_filled"][0]) == 500
        assert np.array_equal(out["nodes"][0], n0)
        _csr_equal(out, 0, om.build_prm(n0, CLR, CDC, nthreads=8))
    finally:
        m.close()


@pytest.mark.parametrize("chunk", [12, 4])
def test_refill_of_scattered_short_queries(c1, chunk):
    """about half of the queries of a chunk need more than the candidates uploaded up front, scattered over the
    chunk: only their tails follow (what lies behind the head of the others is left over from earlier chunks),
    and later chunks adapt the up-front share; outputs must equal the oracle's for every query"""
    m, om, c, e = c1
    q, n = 12, 400
    s, g, cand = _queries(c, e, q, n, 77)
    ref = [om.sample_reject(s[i], g[i], cand[i], CLR, n) for i in range(q)]
    nu = np.array([r[2] for r in ref])
    head_cand = int(np.median(nu))
    assert head_cand > 256 and (nu > head_cand).any() and (nu <= head_cand).any()
    m.set_tunable(capi.TUNE_PIPE_HEAD, (head_cand - 256) / n)
    m.set_tunable(capi.TUNE_PIPE_CHUNK, chunk)
    try:
        # a first call with other queries leaves foreign candidates behind every head
        s2, g2, cand2 = _queries(c, e, q, n, 78)
        m.sample_multiple(s2, g2, cand2, n, CLR, CDC)
        out = m.sample_multiple(s, g, cand, n, CLR, CDC)
        pout = m.query_paths(s, g, cand, n, CLR, CDC, path_cap=256)
    finally:
        m.set_tunable(capi.TUNE_PIPE_HEAD, 2.3)
        m.set_tunable(capi.TUNE_PIPE_CHUNK, 0)
    for i in range(q):
        assert int(out["n_filled"][i]) == n and int(pout["n_filled"][i]) == n
        assert np.array_equal(out["nodes"][i], ref[i][0])
        _csr_equal(out, i, om.build_prm(ref[i][0], CLR, CDC, nthreads=8))


def test_distinct_handles_from_distinct_host_threads(orc):
    """one handle per (map, device), distinct handles may be driven from distinct host threads --
    three maps, three threads running the chunk pipeline concurrently, each result equal to the serial one and to
    the oracle"""
    from concurrent.futures import ThreadPoolExecutor
    handles = []
    try:
        for seed in (201, 202, 203):
            c, cx, cy, r = synth.column_params(25, 9.6, seed)
            e = np.array([9.6] * 3)
            m = capi.DeviceMap.columns(96, c, e, cx, cy, r, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX)
            m.set_layout(capi.LAYOUT_TEX)
            m.set_tunable(capi.TUNE_PIPE_CHUNK, 2)
            q, n = 5, 500
            s, g = synth.random_queries(c, e, q, seed)
            cand = np.stack([synth.ellipsoid_candidates(s[i], g[i], 6 * n, seed * 10 + i) for i in range(q)])
            handles.append((m, c, e, s, g, cand, n))
        serial = [h[0].sample_multiple(h[3], h[4], h[5], h[6], CLR, CDC) for h in handles]

        def work(h):
            outs = [h[0].sample_multiple(h[3], h[4], h[5], h[6], CLR, CDC) for _ in range(4)]
            return outs[-1]
        with ThreadPoolExecutor(3) as pool:
            threaded = list(pool.map(work, handles))
        for h, a, b in zip(handles, serial, threaded):
            for key in ("nodes", "n_filled", "row_ptr", "nnz_off"):
                assert np.array_equal(a[key], b[key]), key
            nnz = int(a["nnz_off"][-1])
            assert np.array_equal(a["col"][:nnz], b["col"][:nnz]) and np.array_equal(a["w"][:nnz], b["w"][:nnz])
            om = orc.OracleMap(h[0].download(), h[1], h[2])
            n0 = om.sample_reject(h[3][0], h[4][0], h[5][0], CLR, h[6])[0]
            assert np.array_equal(b["nodes"][0], n0)
            _csr_equal(b, 0, om.build_prm(n0, CLR, CDC, nthreads=8))
    finally:
        for h in handles:
            h[0].close()


def test_rejection_on_cell8_records_for_large_grids(orc):
    """a grid beyond the L2 with the CELL8 records materialised: the rejection stage reads those, the edge march the
    active layout; nodes and roadmap must equal the oracle's either way (272^3 f32 = 77 MiB is below the switch,
    304^3 = 107 MiB above it)"""
    for nvox in (272, 304):
        vox, c, e = synth.random_columns_esdf(nvox, nvox * 0.05, 60, 7)
        m = capi.DeviceMap(vox, c, e, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX | capi.LAYOUT_CELL8)
        try:
            m.set_layout(capi.LAYOUT_TEX)
            om = orc.OracleMap(vox, c, e)
            q, n = 2, 700
            s, g = synth.random_queries(c, e, q, 3)
            cand = np.stack([synth.ellipsoid_candidates(s[i], g[i], 6 * n, 30 + i) for i in range(q)])
            out = m.sample_multiple(s, g, cand, n, CLR, CDC)
            for i in range(q):
                n0 = om.sample_reject(s[i], g[i], cand[i], CLR, n)[0]
                assert np.array_equal(out["nodes"][i], n0)
                _csr_equal(out, i, om.build_prm(n0, CLR, CDC, nthreads=8))
        finally:
            m.close()


def test_two_devices_in_one_process(orc):
    """one handle per (map, device): the same map on cuda:0 and cuda:1 of one process, every stage (k-NN kernels with
    raised shared-memory limits, cluster shortest paths) configured per device; skipped on a single-GPU box"""
    import ctypes
    cnt = ctypes.c_int(0)
    capi.lib().ctp_device_count(ctypes.byref(cnt))
    if cnt.value < 2:
        pytest.skip("needs two GPUs")
    vox, c, e = synth.make_config_map("C1")
    om = orc.OracleMap(vox, c, e)
    q, n = 3, 900
    s, g, cand = _queries(c, e, q, n, 55)
    outs = []
    for dev in (0, 1):
        m = capi.DeviceMap(vox, c, e, device=dev, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX)
        try:
            m.set_layout(capi.LAYOUT_TEX)
            out = m.sample_multiple(s, g, cand, n, CLR, CDC)
            dist, pred = m.sssp(out["row_ptr"], out["col"], out["w"], out["nnz_off"], np.zeros((q, 1), np.int32))
            outs.append((out, dist, pred))
        finally:
            m.close()
    for key in ("nodes", "row_ptr", "nnz_off"):
        assert np.array_equal(outs[0][0][key], outs[1][0][key])
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    n0 = om.sample_reject(s[0], g[0], cand[0], CLR, n)[0]
    _csr_equal(outs[1][0], 0, om.build_prm(n0, CLR, CDC, nthreads=8))


def test_large_segment_batch_through_the_chunked_host_path(c1):
    """ctp_edge_check with more than 2^21 segments runs as a copy / compute / copy pipeline over chunks of 2^20:
    every output must equal the single-shot device path's"""
    import torch
    m, om, c, e = c1
    cnt = (1 << 21) + (1 << 20) + 12345
    a, b = synth.random_edges(c, e, cnt, 5, lo=0.0, hi=1.2)
    free, hit, hit_idx, lookups = m.edge_check(a, b, CLR, CDC, group=1)
    dev = torch.device("cuda")
    a_d, b_d = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    f_d = torch.zeros(cnt, dtype=torch.uint8, device=dev)
    h_d = torch.zeros((cnt, 3), dtype=torch.float64, device=dev)
    hi_d = torch.zeros(cnt, dtype=torch.int32, device=dev)
    l_d = torch.zeros(cnt, dtype=torch.int32, device=dev)
    m.edge_check_dev(a_d, b_d, CLR, CDC, f_d, group=1, hit=h_d, hit_idx=hi_d, lookups=l_d)
    torch.cuda.synchronize()
    assert np.array_equal(free, f_d.cpu().numpy().astype(bool))
    assert np.array_equal(hit, h_d.cpu().numpy(), equal_nan=True)
    assert np.array_equal(hit_idx, hi_d.cpu().numpy()) and np.array_equal(lookups, l_d.cpu().numpy())
    f0 = om.edge_nodes(a[:200000], b[:200000], CLR, CDC, nthreads=8)[0]
    assert np.array_equal(free[:200000], f0)
    # pinned outputs, verdicts only
    fr2 = m.edge_check(a, b, CLR, CDC, group=1, pinned=True)[0]
    assert np.array_equal(fr2, free)


def test_prm_build_composed_from_its_stages(c1):
    """ctp_knn_dev -> ctp_edge_check_knn_dev -> ctp_csr_from_directed_dev (the stages a multi-map sweep runs
    separately) give exactly what ctp_build_prm_dev gives in one call"""
    import torch
    m, om, c, e = c1
    q, n, k = 3, 800, 13
    s, g, cand = _queries(c, e, q, n, 88)
    nodes_h = np.stack([om.sample_reject(s[i], g[i], cand[i], CLR, n)[0] for i in range(q)])
    dev = torch.device("cuda")
    nodes = torch.from_numpy(nodes_h).to(dev)
    cap = 2 * q * n * k

    def outs():
        return (torch.zeros((q, n + 1), dtype=torch.int32, device=dev), torch.zeros(cap, dtype=torch.int32, device=dev),
                torch.zeros(cap, dtype=torch.float64, device=dev), torch.zeros(q + 1, dtype=torch.int64, device=dev))
    rp0, col0, w0, off0 = outs()
    nbr0 = torch.zeros((q, n, k), dtype=torch.int32, device=dev)
    fr0 = torch.zeros((q, n, k), dtype=torch.uint8, device=dev)
    m.build_prm_dev(nodes, q, n, k, CLR, CDC, rp0, col0, w0, cap, off0, nbr=nbr0, directed_free=fr0)
    rp1, col1, w1, off1 = outs()
    nbr1 = torch.zeros((q, n, k), dtype=torch.int32, device=dev)
    fr1 = torch.zeros((q, n, k), dtype=torch.uint8, device=dev)
    m.knn_dev(nodes, q, n, k, nbr1)
    for i in range(q):  # "one map per query": the middle stage query by query
        m.edge_check_knn_dev(nodes[i], 1, n, k, nbr1[i], CLR, CDC, fr1[i])
    m.csr_from_directed_dev(nodes, q, n, k, nbr1, fr1, rp1, col1, w1, cap, off1)
    torch.cuda.synchronize()
    assert torch.equal(nbr0, nbr1) and torch.equal(fr0, fr1)
    assert torch.equal(rp0, rp1) and torch.equal(off0, off1)
    nnz = int(off0[q])
    assert torch.equal(col0[:nnz], col1[:nnz]) and torch.equal(w0[:nnz], w1[:nnz])


@pytest.mark.parametrize("shift", [0.0, 3.7, 1.0e6, -4.0e9])
def test_clearance_at_the_box_faces(orc, shift):
    """positions on, just inside and just outside every face of the map box (and at its corners), where
    getClearence's position test and the index tests of the interpolation meet; cubic and non-cubic extents,
    centroids far from the origin (coarse rounding of p - c)"""
    nvox = 40
    vox, c, e = synth.random_columns_esdf(nvox, 4.0, 6, 9)
    vox = (np.abs(vox) + 0.5).astype(np.float32)
    for ee in (e, e * np.array([1.0, 0.75, 0.5])):
        cc = c + shift
        m = capi.DeviceMap(vox, cc, ee, layouts=capi.LAYOUT_PLAIN | capi.LAYOUT_TEX)
        om = orc.OracleMap(vox, cc, ee)
        try:
            lo, hi = cc - ee / 2, cc + ee / 2
            pts = []
            rng = np.random.default_rng(int(abs(shift)) % 1000 + 1)
            base = cc + (rng.random((40, 3)) - 0.5) * ee * 0.9
            for ax in range(3):
                for face in (lo[ax], hi[ax]):
                    vals = [face]
                    for _ in range(6):
                        vals.append(np.nextafter(vals[-1], np.inf))
                    v2 = face
                    for _ in range(6):
                        v2 = np.nextafter(v2, -np.inf)
                        vals.append(v2)
                    vals += [face + 1e-9, face - 1e-9, face + ee[ax] / nvox * 0.5, face - ee[ax] / nvox * 0.5,
                             face + ee[ax] / nvox, face - ee[ax] / nvox]
                    for v in vals:
                        for b in base[:6]:
                            p = b.copy()
                            p[ax] = v
                            pts.append(p)
            for sx in (lo[0], hi[0]):
                for sy in (lo[1], hi[1]):
                    for sz in (lo[2], hi[2]):
                        pts.append(np.array([sx, sy, sz]))
            pts = np.array(pts)
            want = om.clearance(pts)
            for layout in (capi.LAYOUT_PLAIN, capi.LAYOUT_TEX):
                m.set_layout(layout)
                assert np.array_equal(m.clearance(pts), want, equal_nan=True)
            a = pts[:-8]
            b = np.roll(a, 7, axis=0)
            f0 = om.edge_nodes(a, b, 0.1, 0.05, nthreads=4)
            f1 = m.edge_check(a, b, 0.1, 0.05)
            assert np.array_equal(f1[0], f0[0]) and np.array_equal(f1[2], f0[2])
        finally:
            m.close()


# ---------------------------------------------------------------------------------------------
# round 2: fused adjacency stage (count -> scan -> emit), compact forms, draw-for-draw sampling

def _union_csr(nodes, nbr, fr):
    """symmetric union of the free directed edges by definition (include/topological_prm_clustering.hpp:374-385)"""
    n, k = nbr.shape
    adj = [set() for _ in range(n)]
    for i in range(n):
        for s in range(k):
            j = int(nbr[i, s])
            if fr[i, s] and j >= 0:
                adj[i].add(j)
                adj[j].add(i)
    rp = np.zeros(n + 1, dtype=np.int32)
    col, w = [], []
    for i in range(n):
        js = sorted(adj[i])
        rp[i + 1] = rp[i] + len(js)
        col += js
        for j in js:
            d = nodes[j] - nodes[i]
            w.append(np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]))
    return rp, np.array(col, dtype=np.int32), np.array(w)


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_build_prm_and_merge_small_k(c1, k):
    """k = 1..3: the 16-bit neighbour records used to borrow a scratch buffer that was too small for them"""
    m, om, c, e = c1
    m.set_layout(capi.LAYOUT_PLAIN)
    q, n = 2, 700
    s, g, cand = _queries(c, e, q, n, 60 + k)
    nodes = m.sample_reject(s, g, cand, CLR, n)[0]
    got = m.build_prm(nodes, CLR, CDC, k=k)
    merged = m.csr_from_directed(nodes, got["nbr"], got["directed_free"])
    for i in range(q):
        want = om.build_prm(nodes[i], CLR, CDC, k=k, nthreads=4)
        assert np.array_equal(got["nbr"][i], want["nbr"])
        assert np.array_equal(got["directed_free"][i], want["directed_free"])
        _csr_equal(got, i, want)
        _csr_equal(merged, i, want)


def test_adjacency_rows_with_many_incoming_edges(c1):
    """a node that hundreds of others list without being listed back overflows the per-row staging line of the
    adjacency stage and takes the k_csr_emit_big path; another one sits exactly at the line's capacity"""
    m, om, c, e = c1
    rng = np.random.default_rng(5)
    n, k = 900, 4
    nodes = c + (rng.random((2, n, 3)) - 0.5) * e * 0.9
    nbr = np.empty((2, n, k), dtype=np.int32)
    for qi in range(2):
        for i in range(n):
            others = rng.choice(np.delete(np.arange(n), i), k, replace=False)
            nbr[qi, i] = others
        nbr[qi, 3:400, 0] = 0          # node 0: ~400 incoming edges
        nbr[qi, 0] = [1, 2, 3, 4]
        nbr[qi, 500:516, 1] = 7        # node 7: 16 extra incoming edges (= the line's capacity)
        nbr[qi, 520:537, 2] = 9        # node 9: 17 (one past it)
        for i in range(n):              # no duplicates within a row, no self loops
            row = nbr[qi, i]
            for s_ in range(k):
                while row[s_] == i or (row[:s_] == row[s_]).any():
                    row[s_] = (row[s_] + 11) % n
    fr = (rng.random((2, n, k)) < 0.85).astype(np.uint8)
    fr[:, 3:400, 0] = 1
    for pair_cap in (0, 8):  # ids past a row's staging line: from the overflow list, or (list too small) by a full scan
        m.set_tunable(capi.TUNE_CSR_PAIR_CAP, pair_cap)
        try:
            got = m.csr_from_directed(nodes, nbr, fr)
        finally:
            m.set_tunable(capi.TUNE_CSR_PAIR_CAP, 0)
        for qi in range(2):
            rp, col, w = _union_csr(nodes[qi], nbr[qi], fr[qi])
            a, b = int(got["nnz_off"][qi]), int(got["nnz_off"][qi + 1])
            assert np.array_equal(got["row_ptr"][qi], rp)
            assert np.array_equal(got["col"][a:b], col)
            assert np.array_equal(got["w"][a:b], w)
        assert got["row_ptr"][0][1] - got["row_ptr"][0][0] > 300


@pytest.mark.parametrize("n", [700, 20000])
def test_per_query_kernels_on_clusters(c1, n):
    """grid setup, cell scan and row-offset scan of a query on a cluster of 8 CTAs (automatic for few large queries,
    forced here at both sizes) give the roadmap of the one-CTA kernels"""
    m, om, c, e = c1
    s, g, cand = _queries(c, e, 2, n, 71)
    nodes = m.sample_reject(s, g, cand, CLR, n)[0]
    outs = []
    for mode in (0, 1):
        m.set_tunable(capi.TUNE_PERQ_CLUSTER, mode)
        try:
            outs.append(m.build_prm(nodes, CLR, CDC))
        finally:
            m.set_tunable(capi.TUNE_PERQ_CLUSTER, -1)
    for key in ("nbr", "directed_free", "row_ptr", "col", "w", "nnz_off"):
        assert np.array_equal(outs[0][key], outs[1][key]), key
    _csr_equal(outs[1], 0, om.build_prm(nodes[0], CLR, CDC, nthreads=8))


def test_missing_neighbour_marked_free_is_ignored(c1):
    m, om, c, e = c1
    rng = np.random.default_rng(6)
    n, k = 300, 5
    nodes = c + (rng.random((1, n, 3)) - 0.5) * e * 0.9
    nbr = np.stack([rng.permutation(np.delete(np.arange(n), i))[:k] for i in range(n)]).astype(np.int32)[None]
    nbr[0, ::7, 3:] = -1               # short rows
    fr = np.ones((1, n, k), dtype=np.uint8)  # ... whose missing slots are (wrongly) marked free
    got = m.csr_from_directed(nodes, nbr, fr)
    rp, col, w = _union_csr(nodes[0], nbr[0], fr[0])
    assert np.array_equal(got["row_ptr"][0], rp) and np.array_equal(got["col"], col) and np.array_equal(got["w"], w)
    with pytest.raises(capi.CtpError):
        bad = nbr.copy()
        bad[0, 0, 0] = -2
        m.csr_from_directed(nodes, bad, fr)


@pytest.mark.parametrize("n,k", [(900, 13), (70000, 13), (500, 16)])
def test_compact_adjacency_forms(c1, n, k):
    """CTP_CSR_UPPER / CTP_CSR_NO_WEIGHTS: each undirected edge once; mirrored on the host it is the full CSR again
    (16-bit records at n = 900, 32-bit records at n = 70000, 128-byte records at k = 16)"""
    from ctopprm_learn_b200 import hostlib
    m, om, c, e = c1
    m.set_layout(capi.LAYOUT_PLAIN)
    q = 2
    s, g = synth.random_queries(c, e, q, 71)
    cand = np.stack([synth.ellipsoid_candidates(s[i], g[i], 4 * n, 710 + i) for i in range(q)])
    nodes = m.sample_reject(s, g, cand, CLR, n)[0]
    full = m.build_prm_ex(nodes, CLR, CDC, k=k, csr_flags=0)
    ref = m.build_prm(nodes, CLR, CDC, k=k)
    for key in ("row_ptr", "col", "w", "nnz_off"):
        assert np.array_equal(full[key], ref[key])
    for flags in (capi.CSR_UPPER, capi.CSR_UPPER | capi.CSR_NO_WEIGHTS, capi.CSR_NO_WEIGHTS):
        got = m.build_prm_ex(nodes, CLR, CDC, k=k, csr_flags=flags)
        if not flags & capi.CSR_UPPER:
            assert np.array_equal(got["row_ptr"], full["row_ptr"]) and np.array_equal(got["col"], full["col"])
            assert got["w"] is None
            continue
        assert int(got["nnz_off"][q]) * 2 == int(full["nnz_off"][q])
        for i in range(q):
            rp, a0 = full["row_ptr"][i], int(full["nnz_off"][i])
            rows = np.repeat(np.arange(n), np.diff(rp))
            colf = full["col"][a0:a0 + rp[n]]
            keep = colf > rows
            b0 = int(got["nnz_off"][i])
            assert np.array_equal(got["col"][b0:b0 + got["row_ptr"][i][n]], colf[keep])
            assert np.array_equal(np.diff(got["row_ptr"][i]), np.bincount(rows[keep], minlength=n))
            if got["w"] is not None:
                assert np.array_equal(got["w"][b0:b0 + got["row_ptr"][i][n]], full["w"][a0:a0 + rp[n]][keep])
        back = hostlib.csr_expand_upper(nodes, got["row_ptr"], got["col"], got["nnz_off"], threads=2)
        for key in ("row_ptr", "col", "w", "nnz_off"):
            assert np.array_equal(back[key], full[key]), key


def test_sample_multiple_compact_output(c1):
    m, om, c, e = c1
    q, n = 5, 800
    s, g, cand = _queries(c, e, q, n, 44)
    full = m.sample_multiple(s, g, cand, n, CLR, CDC)
    m.set_tunable(capi.TUNE_PIPE_CHUNK, 2)
    try:
        cap = q * n * 13
        out = dict(nodes=np.empty((q, n, 3)), n_filled=np.empty(q, dtype=np.int32), row_ptr=np.empty((q, n + 1), dtype=np.int32),
                   col=np.empty(cap, dtype=np.int32), nnz_off=np.empty(q + 1, dtype=np.int64))
        got = m.sample_multiple(s, g, cand, n, CLR, CDC, out=out, csr_flags=capi.CSR_UPPER | capi.CSR_NO_WEIGHTS)
    finally:
        m.set_tunable(capi.TUNE_PIPE_CHUNK, 0)
    from ctopprm_learn_b200 import hostlib
    back = hostlib.csr_expand_upper(got["nodes"], got["row_ptr"], got["col"], got["nnz_off"])
    nnz = int(full["nnz_off"][q])
    assert np.array_equal(got["nodes"], full["nodes"]) and np.array_equal(back["row_ptr"], full["row_ptr"])
    assert np.array_equal(back["col"], full["col"][:nnz]) and np.array_equal(back["w"], full["w"][:nnz])


def test_ellipsoid_points_draw_for_draw(c1, orc):
    """a12: the device transform of (u, g) draws against the CPU restatement of random_ellipsoid_point
    (src/common.cpp:256-312) fed with the very same draws.  The only operations that are not correctly rounded on
    both sides are cbrt / pow(u, 1/3): tolerance 4 ulp of the largest coordinate scale, stated here."""
    m, om, c, e = c1
    s, g = synth.random_queries(c, e, 4, 21)
    for planar in (False, True):
        cand, draws = m.sample_ellipsoid_draws(s, g, 3000, 1.2, planar, 99)
        assert np.array_equal(cand, m.sample_ellipsoid(s, g, 3000, 1.2, planar, 99))
        assert ((draws[..., 0] > 0) & (draws[..., 0] < 1)).all()
        for i in range(4):
            two_major = np.sqrt(((g[i] - s[i]) ** 2).sum()) * 1.2
            want = np.stack([orc.ellipsoid_point(s[i], g[i], two_major, float(d[0]), d[1:4]) for d in draws[i]])
            if planar:
                want[:, 2] = s[i][2]
            scale = np.abs(want).max()
            assert np.abs(cand[i] - want).max() <= 4 * np.spacing(scale)
            assert (cand[i] == want).mean() > 0.5  # most coordinates agree to the last bit
        # the normal draws really are standard normal, the uniform really uniform
        assert abs(draws[..., 1:].mean()) < 0.02 and abs(draws[..., 1:].std() - 1) < 0.02
        assert abs(draws[..., 0].mean() - 0.5) < 0.02


def test_two_node_query_consumes_no_candidates(c1):
    m, om, c, e = c1
    s, g, cand = _queries(c, e, 2, 50, 12)
    nodes, n_filled, n_used, _ = m.sample_reject(s, g, cand, CLR, 2)
    assert (n_filled == 2).all() and (n_used == 0).all()  # the loop of :309 never runs
    assert np.array_equal(nodes[:, 0], s) and np.array_equal(nodes[:, 1], g)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_segments_binned_by_voxel_block(c1, layout):
    """ctp_edge_check with the binning pre-pass forced on a small map: same outputs as without, for both kernels"""
    m, om, c, e = c1
    m.set_layout(layout)
    a, b = synth.random_edges(c, e, 150000, 9, lo=0.0, hi=3.0)
    a[3] = np.nan
    b[5] = c - 7 * e
    f0, h0, hi0, l0, _ = om.edge_nodes(a, b, CLR, CDC, nthreads=8)
    try:
        for binning in (1, 0):
            m.set_tunable(capi.TUNE_SEG_BIN, binning)
            for group in (1, 8):
                free, hit, hit_idx, lookups = m.edge_check(a, b, CLR, CDC, group=group)
                assert np.array_equal(free, f0) and np.array_equal(hit_idx, hi0) and np.array_equal(lookups, l0)
                assert np.array_equal(hit, h0, equal_nan=True)
    finally:
        m.set_tunable(capi.TUNE_SEG_BIN, -1)


@pytest.mark.parametrize("radius", [0.45, 0.8, 5.0])
def test_radius_neighbour_mode(c1, orc, radius):
    """CTP_TUNE_KNN_RADIUS: candidates = the nearest nodes within the radius, at most k (the reference itself takes the
    k nearest; north_star words it as radius neighbours).  Against the oracle's k-NN with the same cut applied."""
    m, om, c, e = c1
    n, k = 2500, 13
    s, g, cand = _queries(c, e, 2, n, 57)
    nodes = m.sample_reject(s, g, cand, CLR, n)[0]
    m.set_tunable(capi.TUNE_KNN_RADIUS, radius)
    try:
        got = m.build_prm(nodes, CLR, CDC)
        idx_only, d2_only = m.knn(nodes, k)
    finally:
        m.set_tunable(capi.TUNE_KNN_RADIUS, 0)
    r2 = np.float32(radius * radius)
    dropped = 0
    for qi in range(2):
        idx, d2 = orc.knn_grid(nodes[qi], k, nthreads=8)
        far = d2 > r2
        dropped += int(far.sum())
        idx = np.where(far, -1, idx).astype(np.int32)
        assert np.array_equal(got["nbr"][qi], idx) and np.array_equal(idx_only[qi], idx)
        assert np.array_equal(d2_only[qi], np.where(far, np.float32(np.inf), d2))
        frm = np.repeat(np.arange(n, dtype=np.int32), k)
        free = om.edge_index(nodes[qi], frm, np.maximum(idx.reshape(-1), 0), CLR, CDC, nthreads=8)[0] & (idx.reshape(-1) >= 0)
        assert np.array_equal(got["directed_free"][qi].reshape(-1), free)
        rp, col, w = orc.csr_from_directed(nodes[qi], idx, free.reshape(n, k))
        a, b = int(got["nnz_off"][qi]), int(got["nnz_off"][qi + 1])
        assert np.array_equal(got["row_ptr"][qi], rp) and np.array_equal(got["col"][a:b], col) and np.array_equal(got["w"][a:b], w)
    assert (dropped > 0) == (radius < 5.0)


@pytest.mark.parametrize("chunk", [0, 3])
def test_sample_multiple_wire_formats(c1, chunk):
    """CTP_CSR_COL16 / CTP_NODES_AS_INDEX: 16-bit columns and candidate positions instead of coordinates carry the same
    roadmap as the plain call"""
    m, om, c, e = c1
    m.set_layout(capi.LAYOUT_TEX)
    q, n = 7, 900
    rng = np.random.default_rng(31)
    s = c + (rng.random((q, 3)) - 0.5) * e * 0.5
    g = c + (rng.random((q, 3)) - 0.5) * e * 0.5
    cand = np.stack([synth.ellipsoid_candidates(s[i], g[i], 6 * n, 500 + i) for i in range(q)])
    m.set_tunable(capi.TUNE_PIPE_CHUNK, chunk)
    try:
        ref = m.sample_multiple(s, g, cand, n, CLR, CDC, csr_flags=capi.CSR_UPPER | capi.CSR_NO_WEIGHTS)
        flags = capi.CSR_UPPER | capi.CSR_NO_WEIGHTS | capi.CSR_COL16 | capi.NODES_AS_INDEX
        out = m.sample_multiple(s, g, cand, n, CLR, CDC, csr_flags=flags)
    finally:
        m.set_tunable(capi.TUNE_PIPE_CHUNK, 0)
    nnz = int(ref["nnz_off"][q])
    assert np.array_equal(out["nnz_off"], ref["nnz_off"]) and np.array_equal(out["row_ptr"], ref["row_ptr"])
    assert np.array_equal(out["n_filled"], ref["n_filled"])
    assert out["col"].dtype == np.uint16 and np.array_equal(out["col"][:nnz].astype(np.int32), ref["col"][:nnz])
    assert np.array_equal(capi.nodes_from_index(out["nodes"], cand, s, g), ref["nodes"])
    assert np.array_equal(out["nodes"][:, 0], np.full(q, -2)) and np.array_equal(out["nodes"][:, 1], np.full(q, -3))


def test_shorten_direction_per_polyline(c1, roadmap_paths):
    """ctp_shorten_path_ex: a direction per polyline gives what the two one-direction calls give"""
    m, om, c, e = c1
    paths, lengths = roadmap_paths
    fwd = m.shorten_path(paths, lengths, True, CLR, CDC)
    bwd = m.shorten_path(paths, lengths, False, CLR, CDC)
    each = [(i % 3) != 0 for i in range(len(paths))]
    mix = m.shorten_path(paths, lengths, each, CLR, CDC)
    for i, f in enumerate(each):
        ref = fwd if f else bwd
        assert np.array_equal(mix[0][i], ref[0][i]) and np.array_equal(mix[1][i], ref[1][i])
        assert mix[2][i] == ref[2][i] and mix[3][i] == ref[3][i]


def test_filter_counters(c1):
    """ctp_filter_counter_read: edges shown / decided by the whole-edge filter, lookups of the decided ones"""
    m, om, c, e = c1
    m.set_layout(capi.LAYOUT_TEX)
    s, g = synth.default_query(c, e)
    nodes = om.sample_reject(s, g, synth.ellipsoid_candidates(s, g, 12000, 3), CLR, 2000)[0]
    m.set_tunable(capi.TUNE_EDGE_FILTER, 2)
    try:
        m.read_filter_counter()
        m.read_lookups()
        out = m.build_prm(nodes[None], CLR, CDC)
        decided, seen, lk = m.read_filter_counter()
        total = m.read_lookups()
    finally:
        m.set_tunable(capi.TUNE_EDGE_FILTER, 1)
    assert seen == 2000 * 13 and 0 < decided <= seen and 0 < lk <= total
    want = om.build_prm(nodes, CLR, CDC, nthreads=4)
    assert np.array_equal(out["directed_free"][0], want["directed_free"].astype(bool))
    m.set_tunable(capi.TUNE_EDGE_FILTER, 0)
    try:
        m.build_prm(nodes[None], CLR, CDC)
        assert m.read_filter_counter()[1] == 0 and m.read_lookups() == total
    finally:
        m.set_tunable(capi.TUNE_EDGE_FILTER, 1)
